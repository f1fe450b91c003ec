"""ctypes/numpy front-end of the CPU oracle (oracle/surfel_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package (g4splat_b200/) never imports it.

The call sequence mirrors the reference host orchestration
(cuda_rasterizer/rasterizer_impl.cu:198-342 forward, :346-448 backward):
project -> count/bin/sort -> blend, and blend-backward -> project-backward.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"
_SRC = _HERE / "surfel_oracle.c"


def build(force: bool = False) -> None:
    """Compile the fp32 and fp64 oracle libraries with gcc (seconds)."""
    _BUILD.mkdir(exist_ok=True)
    for name, real in (("f32", "float"), ("f64", "double")):
        out = _BUILD / f"liboracle_{name}.so"
        if not force and out.exists() and out.stat().st_mtime >= _SRC.stat().st_mtime:
            continue
        cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared",
               f"-DORC_REAL={real}", str(_SRC), "-o", str(out), "-lm"]
        subprocess.run(cmd, check=True)


def _ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One precision of the oracle.  All arrays are numpy, C-contiguous."""

    def __init__(self, precision: str = "f32"):
        assert precision in ("f32", "f64")
        build()
        self.lib = C.CDLL(str(_BUILD / f"liboracle_{precision}.so"))
        self.rt = np.float32 if precision == "f32" else np.float64
        self.creal = C.c_float if precision == "f32" else C.c_double
        assert self.lib.orc_real_size() == np.dtype(self.rt).itemsize
        self.lib.orc_count_instances.restype = C.c_longlong

    def set_threads(self, n: int) -> int:
        """n=1 makes the backward accumulation order deterministic; returns the thread count."""
        return int(self.lib.orc_set_threads(C.c_int(n)))

    def _a(self, x, shape=None):
        if x is None:
            return None
        a = np.ascontiguousarray(np.asarray(x, dtype=self.rt))
        if shape is not None:
            a = a.reshape(shape)
        return a

    # -------------------------------------------------------------------------------- forward
    def forward(self, *, means3D, opacities, view, proj, campos, W, H, tanfovx, tanfovy, bg,
                shs=None, colors_precomp=None, scales=None, rotations=None, transMat_precomp=None,
                sh_degree=0, scale_modifier=1.0, prefiltered=False):
        rt = self.rt
        means3D = self._a(means3D)
        P = means3D.shape[0]
        opacities = self._a(opacities, (P,))
        view = self._a(view, (16,))
        proj = self._a(proj, (16,))
        campos = self._a(campos, (3,))
        bg = self._a(bg, (3,))
        shs = self._a(shs)
        colors_precomp = self._a(colors_precomp)
        scales = self._a(scales)
        rotations = self._a(rotations)
        transMat_precomp = self._a(transMat_precomp)
        M = 0 if shs is None else shs.shape[1]
        st = dict(P=P, W=W, H=H, M=M, D=sh_degree, tanfovx=tanfovx, tanfovy=tanfovy)
        st["radii"] = np.zeros(P, np.int32)
        st["means2D"] = np.zeros((P, 2), rt)
        st["depths"] = np.zeros(P, rt)
        st["transMats"] = np.zeros((P, 9), rt)
        st["rgb"] = np.zeros((P, 3), rt)
        st["normal_opacity"] = np.zeros((P, 4), rt)
        st["clamped"] = np.zeros((P, 3), np.uint8)
        st["tiles_touched"] = np.zeros(P, np.uint32)
        viol = self.lib.orc_project(
            C.c_int(P), C.c_int(sh_degree), C.c_int(M), _ptr(means3D), _ptr(scales),
            self.creal(scale_modifier), _ptr(rotations), _ptr(opacities), _ptr(shs),
            _ptr(transMat_precomp), _ptr(colors_precomp), _ptr(view), _ptr(proj), _ptr(campos),
            C.c_int(W), C.c_int(H), C.c_int(int(prefiltered)), _ptr(st["radii"]), _ptr(st["means2D"]),
            _ptr(st["depths"]), _ptr(st["transMats"]), _ptr(st["rgb"]), _ptr(st["normal_opacity"]),
            _ptr(st["clamped"]), _ptr(st["tiles_touched"])) if P else 0
        if viol:
            raise RuntimeError("Point is filtered although prefiltered is set")
        gx, gy = (W + 15) // 16, (H + 15) // 16
        R = int(self.lib.orc_count_instances(C.c_int(P), _ptr(st["tiles_touched"]))) if P else 0
        st["num_rendered"] = R
        st["point_list"] = np.zeros(max(R, 1), np.uint32)
        st["ranges"] = np.zeros((gx * gy, 2), np.uint32)
        if P:
            rc = self.lib.orc_bin(C.c_int(P), C.c_int(W), C.c_int(H), _ptr(st["means2D"]), _ptr(st["depths"]),
                                  _ptr(st["radii"]), C.c_longlong(R), _ptr(st["point_list"]), _ptr(st["ranges"]))
            assert rc == 0, rc
        st["features"] = colors_precomp if colors_precomp is not None else st["rgb"]
        st["transMat_used"] = transMat_precomp if transMat_precomp is not None else st["transMats"]
        st["out_color"] = np.zeros((3, H, W), rt)
        st["out_others"] = np.zeros((7, H, W), rt)
        st["final_T"] = np.zeros((3, H, W), rt)
        st["n_contrib"] = np.zeros((2, H, W), np.uint32)
        st["pairs_per_pixel"] = np.zeros((H, W), np.uint32)
        if P:
            self.lib.orc_set_pair_counter(_ptr(st["pairs_per_pixel"]))
            self.lib.orc_blend_forward(
                C.c_int(W), C.c_int(H), _ptr(st["ranges"]), _ptr(st["point_list"]), _ptr(st["means2D"]),
                _ptr(st["features"]), _ptr(st["transMat_used"]), _ptr(st["normal_opacity"]), _ptr(bg),
                _ptr(st["out_color"]), _ptr(st["out_others"]), _ptr(st["final_T"]), _ptr(st["n_contrib"]))
            self.lib.orc_set_pair_counter(None)
        # keep inputs for backward
        st.update(_means3D=means3D, _shs=shs, _colors_precomp=colors_precomp, _scales=scales,
                  _rotations=rotations, _transMat_precomp=transMat_precomp, _view=view, _proj=proj,
                  _campos=campos, _bg=bg)
        return st

    # ------------------------------------------------------------------------------- backward
    def backward(self, st, dL_dcolor, dL_dothers):
        rt = self.rt
        P, W, H, M, D = st["P"], st["W"], st["H"], st["M"], st["D"]
        dL_dcolor = self._a(dL_dcolor, (3, H, W))
        dL_dothers = self._a(dL_dothers, (7, H, W))
        g = dict(
            dL_dmeans3D=np.zeros((P, 3), rt), dL_dmeans2D=np.zeros((P, 3), rt),
            dL_dcolors=np.zeros((P, 3), rt), dL_dnormal=np.zeros((P, 3), rt),
            dL_dopacity=np.zeros((P, 1), rt), dL_dtransMat=np.zeros((P, 9), rt),
            dL_dsh=np.zeros((P, M, 3), rt), dL_dscales=np.zeros((P, 2), rt),
            dL_drotations=np.zeros((P, 4), rt))
        if P == 0:
            g["blend_dL_dtransMat"] = g["dL_dtransMat"].copy()
            g["blend_dL_dmean2D"] = g["dL_dmeans2D"].copy()
            return g
        self.lib.orc_blend_backward(
            C.c_int(W), C.c_int(H), _ptr(st["ranges"]), _ptr(st["point_list"]), _ptr(st["_bg"]),
            _ptr(st["means2D"]), _ptr(st["normal_opacity"]), _ptr(st["transMat_used"]), _ptr(st["features"]),
            _ptr(st["final_T"]), _ptr(st["n_contrib"]), _ptr(dL_dcolor), _ptr(dL_dothers),
            _ptr(g["dL_dtransMat"]), _ptr(g["dL_dmeans2D"]), _ptr(g["dL_dnormal"]), _ptr(g["dL_dopacity"]),
            _ptr(g["dL_dcolors"]))
        # stage-level observables of the blend backward (before the projection backward rewrites them)
        g["blend_dL_dtransMat"] = g["dL_dtransMat"].copy()
        g["blend_dL_dmean2D"] = g["dL_dmeans2D"].copy()
        self.lib.orc_project_backward(
            C.c_int(P), C.c_int(D), C.c_int(M), _ptr(st["_means3D"]), _ptr(st["transMat_used"]),
            _ptr(st["radii"]), _ptr(st["_shs"]), _ptr(st["clamped"]), _ptr(st["_scales"]),
            _ptr(st["_rotations"]), _ptr(st["_view"]), _ptr(st["_proj"]), C.c_int(W), C.c_int(H),
            C.c_float(st["tanfovx"]), C.c_float(st["tanfovy"]), _ptr(st["_campos"]), _ptr(g["dL_dtransMat"]),
            _ptr(g["dL_dnormal"]), _ptr(g["dL_dcolors"]), _ptr(g["dL_dsh"]), _ptr(g["dL_dmeans2D"]),
            _ptr(g["dL_dmeans3D"]), _ptr(g["dL_dscales"]), _ptr(g["dL_drotations"]))
        return g

    def mark_visible(self, means3D, view, proj):
        means3D = self._a(means3D)
        P = means3D.shape[0]
        out = np.zeros(P, np.uint8)
        if P:
            self.lib.orc_mark_visible(C.c_int(P), _ptr(means3D), _ptr(self._a(view, (16,))),
                                      _ptr(self._a(proj, (16,))), _ptr(out))
        return out.astype(bool)


