"""Run on the GPU box: parity tables (B200 operator vs CPU oracle vs reference extension) and
per-stage timings; writes gpurun_out/gpu_check.json.  Not part of the product path.

  python tests/tools/gpu_check.py [--cases c0,c0_bg,...] [--time c1,c2]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import torch  # noqa: E402

import helpers as Hh  # noqa: E402
from g4splat_b200 import synthetic as S  # noqa: E402


def load_ref():
    if os.environ.get("G4S_NO_REF"):
        return None
    try:
        from oracle import build_ref
        if not build_ref.up_to_date():
            return None
        return build_ref.import_reference()
    except Exception as ex:  # noqa: BLE001
        print("reference extension unavailable:", ex)
        return None


def fmt(rows):
    out = []
    for k, r in rows.items():
        if "mismatch" in r:
            out.append(f"    {k:28s} mismatch {r['mismatch']}/{r['n']}")
        else:
            out.append(f"    {k:28s} max_rel {r['max_rel']:.3e}  bad_frac {r['bad_frac']:.3e}  scale {r['scale']:.3e}")
    return "\n".join(out)


def time_operator(mod, case, iters=10, warmup=3):
    dev = "cuda"
    sc = case.scene
    t = lambda a, rg=True: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev).requires_grad_(rg)
    means3D, opac, shs, scales, rots = t(sc["means3D"]), t(sc["opacities"]), t(sc["shs"]), t(sc["scales"]), t(sc["rotations"])
    rast = mod.GaussianRasterizer(raster_settings=Hh.make_settings(mod, case, dev))
    gc, go = case.upstream()
    gc, go = torch.from_numpy(gc).to(dev), torch.from_numpy(go).to(dev)
    fw, bw = [], []
    for it in range(warmup + iters):
        means2D = torch.zeros_like(means3D, requires_grad=True)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize()
        e0.record()
        color, radii, allmap = rast(means3D=means3D, means2D=means2D, opacities=opac, shs=shs, scales=scales, rotations=rots)
        e1.record()
        torch.autograd.backward([color, allmap], [gc, go])
        e2.record()
        torch.cuda.synchronize()
        if it >= warmup:
            fw.append(e0.elapsed_time(e1))
            bw.append(e1.elapsed_time(e2))
        for x in (means3D, opac, shs, scales, rots):
            x.grad = None
    return float(np.median(fw)), float(np.median(bw)), int((radii > 0).sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="tiny,c0,c0_bg,c0_deg0,c0_precomp,ragged,scalemod")
    ap.add_argument("--time", default="c1,c2")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "gpu_check.json"))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    import g4splat_b200.diff_surfel_rasterization as b200
    from oracle.oracle import Oracle
    ref = load_ref()
    o32 = Oracle("f32")
    o32.set_threads(1)
    results = {"device": torch.cuda.get_device_name(0), "parity": {}, "timing": {}}

    def build(name):
        if name in Hh.NAMED_CASES:
            return Hh.named_case(name, o32)
        cfg = S.CONFIGS[name]
        return Hh.room_case(name, P=cfg["P"], W=cfg["W"], H=cfg["H"], seed=cfg["seed"], cams=cfg["cams"])

    for name in [c for c in args.cases.split(",") if c]:
        case = build(name)
        print(f"== case {name}: P={case.P} {case.cam.W}x{case.cam.H}")
        t0 = time.time()
        orc = Hh.run_oracle(o32, case)
        print(f"  oracle f32: {time.time() - t0:.2f} s, visible {(orc['radii'] > 0).sum()}, R_ref {orc['_state']['num_rendered']}")
        mine = Hh.run_operator(b200, case)
        rows = Hh.report(mine, orc, Hh.FWD_KEYS + Hh.GRAD_KEYS)
        print("  b200 vs oracle\n" + fmt(rows))
        results["parity"][name] = {"b200_vs_oracle": rows}
        if ref is not None:
            r = Hh.run_operator(ref, case)
            rows = Hh.report(mine, r, Hh.FWD_KEYS + Hh.GRAD_KEYS)
            print("  b200 vs reference\n" + fmt(rows))
            results["parity"][name]["b200_vs_reference"] = rows
            rows = Hh.report(orc, r, Hh.FWD_KEYS + Hh.GRAD_KEYS)
            print("  oracle vs reference\n" + fmt(rows))
            results["parity"][name]["oracle_vs_reference"] = rows
            r2 = Hh.run_operator(ref, case)
            rows = Hh.report(r2, r, Hh.FWD_KEYS + Hh.GRAD_KEYS)
            print("  reference vs reference (noise floor)\n" + fmt(rows))
            results["parity"][name]["reference_noise"] = rows

    for name in [c for c in args.time.split(",") if c]:
        case = build(name)
        print(f"== timing {name}: P={case.P} {case.cam.W}x{case.cam.H}")
        f, b, vis = time_operator(b200, case)
        print(f"  b200      fwd {f:.3f} ms  bwd {b:.3f} ms  total {f + b:.3f} ms  -> {case.P / (f + b) * 1e3 / 1e6:.1f} M Gaussians/s (visible {vis})")
        results["timing"][name] = {"b200": {"fwd_ms": f, "bwd_ms": b}}
        if ref is not None:
            f2, b2, _ = time_operator(ref, case)
            print(f"  reference fwd {f2:.3f} ms  bwd {b2:.3f} ms  total {f2 + b2:.3f} ms  -> {case.P / (f2 + b2) * 1e3 / 1e6:.1f} M Gaussians/s   speedup {(f2 + b2) / (f + b):.2f}x")
            results["timing"][name]["reference"] = {"fwd_ms": f2, "bwd_ms": b2}
            mine = Hh.run_operator(b200, case)
            r = Hh.run_operator(ref, case)
            rows = Hh.report(mine, r, Hh.FWD_KEYS + Hh.GRAD_KEYS)
            print("  b200 vs reference\n" + fmt(rows))
            results["parity"][name] = {"b200_vs_reference": rows}
    with open(args.out, "w") as fh:
        json.dump(results, fh, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
