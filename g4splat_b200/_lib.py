"""ctypes binding of include/g4s_rasterizer.h.  The product path has NO fallback: if the CUDA
library is missing or fails to load, importing the operator raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libg4s_rasterizer.so"

_vp, _i, _f, _i64, _sz, _d = C.c_void_p, C.c_int, C.c_float, C.c_int64, C.c_size_t, C.c_double

# name -> (restype, argtypes); mirrors include/g4s_rasterizer.h one to one
SIGNATURES = {
    "g4s_version": (_i, []),
    "g4s_last_error": (C.c_char_p, []),
    "g4s_set_fast_math": (_i, [_i]),
    "g4s_get_fast_math": (_i, []),
    "g4s_launch_count": (_i64, []),
    "g4s_geom_bytes": (_sz, [_i]),
    "g4s_image_bytes": (_sz, [_i, _i]),
    "g4s_binning_bytes": (_sz, [_i64]),
    "g4s_backward_scratch_bytes": (_sz, [_i]),
    "g4s_forward_plan": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp,
                              _f, _f, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "g4s_forward_render": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i]),
    "g4s_forward_bin": (_i, [_i, _i, _i, _vp, _vp, _vp, _i64, _vp, _i]),
    "g4s_forward_blend": (_i, [_i, _i, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i]),
    "g4s_backward": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f,
                          _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i]),
    "g4s_backward_scratch_bytes_raw": (_sz, [_i]),
    "g4s_forward_plan_raw": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp,
                                  _f, _f, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "g4s_backward_raw": (_i, [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _f, _f,
                              _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i]),
    "g4s_mark_visible": (_i, [_i, _vp, _vp, _vp, _vp, _vp]),
    "g4s_densify_stats": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4s_densify_stats_multimem": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4s_photometric_forward": (_i, [_i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp]),
    "g4s_photometric_backward": (_i, [_i, _i, _i, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp]),
    "g4s_densify_classify": (_i, [_i, _vp, _vp, _vp, _vp, _f, _f, _f, _f, _i, _vp, _vp]),
    "g4s_densify_gather": (_i, [_i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "g4s_multimem_allreduce": (_i, [_vp, _i64, _vp, _i64, _i, _i, _vp]),
    "g4s_normal2curv_forward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4s_normal2curv_backward": (_i, [_i, _i, _vp, _vp, _vp, _vp, _vp]),
    "g4s_depth_order_forward": (_i, [_i, _i, _vp, _vp, _vp, _f, _i, _i, _f, _vp, _vp, _vp]),
    "g4s_depth_order_backward": (_i, [_i, _i, _vp, _vp, _vp, _f, _i, _i, _f, _vp, _vp, _f, _vp, _vp]),
    "g4s_mip_filter": (_i, [_i, _vp, _i, _vp, _f, _f, _f, _vp, _vp, _vp]),
    "g4s_profile_enable": (_i, [_i]),
    "g4s_profile_select": (_i, [C.c_uint]),
    "g4s_profile_num_stages": (_i, []),
    "g4s_profile_stage_name": (C.c_char_p, [_i]),
    "g4s_profile_read": (_i, [_vp, _vp, _i]),
    "g4s_surface_forward": (_i, [_i, _i, _vp, _vp, _vp, _d] + [_vp] * 8 + [_vp]),
    "g4s_surface_backward": (_i, [_i, _i, _vp, _vp, _vp, _d] + [_vp] * 8 + [_vp, _vp]),
    "g4s_debug_decode_geom": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "g4s_debug_decode_lists": (_i, [_i, _i, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "g4s_debug_pair_stats": (_i, [_i, _i, _vp, _i, _vp, _vp, _i64, _vp, _vp]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m g4splat_b200.build` "
            "(there is no CPU or PyTorch fallback for the rasterizer)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().g4s_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"g4s_rasterizer error {rc}: {msg}")
