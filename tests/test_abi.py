"""The C-ABI library loads and exports every symbol include/g4s_rasterizer.h declares (no
compute calls: this runs without a GPU)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "g4s_rasterizer.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(g4s_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for needed in ("g4s_forward_plan", "g4s_forward_render", "g4s_backward", "g4s_mark_visible",
                   "g4s_last_error", "g4s_version", "g4s_geom_bytes", "g4s_image_bytes", "g4s_binning_bytes"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    from g4splat_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_signatures_cover_the_header():
    from g4splat_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_and_sizes():
    from g4splat_b200 import _lib
    lib = _lib.load()
    assert lib.g4s_version() == 100
    # 112-byte record + depth + count + rect + clamp mask + visible-list slot per Gaussian, 256-byte aligned sections
    g = lib.g4s_geom_bytes(1000)
    assert 1000 * (112 + 4 + 4 + 8 + 1 + 4) <= g <= 1000 * (112 + 4 + 4 + 8 + 1 + 4) + 6 * 256
    assert lib.g4s_geom_bytes(0) > 0
    n = 1920 * 1080
    assert lib.g4s_image_bytes(1920, 1080) >= n * 20
    assert lib.g4s_binning_bytes(1 << 20) >= (1 << 20) * 44
    assert lib.g4s_backward_scratch_bytes(1000) >= 1000 * 80
    assert lib.g4s_launch_count() >= 0


def test_error_reporting_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on the CPU."""
    from g4splat_b200 import _lib
    lib = _lib.load()
    rc = lib.g4s_forward_plan(-1, 0, 0, 16, 16, None, None, None, None, None, 1.0, None, None, None, None, None,
                              1.0, 1.0, 0, None, None, None, None, None, 0)
    assert rc == -1
    assert b"bad P/W/H" in lib.g4s_last_error()
    rc = lib.g4s_mark_visible(-5, None, None, None, None, None)
    assert rc == -1


def test_round2_entry_points_validate_their_arguments_without_gpu():
    from g4splat_b200 import _lib
    lib = _lib.load()
    assert lib.g4s_backward_scratch_bytes(1000) >= 1000 * 96          # 21 sums per Gaussian, 6 x float4
    assert lib.g4s_set_fast_math(1) == 0 and lib.g4s_get_fast_math() == 1 and lib.g4s_set_fast_math(0) == 1
    assert lib.g4s_multimem_allreduce(None, 6, None, 0, 0, 2, None) == -1 and b"multiple of 4" in lib.g4s_last_error()
    assert lib.g4s_multimem_allreduce(None, 8, None, 0, 3, 2, None) == -1
    assert lib.g4s_multimem_allreduce(None, 0, None, 0, 0, 1, None) == 0          # nothing to do
    assert lib.g4s_densify_classify(-1, None, None, None, None, 0.0, 0.0, 0.0, 0.0, 2, None, None) == -1
    assert lib.g4s_densify_classify(0, None, None, None, None, 0.0, 0.0, 0.0, 0.0, 2, None, None) == 0
    assert lib.g4s_densify_gather(5, 45, None, None, None, None, 2, None, None, None) == -1
    assert lib.g4s_normal2curv_forward(16, 16, None, None, None, None, None) == -1
    assert lib.g4s_normal2curv_forward(0, 16, None, None, None, None, None) == 0
    assert lib.g4s_depth_order_forward(0, 4, None, None, None, 1.0, 1, 0, 20.0, None, None, None) == -1
    assert lib.g4s_forward_bin(-1, 16, 16, None, None, None, 0, None, 0) == -1
    assert lib.g4s_forward_blend(1, 16, 16, None, None, None, None, 0, None, None, None, 0) == -1
