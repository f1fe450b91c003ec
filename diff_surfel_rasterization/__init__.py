"""Import-name shim: with the repository root on PYTHONPATH, `import diff_surfel_rasterization`
(what 2DGS/gaussian_renderer/__init__.py:14 and matcha/dm_scene/gaussians.py:22-23 do) resolves to
the B200 operator, so the reference's training / rendering scripts run unchanged:

    PYTHONPATH=/path/to/this/repo python train_with_refine_depth.py ...

Everything is re-exported from g4splat_b200.diff_surfel_rasterization (same names as the reference
module: RAST/diff_surfel_rasterization/__init__.py)."""
from g4splat_b200.diff_surfel_rasterization import (  # noqa: F401
    GaussianRasterizationSettings,
    GaussianRasterizer,
    _RasterizeGaussians,
    rasterize_gaussians,
    set_gradient_sink,
)
