"""g4splat_b200 -- B200-native (sm_100a) differentiable 2D-Gaussian (surfel) rasterizer.

The package is the hot path of DaLi-Jack/G4Splat and nothing else:

  g4splat_b200.diff_surfel_rasterization   drop-in operator API (GaussianRasterizer, ...)
  g4splat_b200.csrc                        hand-written CUDA kernels + the C ABI (include/*.h)
  g4splat_b200.view_parallel               view-sharded data parallelism (NCCL all-reduce)
  g4splat_b200.synthetic                   seeded synthetic scenes / cameras for tests + bench

`install()` makes `import diff_surfel_rasterization` resolve to the B200 operator so that the
reference's `gaussian_renderer.render()` and training / rendering scripts run unchanged.
"""
from __future__ import annotations

import sys

__version__ = "0.1.0"


def install(name: str = "diff_surfel_rasterization"):
    """Register the B200 operator module under the reference's import name."""
    from . import diff_surfel_rasterization as mod
    sys.modules[name] = mod
    return mod
