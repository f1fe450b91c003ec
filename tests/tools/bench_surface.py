"""Times render()'s post-processing at 1920x1080 on the GPU: the fused kernels
(g4splat_b200.surface) against the reference's operator sequence (the torch restatement in
oracle/surface_oracle.py run on CUDA tensors = what gaussian_renderer/__init__.py:118-164 launches).
Forward + backward with the gradients a training step uses (normals, distortion, depth).

    python tests/tools/bench_surface.py [--iters 50]      # prints one JSON line
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--P", type=int, default=200_000)
    args = ap.parse_args()
    import helpers as Hh
    import g4splat_b200.diff_surfel_rasterization as op
    from g4splat_b200.surface import surface_attributes
    from oracle import surface_oracle as SO
    case = Hh.room_case("surface_bench", P=args.P, W=1920, H=1080, seed=2, cams=1)
    allmap = torch.tensor(Hh.run_operator(op, case, backward=False)["allmap"], device="cuda")
    V, FP = torch.tensor(case.cam.viewmatrix, device="cuda"), torch.tensor(case.cam.projmatrix, device="cuda")
    used = ("rend_normal", "surf_normal", "rend_dist", "surf_depth", "rend_alpha")
    gen = torch.Generator(device="cuda").manual_seed(1)
    up = {k: torch.randn((c, 1080, 1920), device="cuda", generator=gen) / (1920 * 1080)
          for k, c in zip(SO.KEYS, (1, 3, 3, 1, 1, 3, 3, 1)) if k in used}

    def step(fn):
        am = allmap.clone().requires_grad_(True)
        out = fn(am, V, FP, 1.0)
        torch.autograd.backward([out[k] for k in used], [up[k] for k in used])
        return am.grad

    def timed(fn):
        for _ in range(5):
            step(fn)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            step(fn)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.iters

    ms_fused = timed(surface_attributes)
    ms_torch = timed(SO.surface_attributes)
    g_f, g_t = step(surface_attributes), step(SO.surface_attributes)
    ok = torch.isfinite(g_t)
    err = float((g_f - g_t)[ok].abs().max() / g_t[ok].abs().max())
    N = 1920 * 1080
    alg = N * 4 * (7 + 15) + N * 4 * (sum(up[k].shape[0] for k in used) + 7)   # fwd: read 7 write 15; bwd: read grads (+ allmap, cached) write 7
    print(json.dumps({"workload": "render() post-processing, 1920x1080, forward + backward (normals, distortion, depth, alpha)",
                      "fused_ms": ms_fused, "reference_torch_ops_ms": ms_torch, "speedup": ms_torch / ms_fused,
                      "alg_bytes": alg, "fused_GBps": alg / (ms_fused * 1e-3) / 1e9,
                      "max_rel_grad_diff_vs_torch_fp32": err, "iters": args.iters}))


if __name__ == "__main__":
    main()
