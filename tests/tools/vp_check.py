"""N ranks over NCCL on real GPUs: view-sharded gradients of the B200 operator (kernel-side gradient sink)
brought together by both transports of g4splat_b200.view_parallel must equal ONE rank looping over the same
views, within 1e-4 of each block's largest entry.  Launched by tests/test_view_parallel_gpu.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/tools/vp_check.py [--transports nccl,multimem] [--P 50003]

P is deliberately NOT a multiple of 4 (the flat buffer pads its blocks; the kernels' 16-byte reductions must
cope with any P).  Prints one JSON line on rank 0; exit code 1 on mismatch."""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--transports", default="nccl,multimem,multimem_red")
    ap.add_argument("--P", type=int, default=50003)
    ap.add_argument("--views", type=int, default=3, help="views per rank")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=device)
    import g4splat_b200.diff_surfel_rasterization as op
    from g4splat_b200 import synthetic as S
    from g4splat_b200.view_parallel import ViewShardedGradSync, multimem_available

    W, H = 640, 360
    sc = S.make_scene(args.P, 17)
    t = lambda a: torch.from_numpy(a).to(device).requires_grad_(True)
    params = {"xyz": t(sc["means3D"]), "features": t(sc["shs"]), "opacity": t(sc["opacities"]),
              "scaling": t(sc["scales"]), "rotation": t(sc["rotations"])}
    cams = S.make_cameras(world * args.views, W, H)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
    gc, go = S.make_upstream_grads(W, H, 5)
    g_color, g_allmap = d(gc), d(go)
    bg = torch.zeros(3, device=device)

    def render(sync, k):
        c = cams[k]
        rs = op.GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=bg,
                                              scale_modifier=1.0, viewmatrix=d(c.viewmatrix), projmatrix=d(c.projmatrix),
                                              sh_degree=3, campos=d(c.campos), prefiltered=False, debug=False)
        means2D = torch.zeros_like(params["xyz"], requires_grad=True)
        color, radii, allmap = op.GaussianRasterizer(rs)(means3D=params["xyz"], means2D=means2D, opacities=params["opacity"],
                                                         shs=params["features"], scales=params["scaling"],
                                                         rotations=params["rotation"])
        torch.autograd.backward([color, allmap], [g_color, g_allmap])
        sync.add_view_stats(means2D.grad, radii)

    def snapshot(sync):
        out = {k: v.detach().clone() for k, v in sync._views.items()}
        out.update(accum=sync._accum.clone(), denom=sync._denom.clone(), max_radii=sync.max_radii.clone())
        return out

    # the single-rank answer: every rank loops over ALL views with a local accumulator
    local_sync = ViewShardedGradSync(params, transport="nccl")
    local_sync.bind(op)
    for k in range(world * args.views):
        render(local_sync, k)
    torch.cuda.synchronize()
    want = snapshot(local_sync)

    report, ok = {"world": world, "P": args.P, "multimem_available": bool(multimem_available())}, True
    for transport in args.transports.split(","):
        if transport != "nccl" and not multimem_available():
            report[transport] = "skipped: no multicast support on this box"
            continue
        sync = ViewShardedGradSync(params, transport=transport)
        sync.bind(op)
        for step in range(2):       # two steps: zero() + barrier between them must leave no residue
            sync.zero()
            for k in range(rank * args.views, (rank + 1) * args.views):
                render(sync, k)
            sync.allreduce()
        torch.cuda.synchronize()
        got = snapshot(sync)
        per = {}
        for k in want:
            w_, g_ = want[k].double(), got[k].double()
            per[k] = float((w_ - g_).abs().max() / w_.abs().max().clamp_min(1e-30))
        worst = torch.tensor([max(per.values())], device=device)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        report[transport] = {"max_rel_err": float(worst.item()), "per_block": per}
        ok = ok and float(worst.item()) <= 1e-4
        sync.close()
    local_sync.close()
    if rank == 0:
        report["ok"] = ok
        sys.stdout.write("\n" + json.dumps(report) + "\n")
        sys.stdout.flush()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
