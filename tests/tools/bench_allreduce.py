"""Times the per-step gradient combine of g4splat_b200.view_parallel alone (no rendering): NCCL all-reduce vs the
NVLS multimem all-reduce kernel, on the flat buffer of a P-Gaussian model.  Under torchrun:

    python -m torch.distributed.run --nproc-per-node N ... tests/tools/bench_allreduce.py [--P 1000000] [--iters 20]
"""
import argparse
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--P", type=int, default=1_000_000)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=device)
    from g4splat_b200.view_parallel import ViewShardedGradSync, multimem_available
    P = args.P
    params = {"xyz": torch.zeros(P, 3, device=device, requires_grad=True), "features": torch.zeros(P, 16, 3, device=device, requires_grad=True),
              "opacity": torch.zeros(P, 1, device=device, requires_grad=True), "scaling": torch.zeros(P, 2, device=device, requires_grad=True),
              "rotation": torch.zeros(P, 4, device=device, requires_grad=True)}
    out = {"world": world, "P": P}
    for transport in ("nccl", "multimem"):
        if transport == "multimem" and not multimem_available():
            continue
        sync = ViewShardedGradSync(params, transport=transport)
        sync._store.fill_(1.0)
        for _ in range(3):
            sync.allreduce()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            sync.allreduce()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / args.iters], device=device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        nbytes = sync.bytes_per_step
        out[transport] = {"ms": float(ms.item()), "bytes": nbytes,
                          "busbw_GBps": 2 * (world - 1) / world * nbytes / (float(ms.item()) * 1e-3) / 1e9}
        sync.close()
    if rank == 0:
        sys.stdout.write("\n" + json.dumps(out) + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
