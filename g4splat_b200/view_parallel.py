"""View-sharded data parallelism for the surfel rasterizer.

The reference pipeline is single-GPU (no torch.distributed call on any executed path); rendering
one camera is independent of every other camera, and the only coupling between views is that
they differentiate the same Gaussian parameters.  So the path shards over VIEWS:

  * every rank holds a full replica of the Gaussian parameters;
  * a step's batch of views is split into contiguous blocks, one per rank (`shard_views`);
  * each rank renders its views forward + backward with the local rasterizer, torch autograd
    accumulating the parameter gradients over those views;
  * ONE all-reduce(SUM) of a single flat fp32 buffer carries every parameter gradient plus the
    two densification statistics the trainer derives from the operator's outputs
    (`xyz_gradient_accum`, `denom`: 2DGS/scene/gaussian_model.py:649-651), and one
    all-reduce(MAX) carries `max_radii2D` (train_with_refine_depth.py:583).

After `allreduce()` every rank holds gradients identical (up to fp32 summation order) to a
single-GPU loop over all the views, which is what tests/test_view_parallel.py checks.
No collective sits on the per-view data path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(num_views: int, world_size: int, rank: int) -> range:
    """Contiguous block partition: the first (num_views % world_size) ranks get one extra view
    (50 views on 4 ranks -> 13, 13, 12, 12)."""
    base, extra = divmod(num_views, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class ViewShardedGradSync:
    """Packs parameter grads + densification statistics into one flat buffer and all-reduces it.

    params: name -> leaf tensor with requires_grad (e.g. xyz [P,3], features_dc [P,1,3],
            features_rest [P,15,3], opacity [P,1], scaling [P,2], rotation [P,4]: 58 floats per
            Gaussian); with the two statistics the buffer is [P, 60] fp32 = 240 B per Gaussian.
    """

    def __init__(self, params: Dict[str, torch.Tensor], group: Optional[dist.ProcessGroup] = None):
        self.params = params
        self.group = group
        first = next(iter(params.values()))
        self.P = int(first.shape[0])
        self.device = first.device
        self.sizes = {k: int(v.numel() // max(self.P, 1)) for k, v in params.items()}
        self.width = sum(self.sizes.values()) + 2
        self.flat = torch.zeros((self.P, self.width), dtype=torch.float32, device=self.device)
        self.max_radii = torch.zeros((self.P,), dtype=torch.int32, device=self.device)
        self._stats = self.flat[:, -2:]
        self._handles: List = []

    # -- per view ---------------------------------------------------------------------------
    @torch.no_grad()
    def add_view_stats(self, viewspace_grad: torch.Tensor, radii: torch.Tensor) -> None:
        """What the trainer does after every backward (train_with_refine_depth.py:582-593,
        gaussian_model.py:649-651): accumulate |dL_dmean2D[:, :2]| and the visibility count of
        visible Gaussians, and the running max of the screen-space radius."""
        vis = radii > 0
        self._stats[:, 0] += torch.where(vis, viewspace_grad[:, :2].norm(dim=-1), torch.zeros((), device=self.device))
        self._stats[:, 1] += vis.to(torch.float32)
        torch.maximum(self.max_radii, radii.to(torch.int32), out=self.max_radii)

    def zero(self) -> None:
        self.flat.zero_()
        self.max_radii.zero_()
        for p in self.params.values():
            p.grad = None

    # -- per step ---------------------------------------------------------------------------
    @torch.no_grad()
    def pack(self) -> None:
        col = 0
        for k, p in self.params.items():
            n = self.sizes[k]
            if p.grad is not None:
                self.flat[:, col:col + n] = p.grad.reshape(self.P, n)
            else:
                self.flat[:, col:col + n] = 0
            col += n

    @torch.no_grad()
    def unpack(self) -> None:
        col = 0
        for k, p in self.params.items():
            n = self.sizes[k]
            g = self.flat[:, col:col + n].reshape(p.shape)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            col += n

    def allreduce(self, async_op: bool = False):
        """SUM over ranks of every gradient + statistic, MAX of the radii.  With async_op the
        collectives run on the process group's stream; call `wait()` before reading grads."""
        self.pack()
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            self.unpack()
            return None
        h1 = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        h2 = dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=self.group, async_op=async_op)
        if async_op:
            self._handles = [h1, h2]
            return self._handles
        self.unpack()
        return None

    def wait(self) -> None:
        for h in self._handles:
            h.wait()
        self._handles = []
        self.unpack()

    @property
    def xyz_gradient_accum(self) -> torch.Tensor:
        return self._stats[:, 0:1]

    @property
    def denom(self) -> torch.Tensor:
        return self._stats[:, 1:2]

    @property
    def bytes_per_step(self) -> int:
        return self.flat.numel() * 4 + self.max_radii.numel() * 4


def render_views_sharded(render_one, views: Sequence, sync: ViewShardedGradSync, rank: int, world_size: int,
                         async_allreduce: bool = False):
    """Runs `render_one(view) -> (loss, viewspace_points, radii)` for this rank's block of
    `views`, back-propagates each loss (grads accumulate in the leaves), records the
    densification statistics and all-reduces once.  Returns the local per-view losses."""
    losses = []
    for i in shard_views(len(views), world_size, rank):
        loss, viewspace_points, radii = render_one(views[i])
        loss.backward()
        sync.add_view_stats(viewspace_points.grad, radii)
        losses.append(loss.detach())
    sync.allreduce(async_op=async_allreduce)
    return losses
