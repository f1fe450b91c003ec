"""View-sharded data parallelism for the surfel rasterizer.

The reference pipeline is single-GPU (no torch.distributed call on any executed path); rendering
one camera is independent of every other camera, and the only coupling between views is that
they differentiate the same Gaussian parameters.  So the path shards over VIEWS:

  * every rank holds a full replica of the Gaussian parameters;
  * a step's batch of views is split into contiguous blocks, one per rank (`shard_views`);
  * each rank renders its views forward + backward with the local rasterizer; autograd
    accumulates the parameter gradients IN PLACE into one flat fp32 buffer (every `p.grad` is a
    view of it), which also holds the two densification statistics the trainer derives from
    the operator's outputs (`xyz_gradient_accum`, `denom`: 2DGS/scene/gaussian_model.py:649-651);
  * ONE all-reduce(SUM) of that buffer (58 gradient floats + 2 statistics = 240 B per
    Gaussian) and one all-reduce(MAX) of `max_radii2D` (train_with_refine_depth.py:583) per step.

After `allreduce()` every rank holds gradients identical (up to fp32 summation order) to a
single-GPU loop over all the views, which is what tests/test_view_parallel.py checks.
No collective sits on the per-view data path, and nothing is packed or copied for the
collective: the buffer the kernels accumulate into is the buffer NCCL reduces.
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
    """One flat gradient + statistics buffer shared by autograd and the collective.

    params: name -> leaf tensor with requires_grad, first dimension P (e.g. xyz [P,3],
            features [P,16,3], opacity [P,1], scaling [P,2], rotation [P,4]: 58 floats per
            Gaussian; with the two statistics the buffer holds 60 floats = 240 B per Gaussian).
    """

    def __init__(self, params: Dict[str, torch.Tensor], group: Optional[dist.ProcessGroup] = None):
        self.params = params
        self.group = group
        first = next(iter(params.values()))
        self.P = int(first.shape[0])
        self.device = first.device
        self.sizes = {k: int(v.numel() // max(self.P, 1)) for k, v in params.items()}
        self.width = sum(self.sizes.values()) + 2
        self.flat = torch.zeros((self.P * self.width,), dtype=torch.float32, device=self.device)
        self.max_radii = torch.zeros((self.P,), dtype=torch.int32, device=self.device)
        self._views: Dict[str, torch.Tensor] = {}
        off = 0
        for k, p in params.items():
            n = p.numel()
            self._views[k] = self.flat[off:off + n].view(p.shape)
            off += n
        self._accum = self.flat[off:off + self.P]
        self._denom = self.flat[off + self.P:off + 2 * self.P]
        self._handles: List = []
        self._lib = None
        if self.device.type == "cuda":
            from . import _lib
            self._lib = _lib
        self.attach()

    def bind(self, op_module, names=None) -> None:
        """Let the B200 operator add gradients straight into the flat buffer (kernel-side, visible
        rows only) for parameters that are passed to it as they are; everything else keeps
        flowing through autograd into the same buffer.  `op_module` must offer set_gradient_sink
        (g4splat_b200.diff_surfel_rasterization does; the reference extension does not)."""
        if not hasattr(op_module, "set_gradient_sink"):
            return
        names = list(self.params) if names is None else names
        op_module.set_gradient_sink({self.params[k]: self._views[k] for k in names})

    def attach(self) -> None:
        """Point every p.grad at its block of the flat buffer so that backward accumulates in place."""
        for k, p in self.params.items():
            p.grad = self._views[k]

    # -- per view ---------------------------------------------------------------------------
    @torch.no_grad()
    def add_view_stats(self, viewspace_grad: torch.Tensor, radii: torch.Tensor) -> None:
        """What the trainer does after every backward (train_with_refine_depth.py:582-593,
        gaussian_model.py:649-651): accumulate |dL_dmean2D[:, :2]| and the visibility count of
        visible Gaussians, and the running max of the screen-space radius."""
        if self._lib is not None and viewspace_grad.is_cuda:
            g = viewspace_grad.contiguous()
            r = radii.contiguous()
            lib = self._lib.load()
            with torch.cuda.device(self.device):
                self._lib.check(lib.g4s_densify_stats(self.P, g.data_ptr(), r.data_ptr(), self._accum.data_ptr(),
                                                      self._denom.data_ptr(), self.max_radii.data_ptr(),
                                                      torch.cuda.current_stream(self.device).cuda_stream))
            return
        vis = radii > 0
        self._accum += torch.where(vis, viewspace_grad[:, :2].norm(dim=-1), torch.zeros((), device=self.device))
        self._denom += vis.to(torch.float32)
        torch.maximum(self.max_radii, radii.to(torch.int32), out=self.max_radii)

    def zero(self) -> None:
        self.flat.zero_()
        self.max_radii.zero_()
        self.attach()

    # -- per step ---------------------------------------------------------------------------
    def allreduce(self, async_op: bool = False):
        """SUM over ranks of every gradient + statistic, MAX of the radii.  With async_op the
        collectives run on the process group's stream; call `wait()` before reading grads."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return None
        h1 = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        h2 = dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=self.group, async_op=async_op)
        if async_op:
            self._handles = [h1, h2]
            return self._handles
        return None

    def wait(self) -> None:
        for h in self._handles:
            h.wait()
        self._handles = []

    @property
    def xyz_gradient_accum(self) -> torch.Tensor:
        return self._accum.view(self.P, 1)

    @property
    def denom(self) -> torch.Tensor:
        return self._denom.view(self.P, 1)

    @property
    def bytes_per_step(self) -> int:
        return self.flat.numel() * 4 + self.max_radii.numel() * 4


def render_views_sharded(render_one, views: Sequence, sync: ViewShardedGradSync, rank: int, world_size: int,
                         async_allreduce: bool = False):
    """Runs `render_one(view) -> (loss, viewspace_points, radii)` for this rank's block of
    `views`, back-propagates each loss (grads accumulate in the flat buffer), records the
    densification statistics and all-reduces once.  Returns the local per-view losses."""
    losses = []
    for i in shard_views(len(views), world_size, rank):
        loss, viewspace_points, radii = render_one(views[i])
        loss.backward()
        sync.add_view_stats(viewspace_points.grad, radii)
        losses.append(loss.detach())
    sync.allreduce(async_op=async_allreduce)
    return losses
