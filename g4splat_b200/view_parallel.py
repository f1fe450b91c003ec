"""View-sharded data parallelism for the surfel rasterizer.

The reference pipeline is single-GPU (no torch.distributed call on any executed path); rendering
one camera is independent of every other camera, and the only coupling between views is that
they differentiate the same Gaussian parameters.  So the path shards over VIEWS:

  * every rank holds a full replica of the Gaussian parameters;
  * a step's batch of views is split into contiguous blocks, one per rank (`shard_views`);
  * each rank renders its views forward + backward with the local rasterizer; the parameter
    gradients accumulate IN PLACE into one flat fp32 buffer (every `p.grad` is a view of it), which
    also holds the two densification statistics the trainer derives from the operator's outputs
    (`xyz_gradient_accum`, `denom`: 2DGS/scene/gaussian_model.py:649-651) and `max_radii2D`
    (train_with_refine_depth.py:583).

Three transports bring the ranks' sums together:

  "nccl"          ONE all-reduce(SUM) of the flat buffer (58 gradient floats + 2 statistics = 240 B per
                  Gaussian) and one all-reduce(MAX) of the radii per step, issued on separate process
                  groups so that they run concurrently.
  "multimem"      the flat buffer lives in symmetric memory that every rank maps at the same offset of one
                  NVSwitch multicast object; each rank accumulates its own views locally (kernel-side sink),
                  and the step ends with ONE hand-written kernel (csrc/project.cu multimem_allreduce_kernel):
                  rank r pulls the r-th slice from all ranks with multimem.ld_reduce (the switch adds) and
                  pushes the sum into every replica with multimem.st, bracketed by two symmetric-memory
                  barriers.  Every byte crosses every link once per direction.
  "multimem_red"  no step-end pass at all: the gradient sink of the B200 operator is given the MULTICAST
                  address and `project_bwd` / `densify_stats` add each view's rows with `multimem.red` -- the
                  switch adds them into every rank's replica while the next kernels run; the per-step
                  "all-reduce" is a barrier.  The fused producer + collective of SURVEY.md 8e.  Measured slower
                  than "multimem" (16-byte reductions are a poor NVLink packet size); needs every
                  parameter bound to the operator.

After `allreduce()` every rank holds gradients identical (up to fp32 summation order) to a
single-GPU loop over all the views, which is what tests/test_view_parallel.py checks.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist

_ALIGN = 32   # floats: every block of the flat buffer starts on a 128-byte boundary


def shard_views(num_views: int, world_size: int, rank: int, strided: bool = False) -> range:
    """The views of a step that `rank` renders; the first (num_views % world_size) ranks get one extra view
    (50 views on 4 ranks -> 13, 13, 12, 12).  Default: contiguous blocks.  strided=True deals the views out like
    cards (rank, rank + world_size, ...): neighbouring cameras of a capture cost about the same, so a contiguous
    block gives a rank a correlated -- all cheap or all expensive -- set and the step waits for the unlucky rank;
    dealing them out evens the per-rank sums (measured at 8 GPUs in DESIGN.md 5)."""
    if strided:
        return range(rank, num_views, world_size)
    base, extra = divmod(num_views, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _round_up(n: int, a: int) -> int:
    return (n + a - 1) // a * a


def multimem_available(group: Optional[dist.ProcessGroup] = None) -> bool:
    """True when torch symmetric memory can give this group a multicast mapping (NVSwitch + driver support)."""
    if not (dist.is_available() and dist.is_initialized() and torch.cuda.is_available()):
        return False
    if dist.get_world_size(group) < 2 or dist.get_backend(group) != "nccl":
        return False
    try:
        from torch._C._autograd import DeviceType
        from torch._C._distributed_c10d import _SymmetricMemory
        return bool(_SymmetricMemory.has_multicast_support(DeviceType.CUDA, torch.cuda.current_device()))
    except Exception:  # noqa: BLE001
        return False


class ViewShardedGradSync:
    """One flat gradient + statistics buffer shared by autograd, the kernels and the transport.

    params: name -> leaf tensor with requires_grad, first dimension P (e.g. xyz [P,3],
            features [P,16,3], opacity [P,1], scaling [P,2], rotation [P,4]: 58 floats per
            Gaussian; with the two statistics the buffer holds 60 floats = 240 B per Gaussian).
    transport: "nccl" | "multimem" | "multimem_red" | "auto" (multimem when the hardware offers it, else nccl;
            G4S_TRANSPORT overrides "auto").
    use_native: False keeps every step of this class in plain torch (used by the benchmark's
            reference arm, which must not load this repository's kernels).
    """

    def __init__(self, params: Dict[str, torch.Tensor], group: Optional[dist.ProcessGroup] = None,
                 transport: str = "auto", use_native: bool = True):
        self.group = group
        self.use_native = use_native
        self._lib = None
        self._handles: List = []
        self._max_group = None
        self._symm = None
        self._bound_module = None
        self._bound_names = None
        requested = transport
        if transport == "auto":
            transport = os.environ.get("G4S_TRANSPORT", "auto")
            requested = transport
        if transport == "auto":
            transport = "multimem" if (use_native and multimem_available(group)) else "nccl"
        if transport not in ("nccl", "multimem", "multimem_red"):
            raise ValueError(f"unknown transport {transport!r}")
        self.transport = transport
        try:
            self._build(params)
        except Exception as ex:  # noqa: BLE001
            if requested != "auto" or transport == "nccl":
                raise
            # "auto" promised a working transport: the multicast mapping was refused (no NVLS fabric in this container,
            # a driver without multicast objects, ...).  Every rank takes the same decision.
            import warnings
            warnings.warn(f"view_parallel: multicast transport unavailable ({ex}); falling back to NCCL")
            self.transport = "nccl"
            self._symm = None
            self._build(params)

    # -- layout ---------------------------------------------------------------------------------
    def _build(self, params: Dict[str, torch.Tensor]) -> None:
        self.params = params
        first = next(iter(params.values()))
        self.P = int(first.shape[0])
        self.device = first.device
        self.sizes = {k: int(v.numel() // max(self.P, 1)) for k, v in params.items()}
        self.width = sum(self.sizes.values()) + 2
        # blocks padded to 128 bytes: the kernels' 16-byte reductions need aligned rows whatever P is
        self._offsets, off = {}, 0
        for k, p in params.items():
            self._offsets[k] = off
            off = _round_up(off + p.numel(), _ALIGN)
        self._offsets["accum"] = off
        off = _round_up(off + self.P, _ALIGN)
        self._offsets["denom"] = off
        off = _round_up(off + self.P, _ALIGN)
        self._grad_floats = off                      # what the SUM covers
        self._offsets["max_radii"] = off
        total = _round_up(off + self.P, _ALIGN)
        self._mc_base = None
        if self.transport != "nccl":
            import torch.distributed._symmetric_memory as symm_mem
            g = self.group if self.group is not None else dist.group.WORLD
            store = symm_mem.empty(max(total, _ALIGN), dtype=torch.float32, device=self.device)
            self._symm = symm_mem.rendezvous(store, g)
            if not self._symm.multicast_ptr:
                raise RuntimeError(f"transport {self.transport!r} needs a multicast mapping (NVLS); none was granted")
            self._mc_base = int(self._symm.multicast_ptr)
            store.zero_()
            self._store = store
        else:
            self._store = torch.zeros((max(total, 1),), dtype=torch.float32, device=self.device)
        self.flat = self._store[:self._grad_floats]
        self.max_radii = self._store[self._offsets["max_radii"]:self._offsets["max_radii"] + self.P].view(torch.int32)
        self._views: Dict[str, torch.Tensor] = {
            k: self._store[self._offsets[k]:self._offsets[k] + p.numel()].view(p.shape) for k, p in params.items()}
        self._accum = self._store[self._offsets["accum"]:self._offsets["accum"] + self.P]
        self._denom = self._store[self._offsets["denom"]:self._offsets["denom"] + self.P]
        if self.device.type == "cuda" and self.use_native:
            from . import _lib
            self._lib = _lib
        self.attach()
        if self.transport != "nccl":
            self._barrier()   # every replica is zero before anybody reduces into it

    def _mc(self, name: str) -> int:
        return self._mc_base + 4 * self._offsets[name]

    def rebind(self, params: Dict[str, torch.Tensor]) -> None:
        """Call after densification / pruning replaced the parameter tensors (P may have changed): lays the
        buffer out again for the new tensors and re-registers the operator's gradient sink.  Collective: every
        rank must call it with the same P (densification is replica-identical, see gaussian_model.densify)."""
        self._build(params)
        if self._bound_module is not None:
            self.bind(self._bound_module, self._bound_names)

    def bind(self, op_module, names=None) -> None:
        """Let the B200 operator add gradients straight into the flat buffer (kernel-side, visible
        rows only) for parameters that are passed to it as they are; everything else keeps
        flowing through autograd into the same buffer.  `op_module` must offer set_gradient_sink
        (g4splat_b200.diff_surfel_rasterization does; the reference extension does not).  Must be called
        again (or `rebind`) whenever the parameter tensors are replaced."""
        if not hasattr(op_module, "set_gradient_sink"):
            if self.transport == "multimem_red":
                raise RuntimeError("transport 'multimem_red' needs an operator with a gradient sink")
            return
        self._bound_module, self._bound_names = op_module, names
        names = list(self.params) if names is None else list(names)
        mapping = {self.params[k]: self._views[k] for k in names}
        if self.transport == "multimem_red":
            if set(names) != set(self.params):
                raise RuntimeError("transport 'multimem_red': every parameter must be bound to the operator "
                                   "(gradients that arrive through autograd would stay rank-local)")
            op_module.set_gradient_sink(mapping, multicast={self.params[k]: self._mc(k) for k in names})
        else:
            op_module.set_gradient_sink(mapping)

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
                sp = torch.cuda.current_stream(self.device).cuda_stream
                if self.transport == "multimem_red":
                    self._lib.check(lib.g4s_densify_stats_multimem(self.P, g.data_ptr(), r.data_ptr(), self._mc("accum"),
                                                                   self._mc("denom"), self._mc("max_radii"), sp))
                else:
                    self._lib.check(lib.g4s_densify_stats(self.P, g.data_ptr(), r.data_ptr(), self._accum.data_ptr(),
                                                          self._denom.data_ptr(), self.max_radii.data_ptr(), sp))
            return
        vis = radii > 0
        self._accum += torch.where(vis, viewspace_grad[:, :2].norm(dim=-1), torch.zeros((), device=self.device))
        self._denom += vis.to(torch.float32)
        torch.maximum(self.max_radii, radii.to(torch.int32), out=self.max_radii)

    def zero(self) -> None:
        self._store.zero_()
        self.attach()
        if self.transport == "multimem_red":
            self._barrier()   # nobody reduces into a replica that is still being cleared

    # -- per step ---------------------------------------------------------------------------
    def _barrier(self) -> None:
        """Device-side barrier on the current stream (symmetric-memory signal pads, release / acquire at system
        scope): the reductions every rank issued before it are visible in every replica after it."""
        self._symm.barrier(channel=0)

    def _max_pg(self):
        if self._max_group is None:
            # a second communicator: the MAX of the radii runs beside the SUM instead of behind it
            ranks = list(range(dist.get_world_size(self.group))) if self.group is None else dist.get_process_group_ranks(self.group)
            self._max_group = dist.new_group(ranks=ranks, backend=dist.get_backend(self.group))
        return self._max_group

    def allreduce(self, async_op: bool = False):
        """After this call (and `wait()` when async) every rank's buffer holds the SUM over ranks of every
        gradient + statistic and the MAX of the radii."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return None
        if self.transport == "multimem_red":
            self._barrier()
            return None
        if self.transport == "multimem":
            lib = self._lib.load()
            world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
            self._barrier()       # every rank's partial sums are complete
            with torch.cuda.device(self.device):
                self._lib.check(lib.g4s_multimem_allreduce(self._mc_base, self._grad_floats, self._mc("max_radii"), self.P, rank, world,
                                                           torch.cuda.current_stream(self.device).cuda_stream))
            self._barrier()       # every slice has been written to every replica
            return None
        concurrent = dist.get_backend(self.group) == "nccl"
        h2 = dist.all_reduce(self.max_radii, op=dist.ReduceOp.MAX, group=self._max_pg() if concurrent else self.group,
                             async_op=async_op or concurrent)
        h1 = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        if async_op:
            self._handles = [h1, h2]
            return self._handles
        if concurrent:
            h2.wait()
        return None

    def close(self) -> None:
        """Drop the symmetric-memory mapping (call before destroying the process group)."""
        if self._bound_module is not None and hasattr(self._bound_module, "set_gradient_sink"):
            self._bound_module.set_gradient_sink(None)
        for p in self.params.values():
            p.grad = None
        self._views, self._accum, self._denom, self.flat, self.max_radii = {}, None, None, None, None
        self._store, self._symm = None, None

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
        """Bytes the transport has to combine per step (the flat gradient block and the radii)."""
        return self._grad_floats * 4 + self.max_radii.numel() * 4


def render_views_sharded(render_one, views: Sequence, sync: ViewShardedGradSync, rank: int, world_size: int,
                         async_allreduce: bool = False, strided: bool = False):
    """Runs `render_one(view) -> (loss, viewspace_points, radii)` for this rank's share of
    `views` (`shard_views`), back-propagates each loss (grads accumulate in the flat buffer), records the
    densification statistics and combines the ranks' sums once.  Returns the local per-view losses."""
    losses = []
    for i in shard_views(len(views), world_size, rank, strided=strided):
        loss, viewspace_points, radii = render_one(views[i])
        loss.backward()
        sync.add_view_stats(viewspace_points.grad, radii)
        losses.append(loss.detach())
    sync.allreduce(async_op=async_allreduce)
    return losses
