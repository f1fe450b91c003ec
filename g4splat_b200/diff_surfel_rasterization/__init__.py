"""Drop-in operator API of the B200-native surfel rasterizer.

Mirrors the reference operator module one to one
(RAST/diff_surfel_rasterization/__init__.py: rasterize_gaussians :21-42, _RasterizeGaussians
:44-156, GaussianRasterizationSettings :158-170, GaussianRasterizer :172-222): same names, same
argument order and meaning, same return order `(color, radii, allmap)`, same two one-of
exceptions -- so `gaussian_renderer.render()` (2DGS/gaussian_renderer/__init__.py:14,37-53,
97-106) runs unchanged once this package is importable as `diff_surfel_rasterization`
(`g4splat_b200.install()` registers it under that name).

Below the API everything is different: torch only allocates tensors and provides the stream;
the work is done by hand-written sm_100a kernels reached through the C ABI of
include/g4s_rasterizer.h via ctypes.  There is no CPU / PyTorch fallback.

Host synchronisation: the reference blocks on a device->host copy of `num_rendered` in the
middle of every forward (CR/rasterizer_impl.cu:282).  Here the forward is launched end to end
with a guessed instance capacity; the host then waits only for the small "plan" stage to learn
the true count (the GPU keeps running the rest meanwhile) and re-issues the render stage in the
rare case the guess was too small.  G4S_SYNC=none skips even that wait (capacity overflow is
then reported by the next call).
"""
from __future__ import annotations

import collections.abc
import contextlib
import os
import threading
import weakref
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from .. import _lib

_LIB = _lib.load()  # raises if the CUDA library is missing: fail loudly, never fall back

NUM_CHANNELS = 3  # CR/config.h:14
_OTHERS = 7       # depth, alpha, normal xyz, median depth, distortion (CR/auxiliary.h:23-27)


def _ptr(t: Optional[torch.Tensor]):
    """Device pointer, or NULL for None / empty tensors (the reference's optional-arg convention)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.numel() and t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


class _CapacityPolicy:
    """Capacity (number of (Gaussian, tile) instances) to allocate for the next forward, per device.

    Sticky and bucketed: 1.5x the largest count seen recently, rounded up to {1, 1.25, 1.5, 1.75} x 2^k,
    so that consecutive calls ask the caching allocator for the same block size instead of a new size
    per view (which fragments the pool and ends in synchronous cudaMalloc calls)."""

    def __init__(self):
        self.cap = {}

    @staticmethod
    def bucket(n: int) -> int:
        n = max(int(n), 1 << 16)
        k = n.bit_length() - 1
        for q in (4, 5, 6, 7, 8):
            c = (q << k) >> 2
            if c >= n:
                return c
        return 1 << (k + 1)

    def guess(self, dev: int, P: int) -> int:
        cap = self.cap.get(dev)
        if cap is None:
            return self.bucket(max(1 << 20, 6 * P))
        return cap

    def observe(self, dev: int, R: int) -> None:
        want = self.bucket(int(R * 1.5) + 65536)
        cur = self.cap.get(dev)
        if cur is None or want > cur or want * 4 < cur:
            self.cap[dev] = want


_capacity = _CapacityPolicy()   # keyed by device index inside
# G4S_HOST_TRACE=1: wall-clock of the host-side segments of every call (diagnostics; host_trace_summary())
_HOST_TRACE = os.environ.get("G4S_HOST_TRACE") == "1"
_host_times = {"fwd_launch": [], "fwd_wait": [], "bwd": []}


def host_trace_summary() -> dict:
    import statistics
    return {k: dict(n=len(v), mean_us=1e6 * statistics.fmean(v), p50_us=1e6 * statistics.median(v), max_us=1e6 * max(v))
            for k, v in _host_times.items() if v}


_RING = 256             # pinned count slots per device; a slot is reused only after _RING further forwards


class _ScratchSet:
    """One set of forward scratch buffers (radii, geometry, image, binning) owned by the pipelined path of view_batch().
    The torch caching allocator cannot serve that path: buffers written on the side stream and consumed on the main one
    need `record_stream`, which defers their reuse until the main stream has caught up, so with the host a view or two
    ahead every forward ends in a fresh cudaMalloc (measured: 0.8 ms of host time per view).  The sets are allocated once,
    grown when a view needs more, and recycled explicitly: `released` is recorded on the consuming stream by the backward
    of the view that used the set (or when its autograd context dies), and the side stream waits for it before reuse."""
    __slots__ = ("radii", "geom", "img", "binning", "busy", "released", "generation")

    def __init__(self):
        self.radii = self.geom = self.img = self.binning = None
        self.busy = False
        self.released = None
        self.generation = 0


class _Lease:
    """Held by the autograd context of a pipelined forward: gives the scratch set back when the backward has been
    enqueued, or when the context is dropped without one."""

    def __init__(self, sset, dev):
        self.sset, self.dev, self.generation, self.done = sset, dev, sset.generation, False

    def check(self):
        if self.done or self.generation != self.sset.generation:
            raise RuntimeError("g4s rasterizer: the scratch buffers of this view_batch() forward have been recycled "
                               "(a second backward through the same graph is not supported inside view_batch)")

    def release(self):
        if not self.done:
            self.done = True
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev))
            self.sset.released = ev
            self.sset.busy = False

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


class _DeviceState:
    """Everything the shim remembers between calls, per CUDA device (one process may drive several devices
    from several threads): the pinned ring the plan stage reports its counts into, the overflow checks a
    G4S_SYNC=none caller still owes, and the counts of the most recent forward the host has waited for."""

    def __init__(self):
        self.lock = threading.Lock()
        self.ring = None
        self.ring_next = 0
        self.pending_overflow = []   # (event, pinned counts, capacity) of G4S_SYNC=none calls not yet checked
        self.last_counts = {"num_rendered": 0, "max_tile_list": 0, "visible": 0}
        self.batch = None            # _ViewBatch while a view_batch() block is open on this device
        self.pipe_sets = [_ScratchSet() for _ in range(3)]   # persistent scratch of the pipelined path, used round-robin
        self.pipe_next = 0


_states = {}
_states_lock = threading.Lock()


def _state(dev: torch.device) -> _DeviceState:
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    st = _states.get(idx)
    if st is None:
        with _states_lock:
            st = _states.setdefault(idx, _DeviceState())
    return st


class _LastCounts(collections.abc.Mapping):
    """(num_rendered, longest tile list, visible Gaussians) of the most recent forward on the CURRENT device
    whose plan the host has waited for; benchmarks read it to size the algorithmic-bytes model."""

    def _d(self):
        return _state(torch.device("cuda", torch.cuda.current_device())).last_counts

    def __getitem__(self, k):
        return self._d()[k]

    def __iter__(self):
        return iter(self._d())

    def __len__(self):
        return len(self._d())

    def __repr__(self):
        return repr(self._d())


last_counts = _LastCounts()


def _sync_mode() -> str:
    return os.environ.get("G4S_SYNC", "plan")


def _pinned_counts(st: _DeviceState) -> torch.Tensor:
    """int32[4] view into the device's pinned ring (cudaHostAlloc per call would serialise the device)."""
    with st.lock:
        if st.ring is None:
            st.ring = torch.zeros((_RING, 4), dtype=torch.int32).pin_memory()
        slot = st.ring[st.ring_next]
        st.ring_next = (st.ring_next + 1) % _RING
        wrapped = st.ring_next == 0
    if wrapped and st.pending_overflow:
        _check_pending(st, wait=True)     # the slots about to be reused still carry unchecked counts
    return slot


def _check_pending(st: _DeviceState, dev_index: int = None, wait: bool = False) -> None:
    """Capacity checks owed by G4S_SYNC=none forwards: raise if one of them overflowed its binning buffer
    (its kernels were no-ops), and feed the observed counts back into the capacity policy."""
    still = []
    for ev, counts, cap in st.pending_overflow:
        if wait:
            ev.synchronize()
        if ev.query():
            n = int(counts[0])
            if dev_index is not None:
                _capacity.observe(dev_index, n)
            st.last_counts.update(num_rendered=n, max_tile_list=int(counts[1]), visible=int(counts[2]))
            if n > cap:
                st.pending_overflow.clear()
                raise RuntimeError(
                    f"g4s rasterizer: a previous G4S_SYNC=none forward needed {n} instances "
                    f"but was given capacity {cap}; its outputs are invalid")
        else:
            still.append((ev, counts, cap))
    st.pending_overflow[:] = still


class _ViewBatch:
    def __init__(self, dev):
        self.side = torch.cuda.Stream(device=dev, priority=-1)     # high priority: its small kernels slot in as SMs free up
        self.start = torch.cuda.Event()
        self.start.record(torch.cuda.current_stream(dev))
        self.prefetched = 0


@contextlib.contextmanager
def view_batch(device=None):
    """Pipelines the views rendered inside the block (extension; the reference API is unchanged without it).

    Within one view the stages are strictly ordered, and everything before the blend -- projection, tile counts,
    scan, scatter, per-tile sort: six latency-bound kernels, ~0.2 ms of a 1.6 ms view at 1 M Gaussians / 1080p --
    leaves most of the GPU idle.  Inside this block the operator runs that FRONT END of a view on a high-priority
    side stream that does not wait for the work already queued on the current stream, so it overlaps the previous
    view's blend / backward kernels; the blend itself and the whole backward stay on the current stream.

    The caller's promise: the tensors passed to the operator inside the block -- parameters, camera matrices,
    background -- are complete on the device when the block is entered and are not modified inside it (a step's views
    share one parameter set: exactly the situation of `view_parallel.render_views_sharded`).  Differentiable inputs
    that are not leaves (they were computed inside the block, e.g. `exp(_scaling)`) switch a call back to the
    ordinary path; pass the raw leaves (`rasterize_gaussian_model`, `render(..., fused_activations=True)`) or
    activated leaves to benefit."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    st = _state(dev)
    if st.batch is not None:     # nested: the outer block already pipelines
        yield st.batch
        return
    st.batch = _ViewBatch(dev)
    try:
        yield st.batch
    finally:
        st.batch = None          # everything the side stream did has been waited for by the stream that consumed it


def set_fast_math(on: bool) -> bool:
    """Arithmetic of the forward blend (g4s_set_fast_math): False (default) = IEEE division / expf, results
    bit-identical to the reference; True = rcp.approx / ex2.approx, values within 1e-5, threshold decisions
    may flip within an ulp of alpha = 1/255, T = 1e-4, T = 0.5.  Returns the previous setting."""
    return bool(_LIB.g4s_set_fast_math(int(bool(on))))


if os.environ.get("G4S_MATH") == "fast":
    set_fast_math(True)


# ---- opt-in gradient sink (extension; the reference API is unchanged when it is not used) --------
# Maps a leaf parameter tensor to the buffer its gradient is accumulated in.  When an input of the
# operator IS such a parameter, backward adds that view's gradient straight into the buffer inside
# the kernel (visible rows only) and returns None for it, instead of returning a dense tensor that
# autograd then adds to `.grad` (view_parallel.ViewShardedGradSync.bind()).  Entries hold a weak
# reference to their parameter: a tensor that merely reuses the address of a replaced parameter
# (densification, optimizer re-creation) never matches, and dead entries are dropped.
_ACC_BITS = {"means3D": 1, "sh": 2, "opacities": 4, "scales": 8, "rotations": 16}
_ACC_MULTIMEM = 32


class _Sink:
    __slots__ = ("ref", "buf", "ptr", "multimem")

    def __init__(self, param, buf, ptr, multimem):
        self.ref, self.buf, self.ptr, self.multimem = weakref.ref(param), buf, ptr, multimem


_grad_sink = {}


def set_gradient_sink(mapping, multicast=None) -> None:
    """mapping: {parameter tensor: accumulation buffer of the same shape} (empty / None clears it).
    multicast: optional {parameter tensor: int} -- the NVSwitch multicast address that maps the parameter's
    accumulation buffer on every rank; the kernels then add with multimem.red (all ranks' buffers receive the
    gradient, include/g4s_rasterizer.h G4S_ACC_MULTIMEM).  Either every entry has one or none has."""
    _grad_sink.clear()
    multicast = multicast or {}
    items = list((mapping or {}).items())
    if multicast and len(multicast) != len(items):
        raise ValueError("set_gradient_sink: give a multicast address for every parameter or for none")
    for param, buf in items:
        if buf.shape != param.shape or buf.dtype != torch.float32 or not buf.is_contiguous():
            raise ValueError("gradient sink buffers must be contiguous float32 tensors shaped like their parameter")
        if buf.data_ptr() % 4 != 0:
            raise ValueError("gradient sink buffers must be 4-byte aligned")
        mc = multicast.get(param)
        if mc is not None and int(mc) % 16 != buf.data_ptr() % 16:
            raise ValueError("multicast address and local buffer must share their 16-byte alignment")
        _grad_sink[id(param)] = _Sink(param, buf, int(mc) if mc is not None else buf.data_ptr(), mc is not None)


def _sink_for(t: torch.Tensor):
    if not _grad_sink or t is None or t.numel() == 0 or not t.requires_grad or not t.is_leaf:
        return None
    s = _grad_sink.get(id(t))
    if s is None:
        return None
    if s.ref() is not t:      # the parameter died and its id was reused
        del _grad_sink[id(t)]
        return None
    return s


def _dump_snapshot(name: str, args) -> None:
    """debug=True: what the reference does when a call throws (RAST/diff_surfel_rasterization/__init__.py:83-90,
    133-140): the arguments, copied to the CPU before the call, are written next to the script."""
    torch.save(args, name)


def _cpu_copy(args):
    return tuple(a.cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


def _plan_and_render(dev, P, W, H, bg, rs, plan, prefetchable=False):
    """Forward of one view for P > 0: `plan(radii, geom, img, counts, stream_ptr)` issues g4s_forward_plan
    (or its raw-parameter twin), then the render stage is launched speculatively (module docstring).
    Returns (color, others, radii, geom, binning, img, capacity, num_rendered, pinned counts)."""
    debug = bool(rs.debug)
    f32 = dict(dtype=torch.float32, device=dev)
    st = _state(dev)
    dev_index = dev.index if dev.index is not None else torch.cuda.current_device()
    if _HOST_TRACE:
        import time
        t_begin = time.perf_counter()
    batch = st.batch if (prefetchable and not debug) else None
    if batch is not None:
        out = _plan_and_render_pipelined(dev, dev_index, st, batch, P, W, H, bg, rs, plan)
        if out is not None:
            return out
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        sp = stream.cuda_stream
        color = torch.empty((NUM_CHANNELS, H, W), **f32)
        others = torch.empty((_OTHERS, H, W), **f32)
        radii = torch.empty((P,), dtype=torch.int32, device=dev)
        geom = torch.empty((_LIB.g4s_geom_bytes(P),), dtype=torch.uint8, device=dev)
        img = torch.empty((_LIB.g4s_image_bytes(W, H),), dtype=torch.uint8, device=dev)
        counts = _pinned_counts(st)
        mode = _sync_mode()
        if mode == "none":
            _check_pending(st, dev_index)
        _lib.check(plan(radii, geom, img, counts, sp))
        planned = torch.cuda.Event()
        planned.record(stream)
        cap = _capacity.guess(dev_index, P)
        while True:
            binning = torch.empty((_LIB.g4s_binning_bytes(cap),), dtype=torch.uint8, device=dev)
            if debug:
                planned.synchronize()
                if int(counts[0]) > cap:
                    cap = int(counts[0])
                    continue
            _lib.check(_LIB.g4s_forward_render(
                P, W, H, bg.data_ptr(), geom.data_ptr(), img.data_ptr(), binning.data_ptr(), cap,
                color.data_ptr(), others.data_ptr(), sp, int(debug)))
            if mode == "none":
                st.pending_overflow.append((planned, counts, cap))
                num_rendered = -1
                break
            if _HOST_TRACE:
                t_wait = time.perf_counter()
            planned.synchronize()  # waits for project + scan only; the blend keeps running
            if _HOST_TRACE:
                t_done = time.perf_counter()
                _host_times["fwd_launch"].append(t_wait - t_begin)
                _host_times["fwd_wait"].append(t_done - t_wait)
            num_rendered = int(counts[0])
            st.last_counts.update(num_rendered=num_rendered, max_tile_list=int(counts[1]), visible=int(counts[2]))
            _capacity.observe(dev_index, num_rendered)
            if num_rendered <= cap:
                break
            if num_rendered >= 0x7fffffff:
                raise RuntimeError("g4s rasterizer: more than 2^31 - 1 (Gaussian, tile) instances in one view "
                                   "(the reference fails to allocate its sort buffers at this size)")
            cap = _capacity.bucket(num_rendered + 65536)  # the speculative launch was a no-op: re-issue
        if debug and rs.prefiltered and int(counts[3]) != 0:
            raise RuntimeError("Point is filtered although prefiltered is set. This shouldn't happen!")
    return color, others, radii, geom, binning, img, cap, num_rendered, counts, None


def _plan_and_render_pipelined(dev, dev_index, st, batch, P, W, H, bg, rs, plan):
    """view_batch(): plan + scatter + sort on the batch's side stream into a recycled scratch set, the blend on the
    current stream.  Returns None when no scratch set is free (the caller then takes the ordinary path)."""
    sset = st.pipe_sets[st.pipe_next]
    if sset.busy:
        return None
    st.pipe_next = (st.pipe_next + 1) % len(st.pipe_sets)
    f32 = dict(dtype=torch.float32, device=dev)
    if _HOST_TRACE:
        import time
        t_begin = time.perf_counter()
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        side = batch.side
        side.wait_event(batch.start)
        counts = _pinned_counts(st)
        mode = _sync_mode()
        if mode == "none":
            _check_pending(st, dev_index)
        cap = _capacity.guess(dev_index, P)
        need = (P, _LIB.g4s_geom_bytes(P), _LIB.g4s_image_bytes(W, H), _LIB.g4s_binning_bytes(cap))
        if sset.geom is None or sset.radii.numel() < need[0] or sset.geom.numel() < need[1] or sset.img.numel() < need[2] \
                or sset.binning.numel() < need[3]:
            # (re)allocate from the main stream's pool; the side stream must not touch the new blocks before everything
            # already queued on the main stream -- the possible previous owners of that memory -- has run
            sset.radii = torch.empty((need[0],), dtype=torch.int32, device=dev)
            sset.geom = torch.empty((need[1],), dtype=torch.uint8, device=dev)
            sset.img = torch.empty((need[2],), dtype=torch.uint8, device=dev)
            sset.binning = torch.empty((need[3],), dtype=torch.uint8, device=dev)
            fence = torch.cuda.Event()
            fence.record(main)
            side.wait_event(fence)
        if sset.released is not None:
            side.wait_event(sset.released)       # the view that used this set last has finished its backward
        sset.busy = True
        sset.generation += 1
        lease = _Lease(sset, dev)
        radii_s, geom, img, binning = sset.radii, sset.geom, sset.img, sset.binning
        sp = side.cuda_stream
        _lib.check(plan(radii_s, geom, img, counts, sp))
        planned = torch.cuda.Event()
        planned.record(side)
        _lib.check(_LIB.g4s_forward_bin(P, W, H, geom.data_ptr(), img.data_ptr(), binning.data_ptr(), cap, sp, 0))
        binned = torch.cuda.Event()
        binned.record(side)
        main.wait_event(binned)
        radii = radii_s[:P].clone()              # the caller's tensor: the set's copy is recycled
        color = torch.empty((NUM_CHANNELS, H, W), **f32)
        others = torch.empty((_OTHERS, H, W), **f32)
        _lib.check(_LIB.g4s_forward_blend(P, W, H, bg.data_ptr(), geom.data_ptr(), img.data_ptr(), binning.data_ptr(), cap,
                                          color.data_ptr(), others.data_ptr(), main.cuda_stream, 0))
        batch.prefetched += 1
        if mode == "none":
            st.pending_overflow.append((planned, counts, cap))
            return color, others, radii, geom, binning, img, cap, -1, counts, lease
        if _HOST_TRACE:
            t_wait = time.perf_counter()
        planned.synchronize()            # the side stream ran ahead: usually already complete
        if _HOST_TRACE:
            t_done = time.perf_counter()
            _host_times["fwd_launch"].append(t_wait - t_begin)
            _host_times["fwd_wait"].append(t_done - t_wait)
        num_rendered = int(counts[0])
        st.last_counts.update(num_rendered=num_rendered, max_tile_list=int(counts[1]), visible=int(counts[2]))
        _capacity.observe(dev_index, num_rendered)
        while num_rendered > cap:        # rare: the speculative launches were no-ops, re-issue on this stream
            if num_rendered >= 0x7fffffff:
                raise RuntimeError("g4s rasterizer: more than 2^31 - 1 (Gaussian, tile) instances in one view")
            cap = _capacity.bucket(num_rendered + 65536)
            binning = torch.empty((_LIB.g4s_binning_bytes(cap),), dtype=torch.uint8, device=dev)
            _lib.check(_LIB.g4s_forward_render(P, W, H, bg.data_ptr(), geom.data_ptr(), img.data_ptr(), binning.data_ptr(), cap,
                                               color.data_ptr(), others.data_ptr(), main.cuda_stream, 0))
    return color, others, radii, geom, binning, img, cap, num_rendered, counts, lease


def _leaf_or_constant(*tensors) -> bool:
    """True when none of the tensors was computed by an autograd-tracked operation (see view_batch)."""
    return all(t is None or t.grad_fn is None for t in tensors)


def _settle_pending(dev) -> None:
    """G4S_SYNC=none: before a backward replays a forward, wait for the capacity checks that are still owed
    (a forward that overflowed was a no-op and left its outputs and masks unwritten; the backward kernels
    refuse to run on it too, but the caller must hear about it now, not one call late)."""
    st = _state(dev)
    if st.pending_overflow:
        _check_pending(st, dev.index if dev.index is not None else torch.cuda.current_device(), wait=True)


def _resolve_sinks(named):
    """named: {kernel output name: operator input}.  Returns ({name: _Sink}, ACC_MULTIMEM or 0)."""
    if not _grad_sink:
        return None, 0
    out = {}
    for k, t in named.items():
        s = _sink_for(t)
        if s is not None:
            out[k] = s
    if not out:
        return None, 0
    mm = {s.multimem for s in out.values()}
    if len(mm) != 1:
        raise RuntimeError("gradient sinks of one call must all be multicast or all local")
    return out, (_ACC_MULTIMEM if mm.pop() else 0)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, raster_settings)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, raster_settings):
        rs = raster_settings
        if means3D.dim() != 2 or means3D.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        dev = means3D.device
        P = int(means3D.size(0))
        H, W = int(rs.image_height), int(rs.image_width)
        debug = bool(rs.debug)

        means3D_c = _f32c(means3D, "means3D")
        sh_c = _f32c(sh, "sh")
        colors_c = _f32c(colors_precomp, "colors")
        opac_c = _f32c(opacities, "opacity")
        scales_c = _f32c(scales, "scales")
        rots_c = _f32c(rotations, "rotations")
        cov_c = _f32c(cov3Ds_precomp, "transMat_precomp")
        bg = _f32c(rs.bg, "background")
        view = _f32c(rs.viewmatrix, "viewmatrix")
        proj = _f32c(rs.projmatrix, "projmatrix")
        campos = _f32c(rs.campos, "campos")
        M = int(sh_c.size(1)) if sh_c.numel() != 0 else 0

        f32 = dict(dtype=torch.float32, device=dev)
        num_rendered = 0
        ctx.lease = None
        if P == 0:
            # reference: kernels skipped, zero images returned (rasterize_points.cu:85-99)
            color = torch.zeros((NUM_CHANNELS, H, W), **f32)
            others = torch.zeros((_OTHERS, H, W), **f32)
            radii = torch.zeros((0,), dtype=torch.int32, device=dev)
            geom = binning = img = torch.empty((0,), dtype=torch.uint8, device=dev)
        else:
            def plan(radii, geom, img, counts, sp):
                return _LIB.g4s_forward_plan(
                    P, int(rs.sh_degree), M, W, H, _ptr(means3D_c), _ptr(sh_c), _ptr(colors_c),
                    _ptr(opac_c), _ptr(scales_c), float(rs.scale_modifier), _ptr(rots_c), _ptr(cov_c),
                    _ptr(view), _ptr(proj), _ptr(campos), float(rs.tanfovx), float(rs.tanfovy),
                    int(bool(rs.prefiltered)), radii.data_ptr(), geom.data_ptr(), img.data_ptr(),
                    counts.data_ptr(), sp, int(debug))
            if debug:
                # reference :83-90 -- arguments copied before the call so that a failure can be replayed
                cpu_args = _cpu_copy((rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                                      cov3Ds_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height,
                                      rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered, rs.debug))
                try:
                    color, others, radii, geom, binning, img, cap, num_rendered, counts, lease = _plan_and_render(dev, P, W, H, bg, rs, plan)
                except Exception as ex:
                    _dump_snapshot("snapshot_fw.dump", cpu_args)
                    print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                    raise ex
            else:
                color, others, radii, geom, binning, img, cap, num_rendered, counts, lease = _plan_and_render(
                    dev, P, W, H, bg, rs, plan,
                    prefetchable=_leaf_or_constant(means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp))
            ctx.counts = counts
            ctx.lease = lease

        ctx.sinks, ctx.sink_flags = _resolve_sinks({"means3D": means3D, "sh": sh, "opacities": opacities,
                                                    "scales": scales, "rotations": rotations})
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.M = M
        ctx.capacity = cap if P != 0 else 0
        ctx.save_for_backward(colors_c, means3D_c, scales_c, rots_c, cov_c, radii, sh_c, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return color, radii, others

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        if _HOST_TRACE:
            import time
            t_begin = time.perf_counter()
        rs = ctx.raster_settings
        colors_c, means3D_c, scales_c, rots_c, cov_c, radii, sh_c, geom, binning, img = ctx.saved_tensors
        dev = means3D_c.device
        _settle_pending(dev)
        lease = getattr(ctx, "lease", None)
        if lease is not None:
            lease.check()
        P = int(means3D_c.size(0))
        M = ctx.M
        H, W = int(rs.image_height), int(rs.image_width)
        f32 = dict(dtype=torch.float32, device=dev)
        alloc = torch.zeros if P == 0 else torch.empty  # every element is written by the kernels
        sinks = ctx.sinks or {}
        acc_mask = 0

        def out(name, shape):
            """(device address handed to the kernel, tensor returned to autograd)"""
            nonlocal acc_mask
            sink = sinks.get(name) if P != 0 else None
            if sink is not None:
                acc_mask |= _ACC_BITS[name] | ctx.sink_flags
                return sink.ptr, None
            t = alloc(shape, **f32)
            return _ptr(t), t

        k_means3D, dL_dmeans3D = out("means3D", (P, 3))
        k_sh, dL_dsh = out("sh", (P, M, 3))
        k_opacity, dL_dopacity = out("opacities", (P, 1))
        k_scales, dL_dscales = out("scales", (P, 2))
        k_rots, dL_drotations = out("rotations", (P, 4))
        dL_dmeans2D = alloc((P, 3), **f32)
        # gradients of inputs that were not given (empty tensors) are never read by autograd: not computed
        dL_dcolors = alloc((P, NUM_CHANNELS), **f32) if colors_c.numel() else None
        dL_dtransMat = alloc((P, 9), **f32) if cov_c.numel() else None
        if P != 0:
            if M > 0 and sh_c.numel() == 0 and dL_dsh is not None:
                dL_dsh.zero_()
            g_color = _f32c(grad_out_color, "dL_dout_color")
            g_others = _f32c(grad_depth, "dL_dout_others")
            bg = _f32c(rs.bg, "background")
            view = _f32c(rs.viewmatrix, "viewmatrix")
            proj = _f32c(rs.projmatrix, "projmatrix")
            campos = _f32c(rs.campos, "campos")
            with torch.cuda.device(dev):
                sp = torch.cuda.current_stream(dev).cuda_stream
                scratch = torch.empty((_LIB.g4s_backward_scratch_bytes(P),), dtype=torch.uint8, device=dev)

                def call():
                    _lib.check(_LIB.g4s_backward(
                        P, int(rs.sh_degree), M, W, H, bg.data_ptr(), _ptr(means3D_c), _ptr(sh_c), _ptr(colors_c),
                        _ptr(scales_c), float(rs.scale_modifier), _ptr(rots_c), _ptr(cov_c), _ptr(view), _ptr(proj),
                        _ptr(campos), float(rs.tanfovx), float(rs.tanfovy), radii.data_ptr(), geom.data_ptr(),
                        binning.data_ptr(), int(ctx.capacity), img.data_ptr(), g_color.data_ptr(), g_others.data_ptr(),
                        k_means3D, dL_dmeans2D.data_ptr(), k_sh, _ptr(dL_dcolors),
                        k_opacity, k_scales, k_rots,
                        _ptr(dL_dtransMat), acc_mask, scratch.data_ptr(), sp, int(bool(rs.debug))))
                if rs.debug:
                    # reference :133-140
                    cpu_args = _cpu_copy((rs.bg, means3D_c, radii, colors_c, scales_c, rots_c, rs.scale_modifier, cov_c,
                                          rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, grad_depth,
                                          sh_c, rs.sh_degree, rs.campos, geom, ctx.num_rendered, binning, img, rs.debug))
                    try:
                        call()
                    except Exception as ex:
                        _dump_snapshot("snapshot_bw.dump", cpu_args)
                        print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                        raise ex
                else:
                    call()
        if lease is not None:
            lease.release()      # the scratch set of this view may be recycled once the kernels above have run
        if _HOST_TRACE:
            _host_times["bwd"].append(time.perf_counter() - t_begin)
        # same order as the reference (RAST/diff_surfel_rasterization/__init__.py:144-154)
        return (dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors, dL_dopacity, dL_dscales, dL_drotations,
                dL_dtransMat, None)


# ---- raw-parameter operator (extension) ----------------------------------------
# The reference trainer activates its parameters with ~10 torch kernels (+ autograd) before every
# operator call: exp / sqrt(s^2 + f^2), sigmoid (* mip compensation), normalize, cat(features_dc,
# features_rest)  (2DGS/scene/gaussian_model.py:158-192).  `rasterize_gaussian_model` takes the
# optimiser's leaves as they are stored and returns gradients with respect to them; the activations
# run in registers inside the projection kernels (g4s_forward_plan_raw / g4s_backward_raw).
def rasterize_gaussian_model(xyz, means2D, features_dc, features_rest, opacity, scaling, rotation, mip_filter,
                             raster_settings):
    """(color, radii, allmap) from un-activated parameters: xyz[P,3], features_dc[P,1,3],
    features_rest[P,M-1,3], opacity[P,1], scaling[P,2], rotation[P,4] as GaussianModel stores them
    (_xyz, _features_dc, _features_rest, _opacity, _scaling, _rotation) and mip_filter[P,1] or None
    (use_mip_filter off).  Same outputs as GaussianRasterizer on the activated tensors."""
    return _RasterizeGaussianModel.apply(xyz, means2D, features_dc, features_rest, opacity, scaling, rotation,
                                         mip_filter, raster_settings)


class _RasterizeGaussianModel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, means2D, features_dc, features_rest, opacity, scaling, rotation, mip_filter, raster_settings):
        rs = raster_settings
        if xyz.dim() != 2 or xyz.size(1) != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        dev = xyz.device
        P = int(xyz.size(0))
        H, W = int(rs.image_height), int(rs.image_width)
        xyz_c = _f32c(xyz, "xyz")
        dc_c = _f32c(features_dc, "features_dc")
        rest_c = _f32c(features_rest, "features_rest")
        opac_c = _f32c(opacity, "opacity")
        scal_c = _f32c(scaling, "scaling")
        rot_c = _f32c(rotation, "rotation")
        mip_c = _f32c(mip_filter, "mip_filter") if mip_filter is not None else None
        if P and (dc_c.shape != (P, 1, 3) or rest_c.dim() != 3 or rest_c.shape[0] != P or rest_c.shape[2] != 3 or
                  opac_c.numel() != P or scal_c.shape != (P, 2) or rot_c.shape != (P, 4) or
                  (mip_c is not None and mip_c.numel() != P)):
            raise RuntimeError("rasterize_gaussian_model: parameter shapes do not match GaussianModel's "
                               "(_features_dc [P,1,3], _features_rest [P,M-1,3], _opacity [P,1], _scaling [P,2], _rotation [P,4])")
        M = 1 + int(rest_c.shape[1]) if P else 1
        bg = _f32c(rs.bg, "background")
        view = _f32c(rs.viewmatrix, "viewmatrix")
        proj = _f32c(rs.projmatrix, "projmatrix")
        campos = _f32c(rs.campos, "campos")
        f32 = dict(dtype=torch.float32, device=dev)
        num_rendered, cap = 0, 0
        ctx.lease = None
        if P == 0:
            color = torch.zeros((NUM_CHANNELS, H, W), **f32)
            others = torch.zeros((_OTHERS, H, W), **f32)
            radii = torch.zeros((0,), dtype=torch.int32, device=dev)
            geom = binning = img = torch.empty((0,), dtype=torch.uint8, device=dev)
        else:
            def plan(radii, geom, img, counts, sp):
                return _LIB.g4s_forward_plan_raw(
                    P, int(rs.sh_degree), M, W, H, xyz_c.data_ptr(), dc_c.data_ptr(), _ptr(rest_c), opac_c.data_ptr(),
                    scal_c.data_ptr(), float(rs.scale_modifier), rot_c.data_ptr(), _ptr(mip_c), _ptr(view), _ptr(proj),
                    _ptr(campos), float(rs.tanfovx), float(rs.tanfovy), int(bool(rs.prefiltered)), radii.data_ptr(),
                    geom.data_ptr(), img.data_ptr(), counts.data_ptr(), sp, int(bool(rs.debug)))
            color, others, radii, geom, binning, img, cap, num_rendered, counts, lease = _plan_and_render(
                dev, P, W, H, bg, rs, plan,
                prefetchable=_leaf_or_constant(xyz, features_dc, features_rest, opacity, scaling, rotation, mip_filter))
            ctx.counts = counts
            ctx.lease = lease
        # the SH gradient is one kernel output mode for both tensors: they are sunk together or not at all
        ctx.sinks, ctx.sink_flags = _resolve_sinks({"means3D": xyz, "sh": features_dc, "sh_rest": features_rest,
                                                    "opacities": opacity, "scales": scaling, "rotations": rotation})
        if ctx.sinks is not None and M > 1 and (("sh" in ctx.sinks) != ("sh_rest" in ctx.sinks)):
            ctx.sinks.pop("sh", None)
            ctx.sinks.pop("sh_rest", None)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.M = M
        ctx.capacity = cap
        ctx.has_mip = mip_c is not None
        ctx.save_for_backward(xyz_c, dc_c, rest_c, opac_c, scal_c, rot_c, mip_c if mip_c is not None else xyz_c.new_empty(0),
                              radii, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return color, radii, others

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_depth):
        rs = ctx.raster_settings
        xyz_c, dc_c, rest_c, opac_c, scal_c, rot_c, mip_c, radii, geom, binning, img = ctx.saved_tensors
        dev = xyz_c.device
        P, M = int(xyz_c.size(0)), ctx.M
        H, W = int(rs.image_height), int(rs.image_width)
        f32 = dict(dtype=torch.float32, device=dev)
        alloc = torch.zeros if P == 0 else torch.empty  # every element is written by the kernels
        _settle_pending(dev)
        lease = getattr(ctx, "lease", None)
        if lease is not None:
            lease.check()
        sinks = ctx.sinks or {}
        acc_mask = 0

        def out(name, shape, bit):
            nonlocal acc_mask
            sink = sinks.get(name) if P != 0 else None
            if sink is not None:
                acc_mask |= bit | ctx.sink_flags
                return sink.ptr, None
            t = alloc(shape, **f32)
            return _ptr(t), t

        k_xyz, g_xyz = out("means3D", (P, 3), _ACC_BITS["means3D"])
        k_dc, g_dc = out("sh", (P, 1, 3), _ACC_BITS["sh"])
        k_rest, g_rest = out("sh_rest", (P, M - 1, 3), _ACC_BITS["sh"])
        k_op, g_op = out("opacities", (P, 1), _ACC_BITS["opacities"])
        k_sc, g_sc = out("scales", (P, 2), _ACC_BITS["scales"])
        k_rot, g_rot = out("rotations", (P, 4), _ACC_BITS["rotations"])
        g_means2D = alloc((P, 3), **f32)
        if P != 0:
            g_color = _f32c(grad_out_color, "dL_dout_color")
            g_others = _f32c(grad_depth, "dL_dout_others")
            bg = _f32c(rs.bg, "background")
            view = _f32c(rs.viewmatrix, "viewmatrix")
            proj = _f32c(rs.projmatrix, "projmatrix")
            campos = _f32c(rs.campos, "campos")
            with torch.cuda.device(dev):
                sp = torch.cuda.current_stream(dev).cuda_stream
                scratch = torch.empty((_LIB.g4s_backward_scratch_bytes_raw(P),), dtype=torch.uint8, device=dev)
                _lib.check(_LIB.g4s_backward_raw(
                    P, int(rs.sh_degree), M, W, H, bg.data_ptr(), xyz_c.data_ptr(), dc_c.data_ptr(), _ptr(rest_c),
                    opac_c.data_ptr(), scal_c.data_ptr(), float(rs.scale_modifier), rot_c.data_ptr(),
                    mip_c.data_ptr() if ctx.has_mip else None, _ptr(view), _ptr(proj), _ptr(campos),
                    float(rs.tanfovx), float(rs.tanfovy), radii.data_ptr(), geom.data_ptr(), binning.data_ptr(),
                    int(ctx.capacity), img.data_ptr(), g_color.data_ptr(), g_others.data_ptr(), k_xyz,
                    g_means2D.data_ptr(), k_dc, k_rest, k_op, k_sc,
                    k_rot, acc_mask, scratch.data_ptr(), sp, int(bool(rs.debug))))
        if lease is not None:
            lease.release()
        return g_xyz, g_means2D, g_dc, g_rest, g_op, g_sc, g_rot, None, None


def debug_pair_stats(rendered: torch.Tensor) -> dict:
    """Work counters of the view that produced `rendered` (the colour or allmap output of a call that
    required grad): pairs blended, sum over pixels of the last contributor's position in the culled lists,
    pair slots (256 per list entry), longest tile list, pair evaluations the B200 backward issues (g4s_debug_pair_stats).  Benchmark / test introspection;
    no reference counterpart (SURVEY.md 8d asks for pairs/s against the fp32 peak)."""
    ctx = rendered.grad_fn
    if ctx is None or not hasattr(ctx, "raster_settings"):
        raise RuntimeError("debug_pair_stats needs an output of the rasterizer that is attached to the graph")
    saved = ctx.saved_tensors
    geom, binning, img = saved[-3:]
    rs = ctx.raster_settings
    P = int(saved[1].shape[0])
    keys = ("pairs_blended", "pairs_walked", "pair_slots", "longest_tile_list", "pair_evals_bwd", "block_entry_hits")
    if P == 0:
        return dict.fromkeys(keys, 0)
    lib = _lib.load()
    stats = torch.zeros(8, dtype=torch.int64, device=geom.device)
    with torch.cuda.device(geom.device):
        _lib.check(lib.g4s_debug_pair_stats(int(rs.image_width), int(rs.image_height), geom.data_ptr(), P, img.data_ptr(),
                                            binning.data_ptr(), int(ctx.capacity), stats.data_ptr(),
                                            torch.cuda.current_stream(geom.device).cuda_stream))
    return dict(zip(keys, (int(v) for v in stats[:6].tolist())))


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points with view-space z > 0.2 (reference :177-186)."""
        with torch.no_grad():
            rs = self.raster_settings
            pos = _f32c(positions, "means3D")
            P = int(pos.size(0))
            present = torch.zeros((P,), dtype=torch.bool, device=pos.device)
            if P != 0:
                view = _f32c(rs.viewmatrix, "viewmatrix")
                proj = _f32c(rs.projmatrix, "projmatrix")
                with torch.cuda.device(pos.device):
                    sp = torch.cuda.current_stream(pos.device).cuda_stream
                    _lib.check(_LIB.g4s_mark_visible(P, pos.data_ptr(), view.data_ptr(), proj.data_ptr(),
                                                     present.data_ptr(), sp))
        return present

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None):
        raster_settings = self.raster_settings

        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')

        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')

        empty = torch.empty((0,), dtype=torch.float32, device=means3D.device)
        if shs is None:
            shs = empty
        if colors_precomp is None:
            colors_precomp = empty
        if scales is None:
            scales = empty
        if rotations is None:
            rotations = empty
        if cov3D_precomp is None:
            cov3D_precomp = empty

        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, raster_settings)
