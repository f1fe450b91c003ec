#!/usr/bin/env python
"""bench.py -- forward+backward Gaussians/s of the surfel rasterizer hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (all N): BASELINE.json config 2 -- synthetic "room" scan, 1.0 M surfels, SH degree 3,
1920x1080, --views views per GPU per step cycled from a ring of 64 perimeter cameras; every
rank holds the full scene, renders its own views forward+backward through the operator API
(gradients accumulate in the parameter leaves) and, for N > 1, joins ONE NCCL all-reduce of the
[P,60] gradient + densification-statistics buffer per step (weak scaling: per-GPU work fixed).

  value   = P * views * N * K / t   with parameters, cameras and upstream gradients resident in
            HBM; t = CUDA-event time of the K steps, max over ranks.
  e2e     = same through the public API starting from HOST buffers: per view one pinned uint8
            ground-truth image is copied host->device, an L1 photometric + regulariser loss is
            formed on the device, and the scalar loss is read back device->host.
  roofline= dominant kernel of the step (per-stage CUDA-event timers inside the C library,
            averaged over the timed region) against the measured HBM peak; algorithmic bytes per
            stage are spelled out in DESIGN.md and in `algorithmic_bytes()` below.
  cpu_baseline = the CPU oracle port (oracle/surfel_oracle.c, OpenMP on all host cores) on six
            full views of the same workload (~10 s on the GPU box's 16 cores).

--impl reference runs the UNMODIFIED reference extension (oracle/_ref, the vendored
diff-surfel-rasterization compiled for sm_100a) through its own Python API in the same loops;
when that build is absent it falls back to timing the CPU oracle port.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "gaussians_per_s_fwd_bwd_1080p"
UNIT = "Gaussians/s"
CONFIG = "c2"           # default workload: the configuration BASELINE.json's metric is quoted on
CAM_RING = 64
# BASELINE.json configs 3-5: how a "step" is formed and how it scales with the rank count
CONFIG_DOC = {
    "c2": dict(name="c2: 1.0M surfels, SH degree 3, 1920x1080", scaling="weak", ring=64, step_views=None),
    "c3": dict(name="c3: 2.5M surfels, SH degree 3, 1600x1200, 50 views per step sharded over the ranks",
               scaling="strong", ring=50, step_views=50),
    "c4": dict(name="c4: 5.0M surfels, SH degree 3, 1920x1080, 64 cameras", scaling="weak", ring=64, step_views=None),
}


# ------------------------------------------------------------------------------------------ util
def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled every 40 ms while the timed region runs (the region is ~0.3 s).

    Source: the NVML calls behind `nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.*`
    made from a thread of this process (nvidia_ml_py); G4S_BENCH_CLOCKS=smi spawns nvidia-smi itself
    (`-lms 200`, the profiling recipe's line), =off disables sampling.  In-process NVML is the default
    because a polling nvidia-smi process that lands on the core of the launching thread stretches the
    (host-synchronised) step loop by tens of percent in some runs (profiles/r01t_bench_stability.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.mode = os.environ.get("G4S_BENCH_CLOCKS", "nvml")
        self.proc = None
        self.lines = []
        self.samples = []      # (sm_mhz, max_mhz, reason bits)
        self._stop = threading.Event()
        self._thread = None

    # physical index of the visible device `idx` (CUDA_VISIBLE_DEVICES may renumber)
    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.idx < len(ids) and ids[self.idx].isdigit():
                return int(ids[self.idx])
        return self.idx

    def start(self):
        if self.mode == "off":
            return
        if self.mode == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                self._h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
                self._nv = pynvml
                self._thread = threading.Thread(target=self._poll_nvml, daemon=True)
                self._thread.start()
                return
            except Exception:  # noqa: BLE001
                self.mode = "smi"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self._physical_index())], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _poll_nvml(self):
        nv = self._nv
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((float(sm), float(mx), {n for n, b in bits.items() if r & b}))
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.04)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.mode == "off":
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampling disabled (G4S_BENCH_CLOCKS=off)"]}
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1.0)
            sm = [s[0] for s in self.samples]
            reasons = set().union(*[s[2] for s in self.samples]) if self.samples else set()
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(s[1] for s in self.samples) if sm else None,
                    "reasons": sorted(reasons), "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def algorithmic_bytes(stage: str, P: int, V: int, R: int, N: int, T: int, K: int, M: int, sink: bool = True) -> float:
    """Compulsory bytes per launch of each stage (DESIGN.md 'Kernels'): every input read once,
    every output written once, every (Gaussian, tile) instance written once and read once.
    sink: project_bwd adds the rows of VISIBLE Gaussians into the multi-view sum (read-modify-write) and writes
    no dense per-view gradient except dL_dmeans2D; otherwise it writes every row of every gradient tensor."""
    rec = 112 + 4 + 4 + 8 + 1 + 4                  # record + depth + ntiles + rect + clamp mask + visible-list entry
    g = 12 + 12 * M + 4 + 8 + 16                   # gradient row: means3D, SH, opacity, scales, rotations
    if sink:
        # dL_dmeans2D cleared and its visible rows stored | list entry, radius, acc, record, inputs, SH; RMW of the row
        project_bwd = P * 12 + V * (4 + 4 + 12 + 96 + 48 + 40 + 12 * K + 2 * g)
    else:
        project_bwd = P * (4 + 12 + g) + V * (96 + 48 + 40 + 12 * K)
    return {
        "project_fwd": P * (40 + 4) + V * (12 * K + rec) + R * 4,
        "tile_scan": T * 16,
        "scatter": V * 16 + R * (8 + 4),
        "tile_sort": R * (8 + 4) + T * 8,
        "blend_fwd": R * (4 + 32) + V * 112 + N * 60 + T * 12,   # list + masks written; each visible record once; images
        "acc_clear": P * 4 + V * 96,
        "blend_bwd": R * (4 + 32) + V * (80 + 96) + N * 60 + T * 12,   # list + masks read; records; 96-byte accumulator rows
        "project_bwd": project_bwd,
    }[stage]


def path_bytes(P, V, R, N, K, M):
    """Whole-path figure of SURVEY.md 8d: P(96+12M) + V(40+24K) + 80N + 20R."""
    return P * (96 + 12 * M) + V * (40 + 24 * K) + 80 * N + 20 * R


# ------------------------------------------------------------------------------------- workload
def cached_scene(cfg_name: str, P: int, seed: int, rank: int, world: int):
    """The seeded synthetic scene.  Generating 5 M surfels takes ~20 s of numpy per process; on a multi-rank run
    local rank 0 generates it once into /dev/shm and the others read it (same bytes: the generator is seeded)."""
    from g4splat_b200 import synthetic as S
    if P < 2_000_000:
        return S.make_scene(P, seed)
    path = Path("/dev/shm") / f"g4s_scene_{cfg_name}_{P}_{seed}.npz"
    keys = ("means3D", "scales", "rotations", "opacities", "shs")
    if not path.exists() and rank == 0:
        sc = S.make_scene(P, seed)
        tmp = path.with_suffix(".tmp.npz")
        np.savez(tmp, **sc)
        os.replace(tmp, path)
        return sc
    if world > 1 or path.exists():
        deadline = time.time() + 600
        while not path.exists():
            if time.time() > deadline:
                return S.make_scene(P, seed)
            time.sleep(0.5)
        z = np.load(path)
        return {k: z[k] for k in keys}
    return S.make_scene(P, seed)


class Workload:
    def __init__(self, device, views_per_step: int, rank: int, world: int, cfg_name: str = CONFIG):
        import torch
        from g4splat_b200 import synthetic as S
        cfg = S.CONFIGS[cfg_name]
        self.cfg, self.device, self.vps, self.rank, self.world = cfg, device, views_per_step, rank, world
        self.P, self.W, self.H = cfg["P"], cfg["W"], cfg["H"]
        sc = cached_scene(cfg_name, self.P, cfg["seed"], rank, world)
        t = lambda a: torch.from_numpy(a).to(device).requires_grad_(True)
        # 58 floats per Gaussian (xyz 3, SH 48, opacity 1, scaling 2, rotation 4), already activated:
        # the operator's inputs.  (The trainer's activations / cat of features_dc and features_rest sit
        # outside the operator boundary -- SURVEY.md 8f "next" rows.)
        self.params = {"xyz": t(sc["means3D"]), "features": t(sc["shs"]), "opacity": t(sc["opacities"]),
                       "scaling": t(sc["scales"]), "rotation": t(sc["rotations"])}
        self.doc = CONFIG_DOC[cfg_name]
        self.ring = self.doc["ring"]
        self.cams = S.make_cameras(self.ring, self.W, self.H)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
        self.cam_dev = [dict(view=d(c.viewmatrix), proj=d(c.projmatrix), campos=d(c.campos)) for c in self.cams]
        self.bg = torch.zeros(3, device=device)
        gc, go = S.make_upstream_grads(self.W, self.H, cfg["seed"])
        self.g_color, self.g_allmap = d(gc), d(go)
        # e2e regulariser weights per allmap channel (depth, alpha, normal xyz, median depth, distortion), / pixels
        n_pix = float(self.W * self.H)
        self.reg_weights = torch.tensor([0.01, 0.0025, 0.0025, 0.0025, 0.0025, 0.01, 0.05], device=device).view(7, 1, 1) / n_pix
        # e2e: 8 distinct pinned uint8 "photographs" (random content; the loss only needs the bytes to move)
        rng = np.random.default_rng(123 + rank)
        self.gt_host = [torch.from_numpy(rng.integers(0, 256, size=(3, self.H, self.W), dtype=np.uint8)).pin_memory()
                        for _ in range(8)]
        # double-buffered device copies filled by a side stream: the next view's image travels over
        # PCIe while the current view is rendered (what a data loader with pinned memory does)
        self.gt_dev = [torch.empty((3, self.H, self.W), dtype=torch.uint8, device=device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.copy_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.slot_free = [torch.cuda.Event(), torch.cuda.Event()]
        self._slot = 0

    def prefetch_image(self, index: int):
        """Start the host->device copy of photograph `index` into the next free slot (side stream)."""
        import torch
        slot = self._slot
        self._slot ^= 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.slot_free[slot])   # the loss that last read this slot has been issued
            self.gt_dev[slot].copy_(self.gt_host[index % len(self.gt_host)], non_blocking=True)
            self.copy_done[slot].record(self.copy_stream)
        return slot

    def image(self, slot: int):
        import torch
        torch.cuda.current_stream(self.device).wait_event(self.copy_done[slot])
        return self.gt_dev[slot]

    def release_image(self, slot: int):
        import torch
        self.slot_free[slot].record(torch.cuda.current_stream(self.device))

    def view_ids(self, step: int, rank: int = None, world: int = None):
        """Cameras this rank renders in step `step` (rank / world default to the process's own)."""
        rank = self.rank if rank is None else rank
        world = self.world if world is None else world
        # the step's views are dealt out to the ranks (shard_views(strided=True)): neighbouring ring cameras cost
        # about the same, a contiguous block per rank would make the step wait for the rank with the expensive arc
        from g4splat_b200.view_parallel import shard_views
        n = self.doc["step_views"] if self.doc["step_views"] is not None else self.vps * world
        return [(step * n + i) % self.ring for i in shard_views(n, world, rank, strided=True)]

    def views_per_step_total(self):
        return self.doc["step_views"] if self.doc["step_views"] is not None else self.vps * self.world

    def settings(self, mod, vid: int):
        c, cd = self.cams[vid], self.cam_dev[vid]
        return mod.GaussianRasterizationSettings(
            image_height=self.H, image_width=self.W, tanfovx=c.tanfovx, tanfovy=c.tanfovy, bg=self.bg,
            scale_modifier=1.0, viewmatrix=cd["view"], projmatrix=cd["proj"], sh_degree=3, campos=cd["campos"],
            prefiltered=False, debug=False)

    def rasterize(self, mod, vid: int):
        import torch
        p = self.params
        means2D = torch.zeros_like(p["xyz"], requires_grad=True)
        shs = p["features"]
        rast = mod.GaussianRasterizer(raster_settings=self.settings(mod, vid))
        color, radii, allmap = rast(means3D=p["xyz"], means2D=means2D, opacities=p["opacity"], shs=shs,
                                    scales=p["scaling"], rotations=p["rotation"])
        return color, radii, allmap, means2D


# G4S_BENCH_PIPELINE=1: view_batch() around the views of a step (B200 arm).  Off by default: measured on a B200 it
# gains 2.9 % (621.9 -> 639.9 M Gaussians/s; the blend kernels already fill 70-77 % of the issue slots);
# the headline runs the plain single-stream loop.
PIPELINE = os.environ.get("G4S_BENCH_PIPELINE", "0") == "1"
STEP_TRACE = []   # G4S_BENCH_TRACE=1: a CUDA event after every step (diagnostics; read after the timed region)


def run_steps(wl: Workload, mod, sync, steps: int, first_step: int, e2e: bool):
    """`steps` optimisation-step-shaped passes; returns the last loss value (e2e) or None."""
    import torch
    last = None
    trace = os.environ.get("G4S_BENCH_TRACE") == "1"
    slot = wl.prefetch_image(first_step) if e2e else None
    for s in range(first_step, first_step + steps):
        if trace:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            STEP_TRACE.append((s, e2e, ev, op_counts_snapshot(mod)))
        loss_acc = None
        # the views of a step share one parameter set: the B200 operator may run each view's front end ahead of the
        # previous view's blend kernels (view_batch); the reference operator has no such mode
        with (mod.view_batch() if PIPELINE and hasattr(mod, "view_batch") else contextlib.nullcontext()):
            for k, vid in enumerate(wl.view_ids(s)):
                if e2e:
                    nxt = wl.prefetch_image(s + k + 1)           # next view's photograph: PCIe copy overlaps this view
                color, radii, allmap, means2D = wl.rasterize(mod, vid)
                if e2e:
                    gt = torch.mul(wl.image(slot), 1.0 / 255.0)      # uint8 -> float32 in one kernel
                    wl.release_image(slot)
                    slot = nxt
                    # L1 photometric term + regularisers on the seven allmap channels:
                    #   mean|color - gt| + 0.05 mean(distortion) + 0.01 mean(depth + median depth) + 0.01 mean(alpha, normal)
                    # written as one weighted sum so that both arms spend few launches on it
                    loss = torch.nn.functional.l1_loss(color, gt) + (allmap * wl.reg_weights).sum()
                    loss.backward()
                    loss_acc = loss.detach() if loss_acc is None else loss_acc + loss.detach()
                else:
                    torch.autograd.backward([color, allmap], [wl.g_color, wl.g_allmap])
                sync.add_view_stats(means2D.grad, radii)
        sync.allreduce()
        if e2e:
            last = float(loss_acc.item())  # device -> host read of the step's result
        sync.zero()
    return last


# ------------------------------------------------------------- whole training iteration (8f rows)
class RawModel:
    """The optimiser's leaves as GaussianModel stores them (scene/gaussian_model.py:253-260), derived once
    from the workload's activated parameters, and the properties render() reads.  The properties are the
    reference's own torch expressions (gaussian_model.py:158-192)."""

    def __init__(self, wl: "Workload"):
        import torch
        p = wl.params
        with torch.no_grad():
            op = p["opacity"].clamp(1e-4, 1 - 1e-4)
            leaves = {"_xyz": p["xyz"].clone(), "_features_dc": p["features"][:, :1].contiguous(),
                      "_features_rest": p["features"][:, 1:].contiguous(), "_opacity": torch.log(op / (1 - op)),
                      "_scaling": torch.log(p["scaling"]), "_rotation": p["rotation"].clone()}
        for k, v in leaves.items():
            setattr(self, k, v.detach().requires_grad_(True))
        self.active_sh_degree = self.max_sh_degree = 3
        self.use_mip_filter = False

    def leaves(self):
        return [self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling, self._rotation]

    get_xyz = property(lambda s: s._xyz)
    get_scaling = property(lambda s: __import__("torch").exp(s._scaling))
    get_rotation = property(lambda s: __import__("torch").nn.functional.normalize(s._rotation))
    get_features = property(lambda s: __import__("torch").cat((s._features_dc, s._features_rest), dim=1))
    get_opacity = property(lambda s: __import__("torch").sigmoid(s._opacity))


def train_iteration(wl: "Workload", model: RawModel, mod, vid: int, gt, fused: bool):
    """One view of train_with_refine_depth.py:378-399 from the raw leaves: activations -> rasterizer -> render()'s
    post-processing -> 0.8 L1 + 0.2 D-SSIM + 0.05 normal consistency + 100 distortion -> backward.
    fused=True: this repository's kernels for every row (rasterize_gaussian_model, surface_attributes,
    photometric_loss); fused=False: the reference's torch operator sequences around `mod`'s rasterizer
    (oracle/{surface,loss}_oracle.py restate gaussian_renderer/__init__.py:118-164 and utils/loss_utils.py)."""
    import torch
    c, cd = wl.cams[vid], wl.cam_dev[vid]
    settings = wl.settings(mod, vid)
    means2D = torch.zeros_like(model._xyz, requires_grad=True)
    if fused:
        from g4splat_b200.diff_surfel_rasterization import rasterize_gaussian_model
        from g4splat_b200.surface import surface_attributes
        from g4splat_b200.loss_utils import photometric_loss
        color, radii, allmap = rasterize_gaussian_model(model._xyz, means2D, model._features_dc, model._features_rest,
                                                        model._opacity, model._scaling, model._rotation, None, settings)
        pkg = surface_attributes(allmap, cd["view"], cd["proj"], 0.0)
        loss, _ = photometric_loss(color, gt, 0.2)
    else:
        from oracle import loss_oracle as LO
        from oracle import surface_oracle as SO
        rast = mod.GaussianRasterizer(raster_settings=settings)
        color, radii, allmap = rast(means3D=model.get_xyz, means2D=means2D, opacities=model.get_opacity, shs=model.get_features,
                                    scales=model.get_scaling, rotations=model.get_rotation)
        pkg = SO.surface_attributes(allmap, cd["view"], cd["proj"], 0.0)
        loss = 0.8 * LO.l1_loss(color, gt) + 0.2 * (1.0 - LO.ssim(color, gt))
    normal_error = (1 - (pkg["rend_normal"] * pkg["surf_normal"]).sum(dim=0))[None]
    total = loss + 0.05 * normal_error.mean() + 100.0 * pkg["rend_dist"].mean()
    total.backward()
    return total.detach()


def time_train_iteration(wl: "Workload", mod, fused: bool, views: int, device):
    """ms per view of `train_iteration` over `views` ring cameras (photograph resident on the device)."""
    import torch
    model = RawModel(wl)
    gt = wl.gt_host[0].to(device).to(torch.float32) * (1.0 / 255.0)
    for v in range(3):
        train_iteration(wl, model, mod, v, gt, fused)
    for leaf in model.leaves():
        leaf.grad = None
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for v in range(views):
        train_iteration(wl, model, mod, (3 + v) % wl.ring, gt, fused)
    e1.record()
    torch.cuda.synchronize(device)
    return e0.elapsed_time(e1) / views


def time_operator(wl: "Workload", mod, device, views: int = 20, warm: int = 3):
    """SURVEY.md 8(d) timing protocol, per view through the plain operator API (no gradient sink, dense gradients
    returned to autograd as the reference does): 3 warm-up + 20 timed views, CUDA events on the current stream, the
    forward and the backward timed separately and together, output allocation included; median / p10 / p90 in ms.
    The views run one at a time with a host synchronise between them, so each number is an operator LATENCY (launch
    latency and the mid-forward count read-back included), not the throughput of the step loop above."""
    import torch
    rows = []
    for i in range(warm + views):
        vid = (5 * i + 1) % wl.ring
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        for p_ in wl.params.values():
            p_.grad = None          # autograd keeps the returned tensors as they are: no accumulation pass inside the timing
        torch.cuda.synchronize(device)
        e0.record()
        color, radii, allmap, means2D = wl.rasterize(mod, vid)
        e1.record()
        torch.autograd.backward([color, allmap], [wl.g_color, wl.g_allmap])
        e2.record()
        torch.cuda.synchronize(device)
        if i >= warm:
            rows.append((e0.elapsed_time(e1), e1.elapsed_time(e2), e0.elapsed_time(e2)))
        del color, radii, allmap, means2D
    for p_ in wl.params.values():
        p_.grad = None
    a = np.array(rows)
    q = lambda col: {"median": float(np.median(a[:, col])), "p10": float(np.percentile(a[:, col], 10)),
                     "p90": float(np.percentile(a[:, col], 90))}
    return {"unit": "ms per view", "views": views, "warmup": warm, "forward": q(0), "backward": q(1), "forward_backward": q(2),
            "gaussians_per_s_median": wl.P / (float(np.median(a[:, 2])) * 1e-3),
            "what": "plain operator API, one view at a time with a host synchronise in between (SURVEY.md 8d protocol)"}


def op_counts_snapshot(mod):
    lc = getattr(mod, "last_counts", None)
    return dict(lc) if lc else None


def dump_trace():
    if not STEP_TRACE:
        return
    rows = []
    for (s0, e0, ev0, c0), (s1, e1, ev1, c1) in zip(STEP_TRACE[:-1], STEP_TRACE[1:]):
        rows.append(f"step {s0:5d} {'e2e' if e0 else 'dev'} {ev0.elapsed_time(ev1):7.2f} ms  {c1}")
    sys.stderr.write("\n".join(rows) + "\n")


def time_region(fn, device, dist_on):
    import gc
    import torch
    import torch.distributed as dist
    # no cyclic-GC pause inside the timed region (the step loop is host-synchronised once per view, so a
    # 50 ms generation-2 collection shows up one to one); reference counting still frees every tensor
    gc.collect()
    gc.disable()
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize(device)
    gc.enable()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms, out


def allreduce_self_check(wl: "Workload", mod, sync, device, step: int = 5):
    """N ranks, one step, gradients + statistics brought together by `sync`'s transport  ==  ONE rank looping
    over the views of all ranks (same kernels, local accumulation), within 1e-4 of each block's largest entry.
    Run after the timed regions; every rank computes the single-rank loop itself."""
    import torch
    import torch.distributed as dist
    from g4splat_b200.view_parallel import ViewShardedGradSync

    def render(s, vid):
        color, radii, allmap, means2D = wl.rasterize(mod, vid)
        torch.autograd.backward([color, allmap], [wl.g_color, wl.g_allmap])
        s.add_view_stats(means2D.grad, radii)

    sync.zero()
    for vid in wl.view_ids(step):
        render(sync, vid)
    sync.allreduce()
    torch.cuda.synchronize(device)
    got = {k: v.detach().clone() for k, v in sync._views.items()}
    got.update(accum=sync._accum.clone(), denom=sync._denom.clone(), max_radii=sync.max_radii.clone())
    local = ViewShardedGradSync(wl.params, transport="nccl")   # never all-reduced: a single-rank accumulator
    local.bind(mod)
    for r in range(wl.world):
        for vid in wl.view_ids(step, r, wl.world):
            render(local, vid)
    torch.cuda.synchronize(device)
    want = dict(local._views)
    want.update(accum=local._accum, denom=local._denom, max_radii=local.max_radii)
    worst, per = 0.0, {}
    for k in want:
        w, g = want[k].double(), got[k].double()
        err = float((w - g).abs().max() / w.abs().max().clamp_min(1e-30))
        per[k] = err
        worst = max(worst, err)
    t = torch.tensor([worst], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sync.attach()
    sync.bind(mod)
    sync.zero()
    return {"max_rel_err": float(t.item()), "ok": bool(t.item() <= 1e-4), "tolerance": 1e-4, "per_block": per,
            "what": f"{wl.world}-rank step vs one rank looping over the same {wl.views_per_step_total()} views"}


# ----------------------------------------------------------------------------------- CPU oracle
def cpu_oracle_step(cfg_name: str = CONFIG, threads: int = 0, views: int = 6):
    """`views` full views (forward + backward) of the workload on the host cores with the oracle port: a
    bounded sample of ~10 s on the GPU box's 16 cores.  Returns (Gaussians/s, cores, seconds, views)."""
    from g4splat_b200 import synthetic as S
    from oracle.oracle import Oracle
    cfg = S.CONFIGS[cfg_name]
    o = Oracle("f32")
    cores = o.set_threads(threads if threads > 0 else (os.cpu_count() or 1))
    sc = S.make_scene(cfg["P"], cfg["seed"])
    cams = S.make_cameras(CAM_RING, cfg["W"], cfg["H"])
    gc, go = S.make_upstream_grads(cfg["W"], cfg["H"], cfg["seed"])
    dt = 0.0
    for v in range(views):
        cam = cams[(v * 11) % CAM_RING]
        t0 = time.perf_counter()
        st = o.forward(means3D=sc["means3D"], opacities=sc["opacities"], view=cam.viewmatrix, proj=cam.projmatrix,
                       campos=cam.campos, W=cam.W, H=cam.H, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=np.zeros(3, np.float32),
                       shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"], sh_degree=3)
        o.backward(st, gc, go)
        dt += time.perf_counter() - t0
    return cfg["P"] * views / dt, cores, dt, views


# ----------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--views", type=int, default=8, help="views per GPU per step")
    ap.add_argument("--config", default=CONFIG, choices=sorted(CONFIG_DOC), help="BASELINE.json workload (default c2)")
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "multimem", "multimem_red"],
                    help="N > 1: how the ranks' gradient sums meet (g4splat_b200/view_parallel.py)")
    ap.add_argument("--math", default="exact", choices=["exact", "fast"],
                    help="forward blend arithmetic: exact = bit-identical to the reference (default), fast = rcp/ex2.approx")
    ap.add_argument("--no-fast-math", action="store_true", help="skip the informational fast-math timing")
    ap.add_argument("--no-check", action="store_true", help="N > 1: skip the N-rank == 1-rank gradient self-check")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-iteration", action="store_true", help="skip the informational whole-iteration timing")
    ap.add_argument("--no-settle", action="store_true", help="profiler runs: only W untimed steps, not a pass over the camera ring")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3)

    ref_mod = None
    if args.impl == "reference":
        if rank != 0:
            return 0  # rank 0 alone runs the reference arm
        world = 1
        try:
            from oracle import build_ref
            if build_ref.up_to_date():
                ref_mod = build_ref.import_reference()
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write(f"[bench] reference extension not loadable: {ex}\n")
        import torch
        if ref_mod is None or not torch.cuda.is_available():
            # no GPU build of the reference: time the CPU port of its algorithm instead
            value, cores, dt, nv = cpu_oracle_step()
            line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
                    "ms_per_step": dt * 1e3 / nv, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f32", "data": "synthetic", "impl": "reference",
                    "config": {"workload": "c2: 1.0M surfels, SH3, 1920x1080, 1 view (CPU port of the reference algorithm)"},
                    "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                     "sample": f"{nv} views forward+backward, full c2 size, {dt:.1f} s"},
                    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
            print(json.dumps(line))
            return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the product path)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist_on = world > 1
    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=device)

    from g4splat_b200.view_parallel import ViewShardedGradSync   # plain torch until told otherwise: loads no native library
    if args.impl == "reference":
        # nothing of this repository's native code may be mapped in this arm: the statistics pass of the step
        # loop runs as the reference's own torch expressions (use_native=False below)
        mod = ref_mod
        lib = None
    else:
        from g4splat_b200 import _lib
        import g4splat_b200.diff_surfel_rasterization as mod
        lib = _lib.load()
        mod.set_fast_math(args.math == "fast")

    wl = Workload(device, args.views, rank, world, args.config)
    sync = ViewShardedGradSync(wl.params, transport=args.transport if dist_on else "nccl", use_native=lib is not None)
    sync.bind(mod)  # B200 operator: kernel-side accumulation into the flat buffer; no-op for the reference
    P, N = wl.P, wl.W * wl.H
    T = ((wl.W + 15) // 16) * ((wl.H + 15) // 16)

    # ---- device-resident throughput ("value") -------------------------------------------------
    # Untimed steps: at least W, and at least one pass over the whole camera ring, so that every buffer size
    # the 64 views need (the instance capacity is sticky and grows with the largest view seen) has been through
    # the caching allocator before the clock starts -- a first-time cudaMalloc of a 100-200 MB block inside the
    # timed region costs 10-100 ms on these hosts (profiles/r01u_bench_stability.md).
    views_total = wl.views_per_step_total()
    settle = warmup if args.no_settle else max(warmup, -(-wl.ring // views_total))
    run_steps(wl, mod, sync, settle, 0, e2e=False)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # one more untimed step while nvidia-smi starts up (its NVML initialisation takes a driver lock for
    # tens of milliseconds and would otherwise land inside the timed region)
    run_steps(wl, mod, sync, 1, settle, e2e=False)
    first_timed = settle + 1

    def read_stages():
        import ctypes as C
        n_st = lib.g4s_profile_num_stages()
        buf, cnt = (C.c_float * n_st)(), (C.c_int64 * n_st)()
        lib.g4s_profile_read(buf, cnt, n_st)
        names = [lib.g4s_profile_stage_name(i).decode() for i in range(n_st)]
        return names, {nm: float(buf[i]) for i, nm in enumerate(names)}, {nm: int(cnt[i]) for i, nm in enumerate(names)}

    # Per-stage CUDA-event brackets (g4s_profile_*): every bracket is two event records between kernels, and bracketing
    # all eight stages costs 2.8 % of the throughput (measured: 641.6 vs 659.7 M Gaussians/s).  So: ALL stages are
    # bracketed over one untimed pass over the camera ring here (the `stages` block), and inside the timed region only the dominant kernel
    # -- the one `roofline` is about -- is bracketed (G4S_BENCH_STAGE_EVENTS=all brackets everything there as before,
    # =none nothing).
    stage_ms, stage_n, dom, dom_ms_timed, dom_n_timed = {}, {}, None, None, 0
    events_mode = os.environ.get("G4S_BENCH_STAGE_EVENTS", "dominant")
    if lib is not None:
        lib.g4s_profile_select(0xffffffff)
        lib.g4s_profile_enable(1)
        stage_steps = max(2, min(args.steps, -(-wl.ring // views_total)))   # one pass over the camera ring
        run_steps(wl, mod, sync, stage_steps, first_timed, e2e=False)    # the first steps of the timed window, once more below
        names, stage_ms, stage_n = read_stages()
        lib.g4s_profile_enable(0)
        dom = max((k for k in stage_ms if stage_ms[k] > 0), key=lambda k: stage_ms[k], default=None)
        if events_mode == "all":
            lib.g4s_profile_enable(1)
        elif events_mode == "dominant" and dom is not None:
            lib.g4s_profile_select(1 << names.index(dom))
            lib.g4s_profile_enable(1)
    launches0 = lib.g4s_launch_count() if lib is not None else 0
    ms, _ = time_region(lambda: run_steps(wl, mod, sync, args.steps, first_timed, e2e=False), device, dist_on)
    clk = clocks.stop() if rank == 0 else None
    launches = (lib.g4s_launch_count() - launches0) if lib is not None else None
    if lib is not None:
        if events_mode in ("all", "dominant") and dom is not None:
            _, timed_ms, timed_n = read_stages()
            dom_ms_timed, dom_n_timed = timed_ms[dom], timed_n[dom]
            if events_mode == "all":
                stage_ms, stage_n = timed_ms, timed_n
        lib.g4s_profile_enable(0)
        lib.g4s_profile_select(0xffffffff)
    value = P * views_total * args.steps / (ms * 1e-3)

    # ---- end to end from host buffers ("e2e") ---------------------------------------------------
    run_steps(wl, mod, sync, max(2, settle), 1000, e2e=True)
    ms_e2e, _ = time_region(lambda: run_steps(wl, mod, sync, args.steps, 1000 + max(2, settle), e2e=True), device, dist_on)
    e2e_value = P * views_total * args.steps / (ms_e2e * 1e-3)
    h2d = len(wl.view_ids(0)) * (3 * N)   # one uint8 image per view of this rank
    d2h = 4                             # the scalar loss

    # ---- the opt-in fast forward, same loop (informational; the headline `value` above is the exact mode) ---------
    fast = None
    if lib is not None and args.math == "exact" and not args.no_fast_math:
        mod.set_fast_math(True)
        run_steps(wl, mod, sync, 2, 2000, e2e=False)
        k_fast = max(3, args.steps // 2)
        ms_fast, _ = time_region(lambda: run_steps(wl, mod, sync, k_fast, 2002, e2e=False), device, dist_on)
        mod.set_fast_math(False)
        fast = {"value": P * views_total * k_fast / (ms_fast * 1e-3), "unit": UNIT, "ms_per_step": ms_fast / k_fast, "steps": k_fast,
                "what": "set_fast_math(True): rcp.approx / ex2.approx in the forward blend; 1e-4 parity with <= 2e-5 of the image "
                        "elements flipping a threshold (tests/test_parity_gpu.py::test_fast_math_forward_stays_within_north_star_tolerance)"}

    train_it = None
    if world == 1 and not args.no_train_iteration:
        sync.zero()
        if hasattr(mod, "set_gradient_sink"):
            mod.set_gradient_sink(None)   # the multi-view gradient sink belongs to the step loops above
        fused = args.impl != "reference"
        ms_it = time_train_iteration(wl, mod, fused, 16, device)
        train_it = {"ms_per_view": ms_it, "gaussians_per_s": P / (ms_it * 1e-3),
                    "what": "one view of train_with_refine_depth.py:378-399 from the optimiser's raw leaves: activations, rasterizer, "
                            "render() post-processing, 0.8 L1 + 0.2 D-SSIM + normal + distortion loss, backward (SURVEY.md 8f rows 1, 2, 4)",
                    "arm": "this repository's fused kernels for every row" if fused else
                           "the reference rasterizer with the reference's torch operator sequences around it"}

    op_timing = None
    if world == 1 and not args.no_train_iteration:
        op_timing = time_operator(wl, mod, device)
        sync.zero()                       # re-attaches the flat gradient buffer that time_operator detached

    check = None
    if dist_on and lib is not None and not args.no_check:
        check = allreduce_self_check(wl, mod, sync, device)

    pair_stats = None
    if lib is not None and rank == 0:
        import g4splat_b200.diff_surfel_rasterization as op_mod
        acc = {}
        vids = list(wl.view_ids(3))
        for vid in vids[:8]:             # untimed: work counters of (up to eight) views of one step
            color = wl.rasterize(mod, vid)[0]
            for k, v in op_mod.debug_pair_stats(color).items():
                acc[k] = max(acc.get(k, 0), v) if k == "longest_tile_list" else acc.get(k, 0) + v
            del color
        pair_stats = {k: (v if k == "longest_tile_list" else v / len(vids[:8])) for k, v in acc.items()}
        sync.zero()

    if rank != 0:
        if dist_on:
            sync.close()
            dist.destroy_process_group()
        return 0

    peak, peak_src = load_peaks()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world if args.impl != "reference" else args.gpus,
            "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": wl.doc["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl.doc['name']}, " +
                                   (f"{views_total} views/step sharded over {world} GPU(s)" if wl.doc["step_views"] is not None else
                                    f"{args.views} views/GPU/step") + f" from a ring of {wl.ring} cameras, fwd+bwd" +
                                   (f" + gradient sum over ranks ({sync.transport})" if world > 1 else ""),
                       "P": P, "width": wl.W, "height": wl.H, "views_per_step": views_total,
                       "views_per_gpu_per_step": len(wl.view_ids(0)),
                       "l2_policy": f"inputs larger than L2: {232 * P // 1_000_000} MB of parameters + {192 * P // 1_000_000} MB of SH gradients per view, a different camera every view",
                       "untimed_steps_before_timing": settle + 1,
                       "math": args.math if args.impl != "reference" else "reference",
                       "view_pipeline": bool(PIPELINE and args.impl != "reference"),
                       "transport": sync.transport if world > 1 else None,
                       "parallelism": f"view-sharded dp{world}"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clk}
    if train_it is not None:
        line["train_iteration"] = train_it
    if op_timing is not None:
        line["operator_timing"] = op_timing
    if fast is not None:
        line["fast_math"] = fast
    if check is not None:
        line["allreduce_check"] = check
    if world > 1:
        line["grad_sum"] = {"transport": sync.transport, "bytes_per_rank_per_step": sync.bytes_per_step}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["gpu_launches"] = None
        line["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                                "sample": "the unmodified reference CUDA extension (oracle/_ref, sm_100a) on the same B200; "
                                          "the reference has no CPU implementation of this path"}
    else:
        import g4splat_b200.diff_surfel_rasterization as op
        V, R = op.last_counts["visible"], op.last_counts["num_rendered"]
        Kc, M = 16, 16
        if dom is not None:
            # the dominant kernel's mean launch duration over the TIMED region (its brackets stayed on there)
            dom_ms = dom_ms_timed if dom_ms_timed is not None and dom_ms_timed > 0 else stage_ms[dom]
            ab = algorithmic_bytes(dom, P, V, R, N, T, Kc, M)
            achieved = ab / (dom_ms * 1e-3) / 1e9
            # traffic: dram bytes per launch from the committed `ncu --set full` capture of this workload (c2 only),
            # labelled with the capture it came from -- it is not re-measured by this run
            traffic, traffic_src = None, None
            tp = ROOT / "profiles" / "traffic.json"
            if tp.exists() and args.config == "c2":
                tj = json.loads(tp.read_text())
                traffic, traffic_src = tj.get(dom), tj.get("_source")
            line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                                "peak_source": peak_src,
                                "algorithmic_bytes_per_launch": ab, "kernel_ms": dom_ms,
                                "kernel_ms_source": ("CUDA events around every launch of this kernel inside the timed region (%d launches)" % dom_n_timed)
                                                    if dom_ms_timed is not None and dom_ms_timed > 0 else "untimed stage pass",
                                "note": "the blend kernels are bound by the FP32 pipe and shared-memory bandwidth, not by HBM (DESIGN.md); see stages/path"}
        per_view_ms = sum(v for v in stage_ms.values() if v > 0)
        line["stages_source"] = ("CUDA events around every stage inside the timed region" if events_mode == "all" else
                                 "CUDA events around every stage over an untimed pass over the camera ring right before the timed region (bracketing all "
                                 "eight stages inside it costs 2.8 % of the throughput; the dominant kernel alone stays bracketed there)")
        line["stages"] = {k: {"ms": stage_ms[k], "launches": stage_n[k],
                              "alg_GBps": (algorithmic_bytes(k, P, V, R, N, T, Kc, M) / (stage_ms[k] * 1e-3) / 1e9) if stage_ms[k] > 0 else None}
                          for k in stage_ms}
        pb = path_bytes(P, V, R, N, Kc, M)
        line["path"] = {"alg_bytes_per_view": pb, "kernel_ms_per_view": per_view_ms, "visible": V, "num_rendered": R,
                        "hbm_frac_of_kernel_time": (pb / (per_view_ms * 1e-3) / 1e9 / peak) if per_view_ms > 0 else None,
                        "hbm_frac_of_step_time": pb * len(wl.view_ids(0)) / (ms / args.steps * 1e-3) / 1e9 / peak}
        if pair_stats is not None and stage_ms.get("blend_fwd", 0) > 0 and stage_ms.get("blend_bwd", 0) > 0:
            # secondary roofline: algorithmic ~60 flop per blended pair forward, ~200 backward,
            # against the fp32 SIMT peak 148 SMs x 128 lanes x 2 flop x SM clock under load
            mhz = (clk or {}).get("sm_mhz") or 1965.0
            fp32_peak = 148 * 128 * 2 * mhz * 1e6
            pb_, pe_ = pair_stats["pairs_blended"], pair_stats["pair_evals_bwd"]
            fwd_s, bwd_s = stage_ms["blend_fwd"] * 1e-3, stage_ms["blend_bwd"] * 1e-3
            line["compute_roofline"] = {
                "per_view": pair_stats, "fp32_peak_TFLOPs": fp32_peak / 1e12,
                "blend_fwd": {"Gpairs_per_s": pb_ / fwd_s / 1e9, "alg_flop_per_pair": 60,
                              "frac_of_fp32_peak": 60 * pb_ / fwd_s / fp32_peak},
                "blend_bwd": {"Gpairs_per_s": pb_ / bwd_s / 1e9, "alg_flop_per_pair": 200,
                              "frac_of_fp32_peak": 200 * pb_ / bwd_s / fp32_peak,
                              "lane_utilisation": pb_ / pe_ if pe_ else None},
                "lane_slots_skipped_by_culling": 1.0 - pe_ / pair_stats["pair_slots"] if pair_stats["pair_slots"] else None}
        if not args.no_cpu_baseline and world == 1 and args.config == "c2":
            v, cores, dt, nv = cpu_oracle_step()
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{nv} views forward+backward at full c2 size, {dt:.1f} s"}
    sys.stdout.write("\n" + json.dumps(line) + "\n")   # own line even when a library wrote to stdout before (NCCL banner)
    sys.stdout.flush()
    dump_trace()
    if os.environ.get("G4S_HOST_TRACE") == "1" and args.impl != "reference":
        import g4splat_b200.diff_surfel_rasterization as op_mod
        sys.stderr.write("host trace (all calls of this process): " + json.dumps(op_mod.host_trace_summary()) + "\n")
    if dist_on:
        sync.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
