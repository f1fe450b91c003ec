"""Times the caller-side rows of SURVEY.md 8(f) on the GPU, fused kernels against the reference's
operator sequences (the torch restatements under oracle/ run on CUDA tensors = the kernels the
reference launches):
  * photometric loss (L1 + SSIM, forward + backward) at 3x1080x1920          -- loss_utils.py, train:382-383
  * compute_mip_filter for 1.0 M Gaussians and 64 cameras                    -- gaussian_model.py:388-434
  * normal2curv, compute_depth_order_loss at 1080p                           -- matcha/dm_utils/rendering.py, dm_regularization/depth.py
  * densify_and_prune on 1.0 M Gaussians with Adam moments                   -- gaussian_model.py:528-647

    python tests/tools/bench_trainer_ops.py [--iters 30]      # prints one JSON line per row
"""
import argparse
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_loss(iters):
    import make_golden_loss as MG
    from g4splat_b200.loss_utils import photometric_loss
    from oracle import loss_oracle as LO
    img_np, gt_np = MG.make_images(3, 1080, 1920, 7, "noisy")
    img, gt = torch.tensor(img_np, device="cuda"), torch.tensor(gt_np, device="cuda")
    grads = {}

    def fused():
        x = img.clone().requires_grad_(True)
        loss, _ = photometric_loss(x, gt, 0.2)
        loss.backward()
        grads["fused"] = x.grad

    def reference():
        x = img.clone().requires_grad_(True)
        loss = 0.8 * LO.l1_loss(x, gt) + 0.2 * (1.0 - LO.ssim(x, gt))
        loss.backward()
        grads["ref"] = x.grad

    ms_f, ms_r = timed(fused, iters), timed(reference, iters)
    err = float((grads["fused"] - grads["ref"]).abs().max() / grads["ref"].abs().max())
    n = 3 * 1080 * 1920
    alg = n * 4 * (2 + 3) + n * 4 * (3 + 2 + 1)     # fwd: read x, y, write 3 maps; bwd: read 3 maps, x, y, write grad
    print(json.dumps({"row": "photometric loss (L1 + SSIM) forward + backward, 3x1080x1920", "fused_ms": ms_f,
                      "reference_torch_ops_ms": ms_r, "speedup": ms_r / ms_f, "alg_bytes": alg,
                      "fused_GBps": alg / (ms_f * 1e-3) / 1e9, "max_rel_grad_diff_vs_torch_fp32": err, "iters": iters}))


def bench_mip(iters):
    import make_golden_mip as MG
    from g4splat_b200.gaussian_model import compute_mip_filter
    from g4splat_b200 import synthetic as S
    P, C = 1_000_000, 64
    xyz = torch.tensor(S.make_scene(P, 2)["means3D"], device="cuda")
    _, cams = MG.make_case(8, C, 1920, 1080, 3, 1.0)
    out = {}

    def fused():
        out["fused"] = compute_mip_filter(xyz, cams)

    def reference():                       # gaussian_model.py:388-434 on CUDA tensors
        distance = torch.ones((P,), device="cuda") * 100000.0
        valid_points = torch.zeros((P,), device="cuda", dtype=torch.bool)
        focal_length = 0.0
        for camera in cams:
            R = torch.tensor(camera.R, device="cuda", dtype=torch.float32)
            T = torch.tensor(camera.T, device="cuda", dtype=torch.float32)
            xyz_cam = xyz @ R + T[None, :]
            xyz_to_cam = torch.norm(xyz_cam, dim=1)  # noqa: F841  (computed and unused in the reference too)
            valid_depth = xyz_cam[:, 2] > 0.2
            x, y, z = xyz_cam[:, 0], xyz_cam[:, 1], xyz_cam[:, 2]
            z = torch.clamp(z, min=0.001)
            x = x / z * camera.focal_x + camera.image_width / 2.0
            y = y / z * camera.focal_y + camera.image_height / 2.0
            in_screen = torch.logical_and(torch.logical_and(x >= -0.15 * camera.image_width, x <= camera.image_width * 1.15),
                                          torch.logical_and(y >= -0.15 * camera.image_height, y <= 1.15 * camera.image_height))
            valid = torch.logical_and(valid_depth, in_screen)
            distance[valid] = torch.min(distance[valid], z[valid])
            valid_points = torch.logical_or(valid_points, valid)
            if focal_length < camera.focal_x:
                focal_length = camera.focal_x
        distance[~valid_points] = distance[valid_points].max()
        out["ref"] = (distance / focal_length * (0.2 ** 0.5))[..., None]

    ms_f, ms_r = timed(fused, iters), timed(reference, max(3, iters // 5))
    bad = float(((out["fused"] - out["ref"]).abs() > 2e-6 * out["ref"].abs()).float().mean())
    alg = P * (12 + 4) + C * 80
    print(json.dumps({"row": f"compute_mip_filter, {P} Gaussians x {C} cameras", "fused_ms": ms_f, "reference_torch_ops_ms": ms_r,
                      "speedup": ms_r / ms_f, "alg_bytes": alg, "fused_GBps": alg / (ms_f * 1e-3) / 1e9,
                      "fraction_outside_2e-6": bad, "iters": iters}))


def bench_activations(iters):
    """One training-shaped view at c2 size: activate with the reference's torch operators and call the
    operator, against the raw-parameter operator (activations inside the projection kernels)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import helpers as Hh
    import test_activations as TA
    import g4splat_b200.diff_surfel_rasterization as op
    from oracle import activations_oracle as AO
    case = Hh.room_case("act_bench", P=1_000_000, W=1920, H=1080, seed=2, cams=64, cam_index=5)
    raw_np, mip_np = TA.raw_from_case(case, True, seed=4)
    raw = {k: torch.tensor(v, device="cuda", requires_grad=True) for k, v in raw_np.items()}
    mip = torch.tensor(mip_np, device="cuda")
    settings = Hh.make_settings(op, case, "cuda")
    gc, go = (torch.tensor(a, device="cuda") for a in case.upstream())

    def step(fused):
        for v in raw.values():
            v.grad = None
        means2D = torch.zeros_like(raw["_xyz"], requires_grad=True)
        if fused:
            color, radii, allmap = op.rasterize_gaussian_model(raw["_xyz"], means2D, raw["_features_dc"], raw["_features_rest"],
                                                               raw["_opacity"], raw["_scaling"], raw["_rotation"], mip, settings)
        else:
            act = AO.activate(raw, mip)
            color, radii, allmap = op.GaussianRasterizer(raster_settings=settings)(
                means3D=act["means3D"], means2D=means2D, opacities=act["opacities"], shs=act["shs"],
                scales=act["scales"], rotations=act["rotations"])
        torch.autograd.backward([color, allmap], [gc, go])

    ms_f, ms_r = timed(lambda: step(True), iters), timed(lambda: step(False), iters)
    print(json.dumps({"row": "one view forward + backward at c2 size (1.0 M surfels, 1920x1080, mip filter on), from raw leaves",
                      "fused_ms": ms_f, "reference_torch_ops_ms": ms_r, "saved_ms_per_view": ms_r - ms_f,
                      "speedup": ms_r / ms_f, "iters": iters,
                      "note": "both arms use the B200 rasterizer; the difference is the activations (~10 torch kernels + autograd, "
                              "incl. the 192 B/Gaussian cat of the SH tensors) against in-register activations"}))


def bench_regularizers(iters):
    """normal2curv and the depth-order loss at 1080p, forward + backward, against the reference's torch sequences
    (oracle/regularizers_oracle.py restates matcha/dm_utils/rendering.py:392-406 and matcha/dm_regularization/depth.py:142-222)."""
    from g4splat_b200 import regularization as R
    from oracle import regularizers_oracle as RO
    H, W = 1080, 1920
    gen = torch.Generator(device="cuda").manual_seed(1)
    normal = torch.randn(3, H, W, device="cuda", generator=gen)
    mask = torch.ones(1, H, W, device="cuda")
    depth = 1.0 + 4.0 * torch.rand(1, H, W, device="cuda", generator=gen)
    prior = depth * 0.8 + 0.3 * torch.randn(1, H, W, device="cuda", generator=gen)
    max_shift = round(0.05 * max(H, W))

    def curv(fn):
        def run():
            n = normal.clone().requires_grad_(True)
            fn(n, mask).mean().backward()
        return run

    def order(fused):
        def run():
            d = depth.clone().requires_grad_(True)
            if fused:
                loss = R.compute_depth_order_loss(d, prior, scene_extent=3.3, log_space=True)
            else:
                shifts = torch.randint(-max_shift, max_shift + 1, (H * W, 2), device="cuda")
                loss = RO.depth_order_loss(d, prior, shifts, 3.3, True, True, 20., "mean")
            loss.backward()
        return run

    for row, f, r in (("normal2curv forward + backward, 3x1080x1920", curv(R.normal2curv), curv(RO.normal2curv)),
                      ("compute_depth_order_loss forward + backward, 1080x1920 (incl. the randint draw)", order(True), order(False))):
        ms_f, ms_r = timed(f, iters), timed(r, iters)
        print(json.dumps({"row": row, "fused_ms": ms_f, "reference_torch_ops_ms": ms_r, "speedup": ms_r / ms_f, "iters": iters}))


def bench_densify(iters):
    """densify_and_prune on 1.0 M Gaussians with Adam moments, against the reference's sequence (oracle/densify_oracle.py
    restates scene/gaussian_model.py:528-647 on plain tensors: the same torch kernels the reference launches)."""
    import types
    from g4splat_b200.gaussian_model import densify_and_prune
    from g4splat_b200 import synthetic as S
    from oracle import densify_oracle as DO
    P = 1_000_000
    sc = S.make_scene(P, 2)
    rng = np.random.default_rng(3)
    op = np.clip(sc["opacities"], 1e-4, 1 - 1e-4)
    raw = {"xyz": sc["means3D"], "f_dc": sc["shs"][:, :1], "f_rest": sc["shs"][:, 1:], "opacity": np.log(op / (1 - op)),
           "scaling": np.log(sc["scales"]), "rotation": sc["rotations"]}
    raw = {k: torch.tensor(np.ascontiguousarray(v, dtype=np.float32), device="cuda") for k, v in raw.items()}
    accum = torch.tensor(np.float32(np.abs(rng.normal(scale=0.0004, size=(P, 1))) * 5), device="cuda")
    denom = torch.full((P, 1), 5.0, device="cuda")
    attr = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling", "rotation": "_rotation"}
    sizes = {}

    def fused():
        model = types.SimpleNamespace(percent_dense=0.01, xyz_gradient_accum=accum.clone(), denom=denom.clone(),
                                      max_radii2D=torch.zeros(P, device="cuda"))
        groups = []
        for k, a in attr.items():
            p_ = torch.nn.Parameter(raw[k].clone())
            setattr(model, a, p_)
            groups.append({"params": [p_], "lr": 1e-3, "name": k})
        model.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        for g in groups:
            model.optimizer.state[g["params"][0]] = {"step": torch.tensor(3.0), "exp_avg": torch.zeros_like(g["params"][0]),
                                                     "exp_avg_sq": torch.zeros_like(g["params"][0])}
        sizes["fused"] = densify_and_prune(model, 0.0002, 0.05, 5.0, 20)

    def reference():
        params = {k: v.clone() for k, v in raw.items()}
        moments = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in raw.items()}
        grads = accum / denom
        sel = (grads.squeeze() >= 0.0002) & (torch.exp(params["scaling"]).max(dim=1).values > 0.01 * 5.0)
        stds = torch.exp(params["scaling"][sel]).repeat(2, 1)
        stds = torch.cat([stds, 0 * torch.ones_like(stds[:, :1])], dim=-1)
        samples = torch.normal(mean=torch.zeros_like(stds), std=stds)
        p_, _ = DO.densify_and_prune(params, moments, accum.clone(), denom.clone(), 0.01, 0.0002, 0.05, 5.0, 20, samples)
        sizes["ref"] = int(p_["xyz"].shape[0])

    # both arms pay the same clones of the inputs (six parameters + moments); the difference is the densification itself
    ms_f, ms_r = timed(fused, max(3, iters // 3)), timed(reference, max(3, iters // 3))
    print(json.dumps({"row": f"densify_and_prune, {P} Gaussians with Adam moments ({sizes.get('fused')} afterwards; setup clones included in both arms)",
                      "fused_ms": ms_f, "reference_torch_ops_ms": ms_r, "speedup": ms_r / ms_f, "P_after_fused": sizes.get("fused"),
                      "P_after_reference": sizes.get("ref"), "iters": max(3, iters // 3)}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    rows = {"loss": bench_loss, "mip": bench_mip, "activations": bench_activations, "regularizers": bench_regularizers,
            "densify": bench_densify}
    for name, fn in rows.items():
        if not args.only or name in args.only.split(","):
            fn(args.iters)


if __name__ == "__main__":
    main()
