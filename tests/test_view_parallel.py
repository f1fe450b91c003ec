"""View-sharded data parallelism on CPU: world_size 2 over gloo.  The per-view gradients come
from the CPU oracle (test infrastructure), the thing under test is the sharding + packing +
all-reduce logic of g4splat_b200.view_parallel: summed N-rank gradients and densification
statistics must equal a single-process loop over the same views."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

P, W, H, VIEWS = 800, 48, 32, 5


def _params():
    from g4splat_b200 import synthetic as S
    sc = S.make_scene(P, 31)
    t = lambda a: torch.from_numpy(a.copy()).requires_grad_(True)
    return {"xyz": t(sc["means3D"]), "features_dc": t(sc["shs"][:, :1]), "features_rest": t(sc["shs"][:, 1:]),
            "opacity": t(sc["opacities"]), "scaling": t(sc["scales"]), "rotation": t(sc["rotations"])}


def _render_one_factory(params, oracle):
    """(loss, viewspace_points, radii) for one camera, with gradients injected from the oracle."""
    import helpers as Hh
    from g4splat_b200 import synthetic as S

    class OracleRaster(torch.autograd.Function):
        @staticmethod
        def forward(ctx, xyz, means2D, fdc, frest, opacity, scaling, rotation, cam):
            scene = dict(means3D=xyz.detach().numpy(), shs=torch.cat([fdc, frest], 1).detach().numpy(),
                         opacities=opacity.detach().numpy(), scales=scaling.detach().numpy(),
                         rotations=rotation.detach().numpy())
            ctx.out = Hh.run_oracle(oracle, Hh.Case("v", scene, cam, grad_seed=7))
            radii = torch.from_numpy(ctx.out["radii"].copy())
            ctx.mark_non_differentiable(radii)
            return torch.from_numpy(ctx.out["color"].copy()), torch.from_numpy(ctx.out["allmap"].copy()), radii

        @staticmethod
        def backward(ctx, g_color, g_allmap, g_radii):
            o = ctx.out  # the oracle already applied the fixed upstream gradients of the case
            f = lambda k: torch.from_numpy(o[k].copy())
            sh = f("dL_dsh")
            return (f("dL_dmeans3D"), f("dL_dmeans2D"), sh[:, :1], sh[:, 1:], f("dL_dopacity"), f("dL_dscales"),
                    f("dL_drotations"), None)

    def render_one(cam):
        means2D = torch.zeros(P, 3, requires_grad=True)
        color, allmap, radii = OracleRaster.apply(params["xyz"], means2D, params["features_dc"], params["features_rest"],
                                                  params["opacity"], params["scaling"], params["rotation"], cam)
        gc, go = S.make_upstream_grads(cam.W, cam.H, 7)
        loss = (color * torch.from_numpy(gc)).sum() + (allmap * torch.from_numpy(go)).sum()
        return loss, means2D, radii

    return render_one


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from g4splat_b200 import synthetic as S
        from g4splat_b200.view_parallel import ViewShardedGradSync, render_views_sharded
        from oracle.oracle import Oracle
        torch.set_num_threads(1)
        oracle = Oracle("f32")
        oracle.set_threads(1)
        params = _params()
        sync = ViewShardedGradSync(params)
        views = S.make_cameras(VIEWS, W, H)
        render_views_sharded(_render_one_factory(params, oracle), views, sync, rank, world)
        torch.save({"grads": {k: v.grad.clone() for k, v in params.items()}, "flat": sync.flat.clone(),
                    "max_radii": sync.max_radii.clone()}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_gradients_equal_single_process_loop(tmp_path):
    from g4splat_b200 import synthetic as S
    from g4splat_b200.view_parallel import ViewShardedGradSync, render_views_sharded
    from oracle.oracle import Oracle
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    # single-process reference: world_size 1 runs every view
    oracle = Oracle("f32")
    oracle.set_threads(1)
    params = _params()
    sync = ViewShardedGradSync(params)
    render_views_sharded(_render_one_factory(params, oracle), S.make_cameras(VIEWS, W, H), sync, 0, 1)
    for k, p in params.items():
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k          # replicas stay identical
        ref = p.grad
        assert ref.abs().max() > 0, k
        assert torch.allclose(r0["grads"][k], ref, rtol=1e-4, atol=1e-6 * float(ref.abs().max())), k
    assert torch.equal(r0["max_radii"], sync.max_radii)
    assert torch.allclose(r0["flat"][-2 * P:], sync.flat[-2 * P:], rtol=1e-4, atol=1e-9)
    assert float(sync.denom.sum()) > 0 and float(sync.xyz_gradient_accum.sum()) > 0
    assert sync.width == 60 and sync.bytes_per_step == P * 60 * 4 + P * 4
