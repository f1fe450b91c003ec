"""Host-side logic of the operator shim and the view-sharding helpers (no GPU needed)."""
import numpy as np
import pytest
import torch


def test_one_of_exceptions_match_the_reference_text():
    """RAST/diff_surfel_rasterization/__init__.py:192-196 -- raised before any device work."""
    import g4splat_b200.diff_surfel_rasterization as op
    rs = op.GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                          torch.zeros(3), False, False)
    r = op.GaussianRasterizer(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(x, x, torch.zeros(4, 1), shs=None, colors_precomp=None, scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="excatly one of either SHs or precomputed colors"):
        r(x, x, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), colors_precomp=torch.zeros(4, 3),
          scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(x, x, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), scales=torch.zeros(4, 2))
    with pytest.raises(Exception, match="exactly one of either scale/rotation pair or precomputed 3D covariance"):
        r(x, x, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 9))


def test_settings_fields_match_the_reference():
    import g4splat_b200.diff_surfel_rasterization as op
    assert op.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "sh_degree", "campos", "prefiltered", "debug")


def test_cpu_tensors_are_rejected_loudly():
    """No CPU fallback: a CPU tensor must raise, not silently compute somewhere else."""
    import g4splat_b200.diff_surfel_rasterization as op
    rs = op.GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                          torch.zeros(3), False, False)
    r = op.GaussianRasterizer(rs)
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        r(x, x, torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), scales=torch.zeros(4, 2), rotations=torch.zeros(4, 4))


def test_install_registers_the_reference_import_name():
    import sys
    import g4splat_b200
    mod = g4splat_b200.install()
    import diff_surfel_rasterization as d
    assert d is mod and hasattr(d, "GaussianRasterizer") and hasattr(d, "rasterize_gaussians")
    del sys.modules["diff_surfel_rasterization"]


def test_root_level_import_name_shim():
    """`import diff_surfel_rasterization` with the repo root on sys.path is the B200 operator."""
    import importlib
    import sys
    sys.modules.pop("diff_surfel_rasterization", None)
    mod = importlib.import_module("diff_surfel_rasterization")
    import g4splat_b200.diff_surfel_rasterization as op
    assert mod.GaussianRasterizer is op.GaussianRasterizer
    assert mod.GaussianRasterizationSettings is op.GaussianRasterizationSettings
    assert mod.rasterize_gaussians is op.rasterize_gaussians
    sys.modules.pop("diff_surfel_rasterization", None)


def test_capacity_policy():
    from g4splat_b200.diff_surfel_rasterization import _CapacityPolicy
    p = _CapacityPolicy()
    assert p.guess(0, 1_000_000) == 6 * (1 << 20)          # first call: 6 P rounded up to a bucket
    assert p.guess(0, 10) == 1 << 20
    p.observe(0, 3_000_000)
    g = p.guess(0, 1_000_000)
    assert g >= 4_500_000 + 65536 and g == p.bucket(g)      # 1.5x the count, bucketed
    p.observe(0, 2_900_000)
    assert p.guess(0, 1_000_000) == g                        # sticky: similar views reuse the same block size
    p.observe(0, 9_000_000)
    assert p.guess(0, 1_000_000) > g                         # grows
    assert p.guess(1, 10) == 1 << 20                         # per device
    for n in (1, 65536, 65537, 1_000_000, 5_000_001):
        assert p.bucket(n) >= n and p.bucket(n) <= max(1 << 16, int(n * 1.26))


def bitonic_ascending(keys):
    """Python mirror of bitonic_sort_ascending (g4splat_b200/csrc/binning.cu): same index math."""
    keys = list(keys)
    n = len(keys)
    m, log_m = 1, 0
    while m < n:
        m <<= 1
        log_m += 1
    half = m >> 1
    for lk in range(1, log_m + 1):
        k, hk = 1 << lk, (1 << lk) >> 1
        for i in range(half):
            within = i & (hk - 1)
            blk_base = (i >> (lk - 1)) << lk
            lo, hi = blk_base + within, blk_base + k - 1 - within
            if hi < n and keys[lo] > keys[hi]:
                keys[lo], keys[hi] = keys[hi], keys[lo]
        for lj in range(lk - 2, -1, -1):
            j = 1 << lj
            for i in range(half):
                lo = ((i >> lj) << (lj + 1)) | (i & (j - 1))
                hi = lo + j
                if hi < n and keys[lo] > keys[hi]:
                    keys[lo], keys[hi] = keys[hi], keys[lo]
    return keys


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 32, 33, 100, 257, 1000])
def test_bitonic_network_sorts_any_length(n):
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 1 << 62, size=n).tolist()
    assert bitonic_ascending(keys) == sorted(keys)


def test_shard_views_partition():
    from g4splat_b200.view_parallel import shard_views
    assert [len(shard_views(50, 4, r)) for r in range(4)] == [13, 13, 12, 12]
    for n, w in [(64, 8), (5, 2), (3, 4), (0, 2), (7, 7)]:
        got = [i for r in range(w) for i in shard_views(n, w, r)]
        assert got == list(range(n))


def test_synthetic_scene_is_seeded_and_well_formed():
    from g4splat_b200 import synthetic as S
    a, b = S.make_scene(2000, 3), S.make_scene(2000, 3)
    for k in a:
        assert a[k].dtype == np.float32 and np.array_equal(a[k], b[k])
    assert a["shs"].shape == (2000, 16, 3) and a["rotations"].shape == (2000, 4)
    assert np.allclose(np.linalg.norm(a["rotations"], axis=1), 1, atol=1e-5)
    assert (a["opacities"] > 0).all() and (a["opacities"] < 1).all()
    cam = S.make_cameras(3, 320, 200)[1]
    # view matrix is a rigid transform, camera centre maps to the origin
    R = cam.viewmatrix.T[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)
    assert np.allclose(cam.viewmatrix.T @ np.append(cam.campos, 1.0), [0, 0, 0, 1], atol=1e-4)


# ---- caller-side mirrors: argument handling that needs no GPU ----------------------
def test_caller_side_mirrors_reject_cpu_tensors_and_bad_shapes():
    import g4splat_b200.diff_surfel_rasterization as op
    from g4splat_b200.gaussian_model import compute_mip_filter
    from g4splat_b200.loss_utils import photometric_loss, ssim
    from g4splat_b200.surface import surface_attributes
    img = torch.rand(3, 24, 32)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        photometric_loss(img, img, 0.2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ssim(img, img)
    with pytest.raises(NotImplementedError):
        ssim(img, img, window_size=7)          # checked before the device: the reference's only configuration is 11
    with pytest.raises(ValueError, match=r"\(C, H, W\)"):
        photometric_loss(torch.rand(24, 32), torch.rand(24, 32), 0.2)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        surface_attributes(torch.rand(7, 8, 8), torch.eye(4), torch.eye(4), 1.0)
    with pytest.raises(ValueError, match="7, H, W"):
        surface_attributes(torch.rand(6, 8, 8), torch.eye(4), torch.eye(4), 1.0)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        compute_mip_filter(torch.zeros(4, 3), [])
    rs = op.GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 3,
                                          torch.zeros(3), False, False)
    x = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        op.rasterize_gaussian_model(x, x, torch.zeros(4, 1, 3), torch.zeros(4, 15, 3), torch.zeros(4, 1),
                                    torch.zeros(4, 2), torch.zeros(4, 4), None, rs)


def test_mip_filter_camera_records_round_python_scalars_once():
    """The reference multiplies Python floats (double) and lets torch round the result to fp32 once
    (gaussian_model.py:416-420); the host builds the camera records the same way."""
    import types
    from g4splat_b200.gaussian_model import CAMERA_RECORD_FLOATS, camera_records
    cam = types.SimpleNamespace(R=np.eye(3), T=np.array([0.1, 0.2, 0.3]), focal_x=1111.111111, focal_y=999.9,
                                image_width=1237, image_height=821)
    rec = camera_records([cam, cam])
    assert rec.shape == (2, CAMERA_RECORD_FLOATS) and rec.dtype == np.float32
    assert rec[0, 12] == np.float32(1111.111111) and rec[0, 14] == np.float32(1237 / 2.0)
    assert rec[0, 16] == np.float32(-0.15 * 1237) and rec[0, 17] == np.float32(1237 * 1.15)
    assert rec[0, 18] == np.float32(-0.15 * 821) and rec[0, 19] == np.float32(1.15 * 821)
    assert np.array_equal(rec[0, 9:12], np.float32([0.1, 0.2, 0.3]))


def test_loss_window_is_the_reference_gaussian():
    from math import exp
    from g4splat_b200 import loss_utils as LU
    ref = torch.Tensor([exp(-(x - 5) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    ref = ref / ref.sum()
    assert torch.equal(LU.gaussian(), ref) and len(LU._WINDOW_C) == 11
    assert all(float(LU._WINDOW_C[i]) == float(ref[i]) for i in range(11))


def test_capacity_policy_buckets():
    """Sticky, bucketed instance capacity (operator shim): a bucket holds the request, wastes < 25 % above
    2^16, is monotone, and the policy only moves when the request outgrows it or shrinks four-fold."""
    import g4splat_b200.diff_surfel_rasterization as op
    bucket = op._CapacityPolicy.bucket
    prev = 0
    for n in list(range(1, 70000, 997)) + [2 ** k + d for k in range(16, 31) for d in (-1, 0, 1, 12345)]:
        b = bucket(n)
        assert b >= n and b >= 1 << 16
        if n >= 1 << 16:
            assert b < n * 1.25 + 1
    for n in sorted(range(1 << 16, 1 << 22, 4093)):
        assert bucket(n) >= prev
        prev = bucket(n)
    pol = op._CapacityPolicy()
    assert pol.guess(0, 1_000_000) == bucket(6_000_000)        # first call: 6 instances per Gaussian
    pol.observe(0, 1_300_000)
    first = pol.guess(0, 1_000_000)
    assert first >= 1_300_000 * 1.5
    pol.observe(0, 1_250_000)                                    # smaller view: capacity stays (same allocator block)
    assert pol.guess(0, 1_000_000) == first
    pol.observe(0, 2_450_000)                                    # larger view: grows
    assert pol.guess(0, 1_000_000) >= 2_450_000 * 1.5 > first
    grown = pol.guess(0, 1_000_000)
    pol.observe(0, 100_000)                                      # four-fold smaller: shrinks
    assert pol.guess(0, 1_000_000) < grown


# ---------------------------------------------------------------------------------- round 2 host logic
def test_shard_views_partitions_every_view_exactly_once():
    from g4splat_b200.view_parallel import shard_views
    for n, world in ((50, 4), (64, 8), (7, 3), (3, 8), (0, 2)):
        for strided in (False, True):
            seen = sorted(i for r in range(world) for i in shard_views(n, world, r, strided=strided))
            assert seen == list(range(n)), (n, world, strided)
            sizes = [len(shard_views(n, world, r, strided=strided)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert [len(shard_views(50, 4, r)) for r in range(4)] == [13, 13, 12, 12]        # c3 on 4 GPUs
    assert list(shard_views(16, 8, 3, strided=True)) == [3, 11]


def test_flat_gradient_buffer_layout_pads_blocks_and_follows_rebind():
    """Every block of the view-sharded gradient buffer starts on a 128-byte boundary whatever P is (the kernels' 16-byte
    reductions need aligned rows), p.grad aliases its block, and rebind() lays the buffer out again when P changes."""
    import torch
    from g4splat_b200.view_parallel import ViewShardedGradSync

    def params(P):
        mk = lambda *shape: torch.zeros(*shape, requires_grad=True)
        return {"xyz": mk(P, 3), "f_dc": mk(P, 1, 3), "f_rest": mk(P, 15, 3), "opacity": mk(P, 1), "scaling": mk(P, 2), "rotation": mk(P, 4)}

    p = params(1003)
    sync = ViewShardedGradSync(p, transport="nccl", use_native=False)
    base = sync._store.data_ptr()
    for k, v in p.items():
        assert (sync._views[k].data_ptr() - base) % 128 == 0, k
        assert v.grad.data_ptr() == sync._views[k].data_ptr() and v.grad.shape == v.shape
    assert (sync._accum.data_ptr() - base) % 128 == 0 and (sync.max_radii.data_ptr() - base) % 128 == 0
    assert sync.max_radii.dtype == torch.int32 and sync.max_radii.numel() == 1003
    assert sync.flat.numel() % 4 == 0 and sync.bytes_per_step == sync.flat.numel() * 4 + 1003 * 4
    # statistics in plain torch (the reference arm of the benchmark uses exactly this path)
    g = torch.zeros(1003, 3); g[5] = torch.tensor([3.0, 4.0, 9.0])
    r = torch.zeros(1003, dtype=torch.int32); r[5] = 7; r[6] = 2
    sync.add_view_stats(g, r)
    assert float(sync.xyz_gradient_accum[5]) == 5.0 and float(sync.denom[5]) == 1.0 and float(sync.denom[6]) == 1.0
    assert int(sync.max_radii[5]) == 7 and float(sync.denom.sum()) == 2.0
    q = params(1501)
    sync.rebind(q)
    assert sync.P == 1501 and all(v.grad is not None and v.grad.shape == v.shape for v in q.values())
    assert float(sync._store.abs().sum()) == 0.0
    with __import__("pytest").raises(ValueError):
        ViewShardedGradSync(p, transport="carrier-pigeon", use_native=False)
