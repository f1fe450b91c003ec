"""Golden vectors for the image-space regularisers from the UNMODIFIED reference functions:
normal2curv (matcha/dm_utils/rendering.py:392-406) and compute_depth_order_loss (matcha/dm_regularization/depth.py:142-222),
loaded by file path from /root/reference and run on the CPU with torch autograd.  The depth-order loss draws its pixel
shifts with torch.randint inside the function; the script records that draw (by wrapping torch.randint for the call) so
the tests can hand the same shifts to the code under test.

    python tests/golden/make_golden_regularizers.py      # writes tests/golden/regularizers_*.npz
"""
import importlib.util
import json
import sys
from pathlib import Path
from unittest import mock

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
REF = Path("/root/reference")


def _load(path, name):
    for missing in ("pytorch3d", "pytorch3d.transforms", "pytorch3d.transforms.transform3d"):
        sys.modules.setdefault(missing, mock.MagicMock(name=missing))
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    rendering = _load(REF / "matcha/dm_utils/rendering.py", "_ref_rendering")
    depthreg = _load(REF / "matcha/dm_regularization/depth.py", "_ref_depthreg")
    out = ROOT / "tests" / "golden"

    # ---- normal2curv: all-ones mask (what the trainer passes) and a ragged mask with holes on the border
    for name, (H, W, seed, holes) in {"regularizers_curv_ones": (37, 53, 1, False), "regularizers_curv_mask": (24, 31, 2, True)}.items():
        rng = np.random.default_rng(seed)
        n = rng.normal(size=(3, H, W)).astype(np.float32)
        n /= np.linalg.norm(n, axis=0, keepdims=True)
        n *= rng.random(size=(1, H, W)).astype(np.float32)          # rend_normal is alpha-weighted, not unit length
        mask = np.ones((1, H, W), np.float32)
        if holes:
            mask[0] = (rng.random(size=(H, W)) > 0.25).astype(np.float32)
            mask[0, 0, :5] = 0; mask[0, -1, -3:] = 0; mask[0, 5:9, 0] = 0
        g = rng.normal(size=(1, H, W)).astype(np.float32)
        nt = torch.tensor(n, requires_grad=True)
        curv = rendering.normal2curv(nt, torch.tensor(mask))
        (curv * torch.tensor(g)).sum().backward()
        np.savez_compressed(out / f"{name}.npz", normal=n, mask=mask, g=g, curv=curv.detach().numpy(), dnormal=nt.grad.numpy(),
                            meta=np.array(json.dumps({"reference": "matcha/dm_utils/rendering.py:392-406, CPU fp32", "torch": torch.__version__})))
        print(name, curv.shape, float(curv.sum()))

    # ---- depth-order loss: the trainer's settings (normalised, mean) with and without log space; sum; none
    cases = {"regularizers_order_mean": dict(H=40, W=56, seed=3, normalize_loss=True, log_space=False, reduction="mean", scene_extent=4.7),
             "regularizers_order_log": dict(H=33, W=29, seed=4, normalize_loss=True, log_space=True, reduction="mean", scene_extent=1.0),
             "regularizers_order_raw_sum": dict(H=21, W=34, seed=5, normalize_loss=False, log_space=False, reduction="sum", scene_extent=2.5),
             "regularizers_order_none": dict(H=18, W=25, seed=6, normalize_loss=True, log_space=True, reduction="none", scene_extent=3.0)}
    for name, c in cases.items():
        rng = np.random.default_rng(c["seed"])
        H, W = c["H"], c["W"]
        depth = (2.0 + rng.random(size=(1, H, W)) * 3.0).astype(np.float32)
        prior = (depth * 1.3 + 0.4 * rng.normal(size=(1, H, W))).astype(np.float32)
        prior[0, :2, :3] = depth[0, :2, :3]                        # a few exact ties in the prior
        drawn = []
        real_randint = torch.randint

        def recording_randint(*a, **k):
            t = real_randint(*a, **k)
            drawn.append(t.clone())
            return t

        torch.manual_seed(100 + c["seed"])
        dt = torch.tensor(depth, requires_grad=True)
        with mock.patch.object(torch, "randint", recording_randint):
            loss = depthreg.compute_depth_order_loss(depth=dt, prior_depth=torch.tensor(prior), scene_extent=c["scene_extent"],
                                                     max_pixel_shift_ratio=0.05, normalize_loss=c["normalize_loss"],
                                                     log_space=c["log_space"], log_scale=20., reduction=c["reduction"], debug=False)
        assert len(drawn) == 1
        g = rng.normal(size=tuple(loss.shape)).astype(np.float32) if c["reduction"] == "none" else np.float32(1.7)
        (loss * torch.tensor(g)).sum().backward()
        np.savez_compressed(out / f"{name}.npz", depth=depth, prior=prior, shifts=drawn[0].numpy().astype(np.int64), g=g,
                            loss=loss.detach().numpy(), ddepth=dt.grad.numpy(),
                            meta=np.array(json.dumps({**c, "log_scale": 20.0, "max_pixel_shift_ratio": 0.05,
                                                      "reference": "matcha/dm_regularization/depth.py:142-222, CPU fp32", "torch": torch.__version__})))
        print(name, loss.shape, float(loss.sum()))


if __name__ == "__main__":
    main()
