"""Generate the golden vectors: outputs of the UNMODIFIED reference extension (oracle/_ref, built
by oracle/build_ref.py from /root/reference) on seeded synthetic cases, on a B200.

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/

Each file holds the reference's (color, allmap, radii) and the eight gradients for the upstream
gradients of synthetic.make_upstream_grads, plus the JSON `meta` that rebuilds the inputs
(tests/helpers.py: case_from_meta).  Inputs are regenerated from the seed, never stored.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import helpers as Hh  # noqa: E402

CASES = ["tiny", "scalemod", "c0_deg1", "ragged", "c0_bg", "c0_precomp"]


def main():
    from oracle import build_ref
    import torch
    from oracle.oracle import Oracle
    ref = build_ref.import_reference()
    o32 = Oracle("f32")
    out_dir = ROOT / "gpurun_out" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    for name in CASES:
        meta = dict(Hh.NAMED_CASES[name], name=name)
        case = Hh.case_from_meta(meta, o32)
        r = Hh.run_operator(ref, case)
        arrays = {k: r[k] for k in Hh.FWD_KEYS + Hh.GRAD_KEYS}
        arrays["meta"] = np.array(json.dumps(meta))
        arrays["provenance"] = np.array(json.dumps({
            "device": torch.cuda.get_device_name(0), "torch": torch.__version__,
            "reference": "diff-surfel-rasterization @ G4Splat ec07361, sm_100a build (oracle/build_ref.py)"}))
        np.savez_compressed(out_dir / f"{name}.npz", **arrays)
        print(name, {k: v.shape for k, v in arrays.items() if hasattr(v, "shape") and v.ndim}, flush=True)


if __name__ == "__main__":
    main()
