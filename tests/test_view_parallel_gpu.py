"""Multi-GPU correctness on hardware (skipped below two visible GPUs): NCCL + the CUDA operator + the
kernel-side gradient sink together, for both transports of g4splat_b200.view_parallel -- N-rank sums ==
a one-rank loop over the same views (tests/tools/vp_check.py does the work under torchrun)."""
import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_gradient_sum_matches_single_rank():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "tools" / "vp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-4000:]
    rep = json.loads(lines[-1])
    assert rep["ok"], rep
    assert rep["nccl"]["max_rel_err"] <= 1e-4
    if rep["multimem_available"]:
        assert rep["multimem"]["max_rel_err"] <= 1e-4
        assert rep["multimem_red"]["max_rel_err"] <= 1e-4
