"""Multi-GPU product-path parity: needs >= 2 CUDA devices, skipped on a 1-GPU box.  The same check
runs inside every N > 1 `bench.py` invocation and lands in its JSON line as `multi_rank_parity`."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_update_equals_single_rank(built_lib):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--parity-only"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)["multi_rank_parity"]
    print(res)
    assert res["params_bit_identical_across_ranks"]
    assert res["grad_rel_l2_pass1_vs_1rank"] <= 1e-4
    assert res["param_change_rel_l2_4pass_vs_1rank"] <= 5e-2
