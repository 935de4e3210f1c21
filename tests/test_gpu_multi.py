"""Multi-GPU tests (need >= 2 GPUs on the box; skipped otherwise): the sharded planner with the
peer-memory reward exchange (include/dial_b200.h: dial_exchange_*) — results equal the NCCL
path bit for bit, every rank holds the same plan, the sharded run matches the unsharded one to
fp32 rounding, and the sharded CUDA-graph control step matches the eager sequence."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_peer_memory_exchange_two_ranks(built):
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "dist_p2p_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=900)
    line = [l for l in out.stdout.splitlines() if l.startswith("P2PCHECK ")]
    assert line, out.stdout[-3000:] + out.stderr[-3000:]
    r = json.loads(line[-1][len("P2PCHECK "):])
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out", "parity"), exist_ok=True)
        json.dump(r, open(os.path.join(ROOT, "gpurun_out", "parity", "p2p_exchange.json"), "w"), indent=1)
    except OSError:
        pass
    assert r["exchange_error"] == 0
    assert r["xch_eq_nccl_rews"] and r["xch_eq_nccl_Y"] and r["xch_vs_nccl_xbar"] < 1e-5
    assert r["ranks_agree_Y"] and r["ranks_agree_xbar"]
    assert r["sharded_vs_single_rews"] < 2e-4 and r["sharded_vs_single_Y"] < 2e-3 and r["sharded_vs_single_xbar"] < 2e-3
    assert r["graph_vs_eager"] < 5e-3 and r["graph_rng_equal"] and r["graph_step"] == 47 + 5
