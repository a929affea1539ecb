"""Sequence-sharded encoder in host C (vb_dist.c: NCCL K/V halo + adapter all-gather) == single-GPU pipeline.  Needs >= 2
GPUs; on a 1-GPU box the unsharded run of the same call (own-rows conv stem, split layers, vox_cuda_decode_adapter) is still
checked against the streaming API."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _run(nproc, seconds):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tools", "sharded_run.py"), str(seconds)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def test_sharded_call_unsharded_matches_stream_api_single_gpu():
    out = _run(1, 8)
    assert out["tokens_equal_stream_api"] and out["adapter_max_abs_diff_vs_unsharded"] == 0.0


def test_two_rank_halo_exchange_is_exact():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, 60)
    assert out["world"] == 2 and out["tokens_equal_stream_api"]
    assert out["adapter_max_abs_diff_vs_unsharded"] <= 1e-4 * max(out["adapter_scale"], 1.0)
