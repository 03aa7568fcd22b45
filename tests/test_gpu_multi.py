"""GPU, >= 2 devices (skipped on a single-GPU box): the product-side multi-GPU path.  SeqLikelihoodEngine(shard=...)
under torch.distributed.run, one rank per GPU: each rank's kernel writes its loci's values into its slice of a device
vector, one in-place ncclAllGather on the same stream, one pinned D2H copy; chains with single-locus and all-loci
proposals, accepts and rejects, must reproduce the single-GPU engine bit for bit on every rank
(reference loop: src/_mcmc_seq.py:746-761 is a plain sum over loci)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_sharded_engine_on_devices():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-3000:])
    assert "SHARDED_ENGINE_OK ranks=2" in p.stdout, p.stdout[-2000:]
