"""The driver's contract for bench.py, checked on the CPU through the reference arm (`--impl reference` times the
reference's own CPU path -- the unmodified package from baseline/_ref when it is installed, else the oracle port --
on host cores and needs no GPU): ONE JSON line with the agreed keys; the GPU arm refuses to run without a device
instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=170):
    env = dict(os.environ, OPENBLAS_NUM_THREADS="2", OMP_NUM_THREADS="2")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--scale", "0.02"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == "site-node partial updates/sec" and d["unit"] == "updates/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    have_ref = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "phynetpy"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert "cfg5" in d["config"]["workload"]          # the same configuration at every N
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present: the product arm runs")
    p = _run(["--steps", "1", "--warmup", "0", "--scale", "0.01"])
    assert p.returncode != 0
    text = p.stderr + p.stdout
    assert "no CPU fallback" in text or "CUDA" in text or "NVIDIA driver" in text
    assert not [ln for ln in p.stdout.splitlines() if ln.strip().startswith("{") and '"value"' in ln]
