"""bench.py contract on CPU: the reference arm (`--impl reference`, the oracle port on the host cores) prints ONE JSON
line with the keys the driver reads; the product arm must refuse to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.slow
def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"], cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "images/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("full-batch step images/s") and line["value"] > 0
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == pytest.approx(line["value"])
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0 and e2e["value"] == pytest.approx(line["value"])


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_product_arm_fails_loudly_without_a_gpu():
    out = subprocess.run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], cwd=ROOT, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode != 0
    assert "no CPU path" in out.stderr
