"""bench.py contract on the CPU: the reference arm (`--impl reference`: the reference's own CPU code on the host cores) prints ONE JSON
line with the keys the driver reads, for every workload switch; the GPU arm cannot run here (no device) and must fail loudly rather
than fall back."""
import json
import os
import subprocess
import sys

import pytest

import bindings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*extra):
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--frames", "1", "--ref-frames", "1", "--scale", "0.1",
           "--iterations", "2"] + list(extra)
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


@pytest.mark.skipif(not bindings.Reference.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("extra", [(), ("--condition", "ra"), ("--bits", "11", "--scale", "0.08")])
def test_reference_arm_prints_the_contract_line(extra):
    res = run_bench("--impl", "reference", *extra)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpoints/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["metric"].startswith("Mpoints/s patch-gen+image-formation")


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    res = run_bench()
    assert res.returncode != 0 and not any(l.startswith("{") for l in res.stdout.splitlines())
