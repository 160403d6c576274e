"""bench.py's reference arm (the CPU leg the driver launches as `bench.py --impl reference`) runs without a GPU and prints ONE JSON
line that carries the contract's keys; the product arm must not fall back to the CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "slices/s" and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["data"] == "synthetic" and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("C2") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    head = "CT slices/sec ViT dense-descriptor extraction"          # BASELINE.json's metric, up to its "at 1/2/4/8 B200; % peak" tail
    assert base["metric"].startswith(head) and d["metric"].startswith(head)


def test_product_arm_has_no_cpu_fallback():
    """Without a GPU the product arm must fail loudly, never print a benchmark line from a CPU path."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0
    assert not any(ln.startswith("{") for ln in r.stdout.splitlines())
