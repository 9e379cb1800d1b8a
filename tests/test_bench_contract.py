"""bench.py contract checks that need no GPU: the reference arm (the CPU oracle port timed on the host cores) prints
ONE JSON line with the keys the driver reads, and our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--scale", "0.01", "--ref-batches", "4")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sampled_edges_per_sec_3hop_15_10_5" and d["unit"] == "edges/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"]
    cb, e2e = d["cpu_baseline"], d["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_our_arm_needs_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        return  # on a GPU box the arm runs for real (the driver does that)
    p = _run("--steps", "1", "--warmup", "0", "--scale", "0.01")
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
