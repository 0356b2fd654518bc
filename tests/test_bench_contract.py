"""bench.py contract checks that need no GPU: the reference arm (CPU restatement of the path on the host cores)
prints exactly one JSON line with the keys the driver reads, and the product arm refuses to run without a B200."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--config", "tiny_ukbb", "--steps", "1", "--warmup", "1", "--cpu-batch", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hvae_elbo_train_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_the_staged_reference_itself():
    """with baseline/_ref staged (build container / GPU box) the arm times the UNMODIFIED reference (kind = reference)"""
    import pytest
    if not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "src", "vae.py")):
        pytest.skip("reference not staged (run __graft_entry__.build() where /root/reference exists)")
    r = _run(["--impl", "reference", "--config", "morphomnist", "--steps", "1", "--warmup", "1", "--cpu-batch", "4"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(["--steps", "1"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
