"""CPU tests of bench.py's contract: the reference arm's JSON line (keys the driver reads) and the refusal of the B200 arm to
run without a CUDA device (no CPU fallback on the measured path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, env=e,
                          timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run("--impl", "reference", "--workload", "c3", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gll_fwd_bwd_calls_per_sec" and d["unit"] == "calls/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("c3")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and abs(cb["value"] - d["value"]) < 1e-12 and cb["sample"]
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0 and abs(e2e["value"] - d["value"]) < 1e-12
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    """Under torchrun only rank 0 times the CPU arm; the other ranks print nothing and exit 0."""
    p = _run("--impl", "reference", "--workload", "c3", "--steps", "1", "--warmup", "0", "--gpus", "2",
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29571"})
    assert p.returncode == 0, p.stderr[-2000:]
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]


def test_b200_arm_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)
