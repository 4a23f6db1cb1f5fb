"""CPU tier: the parts of bench.py's contract that do not need a GPU -- the reference arm (the CPU oracle prover on the
16-byte circuit) prints exactly ONE JSON line on stdout with the keys the driver reads, whatever native code writes to fd 1;
the device arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    # quick mode (a 2^14-constraint prefix of the circuit): the arm's default, the full 16-byte proof, takes ~2 minutes on 8 cores
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-constraints", "16384"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "encrypt_prove_constraints_per_s" and d["unit"] == "constraints/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm names the configuration it really ran: the 16-byte circuit (BASELINE configs[0]), here its prefix sample -- never "4096-byte"
    assert "16-byte message" in d["config"]["workload"] and "PREFIX SAMPLE" in d["config"]["workload"] and "4096" not in d["config"]["workload"]
    assert d["config"]["constraints"] == 16384 and len(d["config"]["proof_sha256"]) == 64


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_device_arm_has_no_cpu_fallback():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert p.returncode != 0 and p.stdout.strip() == ""
    assert "no CUDA device" in p.stderr
