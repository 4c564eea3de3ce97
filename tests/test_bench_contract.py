"""bench.py's contract on a machine without a GPU: the reference arm runs the compiled reference on the host cores and
prints one JSON line with the keys the driver reads; the CUDA arm refuses to run (there is no CPU fallback)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import oracle_py as orc

ROOT = Path(__file__).resolve().parent.parent


def run_bench(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.fspath(ROOT / "bench.py"), *args], capture_output=True, text=True, env=env, timeout=600)


@pytest.mark.skipif(not orc.have_reference(), reason="oracle/_ref/psi_ref_driver not built (needs /root/reference)")
def test_reference_arm_prints_the_contract_line():
    p = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" not in line
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config"):
        assert key in line, key
    assert line["unit"] == "reads/s" and line["value"] > 0 and line["higher_is_better"] is True
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_does_no_work_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.fspath(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_cuda_arm_fails_loudly_without_a_gpu():
    p = run_bench("--steps", "1")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout) and "no CPU fallback" in (p.stderr + p.stdout)
