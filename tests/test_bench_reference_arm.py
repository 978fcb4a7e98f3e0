"""bench.py --impl reference: the reference's own CPU code (oracle/_ref) timed on a bounded
sample, printed as the contract's JSON line. Runs on the CPU; skipped where _ref was not built."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

import oracle

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built")
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--genomes", "6", "--length", "60000",
           "--steps", "1", "--warmup", "0", "--cpu-queries", "4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    """Under torchrun only rank 0 runs the arm; the other ranks exit 0 without work."""
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
