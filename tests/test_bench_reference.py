"""CPU: `bench.py --impl reference` (the unmodified reference over the product's chunks on the host cores) prints the
contract's JSON line and respects its wall-clock budget - round 1's arm overran the driver's limit at every N."""
import json
import os
import subprocess
import sys
import time

import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/slimfastq is not built")
def test_reference_arm_is_bounded_and_well_formed():
    t0 = time.time()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-budget-s", "20"],
                       capture_output=True, text=True, timeout=300)
    wall = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1
    assert line["config"]["workload"].startswith("illumina_2x150_phred40_10GB") and line["config"]["chunk_bytes"] == 1 << 20
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == (os.cpu_count() or 1) and cb["spread"]["passes"] == 2
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # calibration + 3 passes inside 20 s, plus the best-case leg (one pass over `cores` whole files): generous bound
    assert wall < 150, wall
