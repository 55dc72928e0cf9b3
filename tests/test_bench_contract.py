"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the keys the driver reads, from the
oracle's timing copy on a bounded sample; the b200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-columns", "32", "--days-per-step", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "column-cell-steps/sec" and d["unit"] == "cell-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_timing_copy_is_reproducible_and_equals_the_parity_oracle(orc):
    # The timing copy of the oracle (oracle/libcg_oracle_fast.so: -O3, hoisted invariants) caches the grid arrays of a work block by
    # address.  A second batch call used to run on a zeroed grid when its fresh work block landed on the address just freed (garbage
    # trajectories, a reference arm 4-5x too slow).  Gate: a repeated call reproduces the first bit for bit, and both equal the parity
    # oracle's trajectory (same arithmetic, different optimisation flags: steps equal, H within 1e-9 relative).
    import numpy as np
    import common
    import cryogrid_jl_b200 as cg
    tile = common.samoylov_freew_tile(ncolumns=16, seed=1, forcing_days=12)
    H0 = cg.initialcondition(tile)
    cols = np.arange(16)

    def run(prepared):
        H = np.ascontiguousarray(H0.T); t = np.zeros(16); dt = np.full(16, 60.0)
        st, ns = orc.cgeuler_advance_batch(tile, cols, H, t, dt, 5 * 86400.0, nthreads=4, dtmax=600.0, prepared=prepared)
        assert not np.any(st)
        return H, ns.copy()
    try:
        orc.use_timing_copy(True)
        prep = orc.batch_columns(tile, cols)
        Ha, na = run(prep)
        Hb, nb = run(prep)
        Hc, nc = run(orc.batch_columns(tile, cols))
    finally:
        orc.use_timing_copy(False)
    Hp, np_ = run(orc.batch_columns(tile, cols))
    assert np.array_equal(Ha, Hb) and np.array_equal(Ha, Hc) and np.array_equal(na, nb) and np.array_equal(na, nc)
    assert np.array_equal(na, np_)
    assert np.max(np.abs(Ha - Hp) / (np.abs(Hp) + 1e3)) < 1e-9
