#!/usr/bin/env python
"""Throughput of the five BASELINE.json configurations (scaled so that each finishes in seconds) on one B200.
Prints one JSON line per configuration; the headline config (cfg2) is what bench.py measures in full."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cryogrid_jl_b200 as cg  # noqa: E402
from cryogrid_jl_b200 import presets as common  # noqa: E402  (the synthetic BASELINE problem builders)

DAY = 86400.0


MAX_LAUNCHES = 0


def timed(eng, targets):
    if MAX_LAUNCHES:
        targets = list(targets)[:MAX_LAUNCHES]
    eng.profile_enable(True); eng.profile_read()
    s0 = eng.stats()
    t0 = time.perf_counter()
    for tt in targets:
        eng.advance_to(tt, sync=False)
    eng.synchronize()
    wall = time.perf_counter() - t0
    n, ms = eng.profile_read()
    s1 = eng.stats()
    cs = s1["cell_steps"] - s0["cell_steps"]
    it = s1["lite_iterations"] - s0["lite_iterations"]
    return dict(cell_steps=cs, kernel_s=ms / 1e3, wall_s=wall, cell_steps_per_s=cs / (ms / 1e3), launches=n, lite_iterations=it,
                steps_per_column=(s1["column_steps"] - s0["column_steps"]) / s1["n_columns"], status_or=s1["status_or"])


def tile_repeat(tile_small, H0_small, M):
    """Repeat a small ensemble to M columns (parameters and initial state), for throughput measurements."""
    reps = -(-M // tile_small.M)
    idx = np.tile(np.arange(tile_small.M), reps)[:M]
    return idx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="")
    ap.add_argument("--launches", type=int, default=0, help="time only the first N launches of every configuration (ncu captures)")
    args = ap.parse_args()
    global MAX_LAUNCHES
    MAX_LAUNCHES = args.launches
    out = []
    only = set(args.only.split(",")) if args.only else None

    def want(name):
        return only is None or name in only or name[:4] in only      # cfg4 selects cfg4a, cfg4b, cfg4c

    if want("cfg1"):
        # cfg1a: examples/heat_sfcc_constantbc.jl AS WRITTEN -- the script defines an SFCC but SimpleSoil() keeps its default
        # FreeWater (oddity 1): one FreeWater column, constant 10 W/m² in/out, dtmax = 900 s, 30 days
        tile = cg.SoilHeatTile("H", cg.ConstantBC(cg.Neumann, 10.0), cg.ConstantBC(cg.Neumann, 10.0), cg.SoilProfile((0.0, cg.SimpleSoil())), None,
                               cg.initializer("T", cg.TemperatureProfile((0.0, -20.0), (1000.0, 10.0))), grid=cg.DefaultGrid_10cm, ncolumns=1)
        H0 = cg.initialcondition(tile)
        eng = cg.Engine(tile, dtmax=900.0); eng.set_state(H0, 0.0); eng.advance_to(DAY)
        r = timed(eng, [DAY * d for d in range(2, 32)]); r.update(config="cfg1a single FreeWater column (example as written), N=178, 30 days", columns=1)
        out.append(r); eng.close()
        # cfg1b: single SFCC column (PainterKarra α=1, n=2), constant 10 W/m² in/out, dtmax=900 s, 30 days
        tile = common.sfcc_tile(ncolumns=1, kind="painterkarra", solver="lut")
        H0 = cg.initialcondition(tile)
        eng = cg.Engine(tile, dtmax=900.0); eng.set_state(H0, 0.0); eng.advance_to(DAY)
        r = timed(eng, [DAY * d for d in range(2, 32)]); r.update(config="cfg1b single SFCC-LUT column, N=178, 30 days", columns=1)
        out.append(r); eng.close()
    if want("cfg2"):
        M = int(10000 * args.scale)
        tile = common.samoylov_freew_tile(ncolumns=M, seed=1, forcing_days=70)
        H0 = cg.initialcondition(tile)
        eng = cg.Engine(tile); eng.set_state(H0, 0.0); eng.advance_to(10 * DAY)
        r = timed(eng, [DAY * d for d in range(11, 61)]); r.update(config="cfg2 FreeWater CGEuler, N=218, 50 days", columns=M)
        out.append(r); eng.close()
    if want("cfg3"):
        M = int(100000 * args.scale)
        small = common.lite_ensemble_tile(ncolumns=2048, seed=1234, forcing_days=800)
        tile = common.lite_ensemble_tile(ncolumns=M, seed=1234, forcing_days=800)
        # initial state from 2048 distinct members repeated (host steady-state init of 1e5 members is a "next" row)
        H0s = cg.initialcondition(small)
        idx = tile_repeat(small, H0s, M)
        for name in ("por", "sat", "org"):
            setattr(tile, name, getattr(small, name)[:, idx])
        tile.nf, tile.nt = small.nf[idx], small.nt[idx]
        H0 = np.ascontiguousarray(H0s[:, idx])
        eng = cg.Engine(tile, stepper="lite"); eng.set_state(H0, 0.0); eng.advance_to(5 * DAY)
        r = timed(eng, [DAY * d for d in range(6, 371)]); r.update(config="cfg3 LiteImplicitEuler ensemble, N=278, 365 daily steps", columns=M)
        out.append(r)
        # the same steps with one launch per 5 days (a column stays in registers for 5 implicit steps)
        r = timed(eng, [DAY * d for d in range(375, 741, 5)]); r.update(config="cfg3 LiteImplicitEuler ensemble, N=278, 365 daily steps, 5 per launch", columns=M)
        out.append(r); eng.close()
    for name, variant, label in (("cfg4a", "newton", "cfg4 SFCC DallAmico per-column parameters (Newton), N=218, 4 days"),
                                 ("cfg4b", "lut", "cfg4' SFCC DallAmico presolver LUT (one soil class: SFCC fast path), N=218, 4 days"),
                                 ("cfg4c", "lut_n2", "cfg4'' SFCC PainterKarra n=2 presolver LUT (one soil class: SFCC fast path), N=218, 4 days"),
                                 ("cfg4d", "lut", "cfg4' through the GENERAL LUT kernel (CGB_SFCC_FAST=0), N=218, 4 days")):
        if not want(name):
            continue
        M = int(100000 * args.scale)
        tile = common.cfg4_tile(M, variant)
        H0 = cg.initialcondition(tile)
        if name == "cfg4d":
            os.environ["CGB_SFCC_FAST"] = "0"
        eng = cg.Engine(tile); eng.set_state(H0, 0.0); eng.advance_to(1 * DAY)
        os.environ.pop("CGB_SFCC_FAST", None)
        r = timed(eng, [DAY * d for d in range(2, 6)]); r.update(config=label, columns=M)
        out.append(r); eng.close()
    if want("cfg5"):
        M = int(100000 * args.scale)
        tile = common.cfg5_tile(M)
        eng = cg.SaltEngine(tile); eng.set_state(tile.initialcondition(), 0.0); eng.advance_to(1 * DAY)
        r = timed(eng, [DAY * d for d in range(2, 6)]); r.update(config="cfg5 heat+salt, N=178, 4 days", columns=M)
        out.append(r); eng.close()
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
