#!/usr/bin/env python
"""BASELINE.json north_star target: ">= 10^5 one-year SFCC columns on DefaultGrid_5cm integrated in under one minute on
8 x B200".  Integrates `--columns` SFCC columns (DallAmico, presolver LUT or on-device Newton with per-column parameters)
for one simulated year on ONE GPU and prints the wall time; 10^5 columns on one GPU is 8x the per-GPU share of the target."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import common  # noqa: E402
import cryogrid_jl_b200 as cg  # noqa: E402

DAY = 86400.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, default=100000)
    ap.add_argument("--solver", default="lut", choices=["lut", "newton"])
    ap.add_argument("--days", type=int, default=365)
    args = ap.parse_args()
    M = args.columns
    rng = np.random.default_rng(4)
    kw = dict(grid=cg.DefaultGrid_5cm, top=cg.PeriodicBC(cg.Dirichlet, common.YEAR, 18.0, 0.0, -10.0),
              bottom=cg.GeothermalHeatFlux(0.053), temp=cg.SamoylovDefault_tempprofile)
    if args.solver == "lut":
        tile = common.sfcc_tile(ncolumns=M, kind="dallamico", solver="lut", alpha=2.0, n=1.6, **kw)
    else:
        tile = common.sfcc_tile(ncolumns=M, kind="dallamico", solver="newton", per_column=True, seed=4, **kw)
    tile.top_offset = rng.normal(0.0, 2.0, M)
    eng = cg.Engine(tile)
    t0 = time.perf_counter()
    eng.initialcondition(0.0)                       # on-device initial condition
    eng.synchronize()
    t_init = time.perf_counter() - t0
    eng.profile_enable(True); eng.profile_read()
    t0 = time.perf_counter()
    hops = list(range(73, args.days + 1, 73))
    if not hops or hops[-1] != args.days:
        hops.append(args.days)
    for d in hops:
        eng.advance_to(d * DAY, sync=False)
    eng.synchronize()
    wall = time.perf_counter() - t0
    n, ms = eng.profile_read()
    s = eng.stats()
    print(json.dumps(dict(config=f"{M} SFCC DallAmico columns ({args.solver}), DefaultGrid_5cm (N={tile.grid.ncells}), {args.days} days, CGEuler",
                          columns=M, wall_s=wall, kernel_s=ms / 1e3, init_s=t_init, cell_steps=s["cell_steps"],
                          cell_steps_per_s=s["cell_steps"] / (ms / 1e3), steps_per_column=s["column_steps"] / M,
                          status_or=s["status_or"], launches=n)))
    eng.close()


if __name__ == "__main__":
    main()
