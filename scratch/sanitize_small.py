"""Tiny run of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import sys; sys.path[:0]=['.','oracle','tests']
import numpy as np, common, cryogrid_jl_b200 as cg
DAY=86400.0
# explicit fast + diag (ragged M, N=218)
t=common.samoylov_freew_tile(ncolumns=13, seed=1, forcing_days=5); H0=cg.initialcondition(t)
e=cg.Engine(t); e.set_state(H0,0.0); e.advance_to(0.25*DAY); e.diag("T",0.25*DAY); e.rhs(0.25*DAY); e.ensemble_partial("T",0.25*DAY)
e.solve_saveat(np.array([0.3*DAY,0.35*DAY]),("T","H","θw")); e.close()
# general explicit: SFCC lut + newton, small grid (N=60 -> CPL=2)
t=common.sfcc_tile(ncolumns=5, kind="painterkarra", solver="lut", grid=cg.DefaultGrid_1m); H0=cg.initialcondition(t)
e=cg.Engine(t); e.set_state(H0,0.0); e.advance_to(0.5*DAY); e.diag("dHdT",0.5*DAY); e.close()
t=common.sfcc_tile(ncolumns=5, kind="dallamico", solver="newton", per_column=True, grid=cg.DefaultGrid_20cm); H0=cg.initialcondition(t)
e=cg.Engine(t); e.set_state(H0,0.0); e.advance_to(0.2*DAY); e.close()
# lite warp + thread
import os
for v in ("warp","thread"):
    os.environ["CGB_LITE_VARIANT"]=v
    t=common.lite_ensemble_tile(ncolumns=7, seed=3, forcing_days=10); H0=cg.initialcondition(t)
    e=cg.Engine(t, stepper="lite"); e.set_state(H0,0.0); e.advance_to(3*DAY); e.diag("DT_ap",3*DAY); e.close()
# salt
t=cg.SalineTile(ncolumns=6, salt_bc=cg.SaltGradient(np.linspace(600,1000,6),0)); e=cg.SaltEngine(t); e.set_state(t.initialcondition(),0.0)
e.advance_to(0.1*DAY); e.diag("dc",0.1*DAY); e.close()
print("sanitize run done")
