import sys; sys.path[:0]=['.','oracle','tests']
import numpy as np, common, cryogrid_jl_b200 as cg
for M in (2, 1):
    for lite in (False, True):
        try:
            if lite:
                tile = common.lite_ensemble_tile(ncolumns=M)
            else:
                tile = common.samoylov_freew_tile(ncolumns=M)
            H0 = cg.initialcondition(tile)
            eng = cg.Engine(tile, stepper="lite" if lite else "euler")
            eng.set_state(H0, 0.0)
            eng.advance_to(86400.0)
            print("M", M, "lite", lite, "ok", eng.stats()["column_steps"])
            eng.close()
        except Exception as e:
            print("M", M, "lite", lite, "FAIL", e)
