import sys, time; sys.path[:0]=['.','oracle','tests']
import numpy as np, torch, common, cryogrid_jl_b200 as cg
DAY=86400.0
M=10000; D=73; W=219
tile = common.samoylov_freew_tile(ncolumns=M, seed=1, forcing_days=W+D+3)
H0 = cg.initialcondition(tile)
N=tile.grid.ncells
eng = cg.Engine(tile)
eng.set_state(H0,0.0); eng.advance_to(W*DAY)
u0 = torch.empty((N,M), dtype=torch.float64, pin_memory=True).numpy(); u0[:]=eng.get_state()
out = torch.empty((D+1,1,N,M), dtype=torch.float64, pin_memory=True).numpy()
t0w=W*DAY
prob = cg.CryoGridProblem(tile, u0, (t0w, t0w+D*DAY), saveat=DAY, savevars=("T",))
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0=T(); sol = cg.solve(prob, cg.CGEuler(), engine=eng, out=out, return_state=False); t1=T()
    print("solve total %.1f ms  cell_steps %.3e -> %.3e /s"%((t1-t0)*1e3, sol.stats["cell_steps"], sol.stats["cell_steps"]/(t1-t0)))
saves = t0w + DAY*np.arange(1,D+1)
t0=T(); eng.set_state(u0, t0w); t1=T(); print("set_state %.2f ms"%((t1-t0)*1e3))
t0=T(); x=eng.diag("T",t0w); t1=T(); print("diag T (pageable out) %.2f ms"%((t1-t0)*1e3))
t0=T(); eng.solve_saveat(saves, ("T",), out=out[1:]); t1=T(); print("solve_saveat T %.2f ms"%((t1-t0)*1e3))
eng.set_state(u0, t0w)
t0=T(); eng.solve_saveat(saves, ("H",), out=out[1:]); t1=T(); print("solve_saveat H %.2f ms"%((t1-t0)*1e3))
eng.set_state(u0, t0w)
t0=T()
for s in saves: eng.advance_to(s, sync=False)
eng.synchronize(); t1=T(); print("advance only %d launches %.2f ms"%(len(saves),(t1-t0)*1e3))
eng.set_state(u0, t0w)
t0=T(); eng.advance_to(saves[-1]); t1=T(); print("advance one launch %.2f ms"%((t1-t0)*1e3))
t0=T(); st=eng.status(); s=eng.stats(); t1=T(); print("status+stats %.2f ms"%((t1-t0)*1e3))
