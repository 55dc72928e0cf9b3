"""CryoGridProblem / solve / CGEuler / LiteImplicitEuler / CryoGridOutput -- the reference's solver front-end
(problem.jl:47-103, Solvers/basic_solvers.jl:9-79, Solvers/integrator.jl:84-132,
Solvers/LiteImplicit/cglite_types.jl:1-68, IO/output.jl:59-130) for a batched Tile on one GPU."""
from dataclasses import dataclass

import numpy as np

from .engine import Engine


@dataclass
class CGEuler:
    pass


@dataclass
class LiteImplicitEuler:
    miniters: int = 2
    maxiters: int = 1000
    tolerance: float = 1e-3
    verbose: bool = True


class CryoGridProblem:
    def __init__(self, tile, u0, tspan, saveat=3600.0, savevars=()):
        self.tile, self.u0, self.tspan = tile, np.asarray(u0, dtype=np.float64), (float(tspan[0]), float(tspan[1]))
        if np.ndim(saveat) == 0:
            n = int(np.floor((self.tspan[1] - self.tspan[0]) / float(saveat)))
            saveat = self.tspan[0] + float(saveat) * np.arange(n + 1)   # expandtstep: tspan[1]:saveat:tspan[end]
        self.saveat = np.asarray(saveat, dtype=np.float64)
        self.savevars = tuple(savevars)


class CryoGridSolution:
    def __init__(self, prob, t, u, saved, stats, status, steps, retcode):
        self.prob, self.t, self.u, self.saved, self.stats, self.status, self.steps, self.retcode = \
            prob, t, u, saved, stats, status, steps, retcode


def solve(prob, alg=None, dt=None, dtmax=None, dtmin=1.0, device=0, engine=None, out=None, return_state=True):
    """solve(prob, alg; dt, dtmax, dtmin).  Saved variables come back as [nsave, nvars, N, M]."""
    alg = alg or CGEuler()
    lite = isinstance(alg, LiteImplicitEuler)
    eng = engine or Engine(prob.tile, stepper="lite" if lite else "euler", dt0=dt, dtmin=dtmin, dtmax=dtmax, device=device,
                           **(dict(miniters=alg.miniters, maxiters=alg.maxiters, tolerance=alg.tolerance) if lite else {}))
    t0 = prob.tspan[0]
    eng.set_state(prob.u0, t0)
    saves = prob.saveat[prob.saveat >= t0]
    savevars = prob.savevars or ("H",)
    shape = (saves.size, len(savevars), eng.N, eng.M)
    if out is None:
        out = np.empty(shape)          # pass a pinned buffer for full-speed D2H
    assert out.shape == shape
    first = 0
    if saves.size and saves[0] == t0:           # initial save: basic_solvers.jl:33-37
        for v, name in enumerate(savevars):
            if name == "H":
                out[0, v] = prob.u0
            else:   # written straight into the (possibly pinned) output buffer
                eng.diag({"θw": "thw", "T_now": "T", "θw_now": "thw", "thw_now": "thw"}.get(name, name), t0, out=out[0, v])
        first = 1
    if saves.size > first:
        eng.solve_saveat(saves[first:], savevars, out=out[first:])
    if saves.size == 0 or saves[-1] < prob.tspan[1]:
        eng.advance_to(prob.tspan[1])
    u = eng.get_state() if return_state else None      # sol.u[end]; skip when only the saved outputs are consumed
    status, steps = eng.status()
    retcode = "Success" if not status.any() else ("MaxIters" if (status & 1).any() and not (status & ~1).any() else "Failure")
    sol = CryoGridSolution(prob, saves, u, out, eng.stats(), status, steps, retcode)
    if engine is None:
        eng.close()
    return sol


class CryoGridOutput:
    """Saved variables by name: out.T is [nsave, N, M] (time, depth, column)."""

    def __init__(self, sol):
        self.ts = sol.t
        self.z = sol.prob.tile.grid.cells
        self.vars = {}
        for v, name in enumerate(sol.prob.savevars or ("H",)):
            self.vars[name] = sol.saved[:, v]

    def __getattr__(self, name):
        try:
            return self.__dict__["vars"][name]
        except KeyError:
            raise AttributeError(name)
