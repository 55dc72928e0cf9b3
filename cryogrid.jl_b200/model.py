"""Host-side mirror of the reference's model-description API for the heat-column hot path.

Names follow CryoGrid.jl: SimpleSoil (Soils/para/simple.jl:8-15), FreeWater / PainterKarra / DallAmico /
VanGenuchten (FreezeCurves.jl), SFCCPreSolver (Soils/ground.jl:27), SoilProfile / TemperatureProfile
(heat_methods.jl:113-114), TemperatureBC + NFactor (heat_bc.jl:42-86), ConstantBC / PeriodicBC
(simple_bc.jl:14-38), GeothermalHeatFlux (heat_bc.jl:108-115), HeatBalance (heat_types.jl:28-45),
Ground (Soils/ground.jl:15-21), Stratigraphy / Tile / SoilHeatTile (Tiles/tile.jl:66-98, models.jl:6-24),
Interpolated1D / Input (IO/forcings/providers.jl:9-32, IO/input.jl:9-22).

The batch dimension: every numeric parameter may be a scalar or an array of length `ncolumns`; a Tile
therefore describes M independent columns sharing one grid and one stratigraphy -- the trajectory index of
the reference's EnsembleProblem (Solvers/DiffEq/ensemble.jl:33-45) becomes the column index.
"""
from dataclasses import dataclass, field
from typing import Any, Optional, Sequence

import numpy as np

from . import _lib as L
from . import hostmath as hm
from .grids import Grid, DefaultGrid_5cm

Dirichlet = "Dirichlet"
Neumann = "Neumann"


# ---- freeze curves ---------------------------------------------------------------------------------
@dataclass
class FreeWater:
    pass


@dataclass
class VanGenuchten:
    alpha: Any = 1.0
    n: Any = 2.0
    theta_res: Any = 0.0


@dataclass
class DallAmico:
    swrc: VanGenuchten = field(default_factory=VanGenuchten)
    Tm: Any = 0.0
    kind = L.FC_DALLAMICO


@dataclass
class PainterKarra:
    swrc: VanGenuchten = field(default_factory=VanGenuchten)
    beta: Any = 1.0
    omega: Any = None   # default 1/β
    Tm: Any = 0.0
    kind = L.FC_PAINTERKARRA


@dataclass
class DallAmicoSalt:
    swrc: VanGenuchten = field(default_factory=VanGenuchten)
    Tm: Any = 0.0
    kind = L.FC_DALLAMICOSALT


@dataclass
class SFCCPreSolver:
    """Presolver LUT.  `table` may carry caller-built knots (Hk, Tk) -- e.g. sampled from the Julia
    interpolant -- which makes LUT parity independent of our table builder."""
    Tmin: float = -60.0
    errtol: float = 1e-4
    Tmax: float = 0.0
    extrap: str = "line"
    table: Optional[tuple] = None


@dataclass
class SFCCNewtonSolver:
    """On-device Newton to ~1e-13 K (accuracy mode; needed for per-column SFCC parameters)."""
    pass


# ---- soil ------------------------------------------------------------------------------------------
@dataclass
class SimpleSoil:
    por: Any = 0.5
    sat: Any = 1.0
    org: Any = 0.0
    freezecurve: Any = field(default_factory=FreeWater)


@dataclass
class Heterogeneous:
    """Heterogeneous{SimpleSoil} (Soils/para/simple.jl:92-182): soil parameters vary from cell to cell; the reference
    initialises the per-cell fields por/sat/org by interpolating the profile linearly in depth (InterpInitializer, flat
    outside).  Use it as the `para` of a Ground layer; the Tile expands it to one parameter set per cell."""
    profile: Any                  # SoilProfile((z, SimpleSoil(...)), ...)


class Profile(list):
    """Ordered (depth, value) pairs."""

    def __init__(self, *pairs):
        super().__init__(sorted(pairs, key=lambda p: p[0]))


SoilProfile = Profile
TemperatureProfile = Profile


# ---- boundary conditions / inputs ------------------------------------------------------------------
@dataclass
class Input:
    name: str


class Interpolated1D:
    """Forcing provider: gridded-linear interpolation in time, error outside the knots."""

    def __init__(self, ts, **series):
        self.ts = np.asarray(ts, dtype=np.float64)
        self.data = {k: np.asarray(v, dtype=np.float64) for k, v in series.items()}
        for k, v in self.data.items():
            if v.shape != self.ts.shape:
                raise ValueError(f"forcing '{k}' does not match the time axis")

    def __call__(self, inp, t):
        if isinstance(inp, Input):
            inp = inp.name
        t = np.asarray(t, dtype=np.float64)
        if np.any(t < self.ts[0]) or np.any(t > self.ts[-1]):
            raise IndexError("BoundsError: forcing queried outside its time range")
        y = self.data[inp]
        i = np.clip(np.searchsorted(self.ts, t, side="right") - 1, 0, self.ts.size - 2)
        w = (t - self.ts[i]) / (self.ts[i + 1] - self.ts[i])
        return (1.0 - w) * y[i] + w * y[i + 1]


@dataclass
class NFactor:
    nf: Any = 1.0
    nt: Any = 1.0


@dataclass
class TemperatureBC:
    T: Any                     # float | Input | PeriodicBC-like
    effect: Optional[NFactor] = None
    offset: Any = 0.0          # per-column additive shift of the forcing (ensemble perturbation)
    kind = Dirichlet


@dataclass
class ConstantBC:
    kind: str
    value: Any


@dataclass
class PeriodicBC:
    kind: str
    period: float = 1.0
    amplitude: float = 1.0
    phaseshift: float = 0.0
    level: float = 0.0


def ConstantTemperature(value):
    return ConstantBC(Dirichlet, value)


def ConstantFlux(value):
    return ConstantBC(Neumann, value)


@dataclass
class GeothermalHeatFlux:
    Qgeo: Any = 0.053
    kind = Neumann


# ---- processes / layers ----------------------------------------------------------------------------
@dataclass
class MaxDelta:
    dmax: float = 50e3      # 50 kJ (heat_types.jl:18)


@dataclass
class CFL:
    courant_number: float = 0.5
    maxdelta: MaxDelta = field(default_factory=lambda: MaxDelta(np.inf))


@dataclass
class HeatBalance:
    op: str = "H"            # "H" = Diffusion1D(:H); "T" = Diffusion1D(:T); "implicit" = EnthalpyImplicit
    dtlim: Any = None        # default_dtlim (heat_types.jl:17-18): MaxDelta(50 kJ) for :H, CFL(0.5, MaxDelta(Inf)) for :T
    prop: hm.ThermalProperties = field(default_factory=hm.ThermalProperties)

    def __post_init__(self):
        if self.dtlim is None:
            self.dtlim = CFL(0.5, MaxDelta(float("inf"))) if self.op == "T" else MaxDelta()


def EnthalpyImplicit():
    return "implicit"


@dataclass
class Ground:
    para: SimpleSoil = field(default_factory=SimpleSoil)
    heat: HeatBalance = field(default_factory=HeatBalance)
    fcsolver: Any = field(default_factory=SFCCPreSolver)


@dataclass
class Top:
    bc: Any


@dataclass
class Bottom:
    bc: Any


class Stratigraphy:
    def __init__(self, top, layers, bottom):
        self.z_top, self.top = top
        self.layers = sorted(list(layers), key=lambda p: p[0])
        self.z_bot, self.bottom = bottom


@dataclass
class ThermalSteadyStateInit:
    T0: float = 0.0
    Qgeo: float = 0.053
    maxiters: int = 100
    thresh: float = 1e-6


def initializer(varname, value):
    if varname != "T":
        raise ValueError("only :T initializers are part of the heat hot path")
    if isinstance(value, Profile):
        return ("T", value)
    return ("T", Profile((0.0, float(value))))


def _col(v, M):
    a = np.asarray(v, dtype=np.float64)
    if a.ndim == 0:
        return np.full(M, float(a))
    if a.shape != (M,):
        raise ValueError(f"per-column parameter must have length {M}, got {a.shape}")
    return a.copy()


class Tile:
    """Batched tile: M columns, one grid, one stratigraphy (Tiles/tile.jl:6-98)."""

    def __init__(self, strat: Stratigraphy, grid: Grid, inputs: Optional[Interpolated1D] = None, *inits, ncolumns: int = 1):
        self.strat, self.grid, self.inputs, self.inits = strat, grid, inputs, inits
        self.M = int(ncolumns)
        M, N = self.M, grid.ncells
        layers = self._expand_heterogeneous(strat.layers, grid, strat.z_bot, self.M)
        self._layers = layers
        nl = len(layers)
        # layer -> cells (Numerics/grid.jl:49-58 subgridinds: boundaries snap to the grid edge <= z)
        starts = [int(np.clip(np.searchsorted(grid.edges, z, side="right") - 1, 0, N - 1)) for z, _ in layers]
        starts[0] = 0
        cl = np.zeros(N, dtype=np.int32)
        for l, s in enumerate(starts):
            cl[s:] = l
        self.cell_layer = cl
        self.n_layers = nl
        self.por = np.stack([_col(g.para.por, M) for _, g in layers])
        self.sat = np.stack([_col(g.para.sat, M) for _, g in layers])
        self.org = np.stack([_col(g.para.org, M) for _, g in layers])
        self.fc_kind = np.zeros(nl, dtype=np.int32)
        self.fc_solver = np.zeros(nl, dtype=np.int32)
        self.vg_alpha = np.ones((nl, M)); self.vg_n = np.full((nl, M), 2.0); self.theta_res = np.zeros((nl, M))
        self.Tm = np.zeros((nl, M)); self.beta = np.ones((nl, M)); self.omega = np.ones((nl, M))
        self.fcsolvers = []
        for l, (_, g) in enumerate(layers):
            fc = g.para.freezecurve
            self.fcsolvers.append(g.fcsolver)
            if isinstance(fc, FreeWater):
                continue
            self.fc_kind[l] = fc.kind
            self.fc_solver[l] = L.SOLVER_NEWTON if isinstance(g.fcsolver, SFCCNewtonSolver) else L.SOLVER_LUT
            self.vg_alpha[l] = _col(fc.swrc.alpha, M); self.vg_n[l] = _col(fc.swrc.n, M)
            self.theta_res[l] = _col(fc.swrc.theta_res, M); self.Tm[l] = _col(fc.Tm, M)
            if isinstance(fc, PainterKarra):
                self.beta[l] = _col(fc.beta, M)
                self.omega[l] = _col(fc.omega, M) if fc.omega is not None else 1.0 / self.beta[l]
        heat = layers[0][1].heat
        self.heat = heat
        self.tp = heat.prop
        self.implicit = heat.op == "implicit"
        self.tform = heat.op == "T"           # HeatBalance(:T): the prognostic state is T (heat_conduction.jl:110-116)
        # boundary conditions
        self.top_kind, self.top_forcing, self.use_nfactor, self.nf, self.nt, self.top_offset = self._top(strat.top.bc)
        self.bot_kind, self.bot_forcing, self.bot_value = self._bottom(strat.bottom.bc)
        self.tables = None   # built lazily: list of (Hk, Tk, extrap) + map [nl, M]

    @staticmethod
    def _expand_heterogeneous(layers, grid, z_bot, M):
        out = []
        for li, (z, g) in enumerate(layers):
            if not isinstance(g.para, Heterogeneous):
                out.append((z, g))
                continue
            z_end = layers[li + 1][0] if li + 1 < len(layers) else z_bot
            prof = g.para.profile
            zk = np.array([p[0] for p in prof], dtype=np.float64)
            fc = prof[0][1].freezecurve
            idx = np.nonzero((grid.edges[:-1] >= z - 1e-12) & (grid.edges[:-1] < z_end - 1e-12))[0]
            for i in idx:
                zc = grid.cells[i]
                vals = {}
                for name in ("por", "sat", "org"):
                    vk = np.stack([_col(getattr(p[1], name), M) for p in prof])     # [nk, M]
                    vals[name] = np.array([np.interp(zc, zk, vk[:, c]) for c in range(M)])
                out.append((float(grid.edges[i]), Ground(SimpleSoil(freezecurve=fc, **vals), heat=g.heat, fcsolver=g.fcsolver)))
        return out

    # -- boundary flattening
    def _forcing(self, src):
        if isinstance(src, Input):
            if self.inputs is None:
                raise ValueError(f"Input({src.name}) needs an InputProvider")
            return ("table", self.inputs.ts, self.inputs.data[src.name])
        if isinstance(src, PeriodicBC):
            return ("periodic", src.period, src.amplitude, src.phaseshift, src.level)
        return ("constant", float(src))

    def _top(self, bc):
        M = self.M
        if isinstance(bc, TemperatureBC):
            eff = bc.effect
            src = bc.T
            return (L.BC_DIRICHLET, self._forcing(src), int(eff is not None),
                    _col(eff.nf if eff else 1.0, M), _col(eff.nt if eff else 1.0, M), _col(bc.offset, M))
        if isinstance(bc, ConstantBC):
            kind = L.BC_DIRICHLET if bc.kind == Dirichlet else L.BC_NEUMANN
            v = np.asarray(bc.value, dtype=np.float64)
            if v.ndim == 0:
                return (kind, ("constant", float(v)), 0, np.ones(M), np.ones(M), np.zeros(M))
            return (kind, ("constant", 0.0), 0, np.ones(M), np.ones(M), _col(v, M))
        if isinstance(bc, PeriodicBC):
            kind = L.BC_DIRICHLET if bc.kind == Dirichlet else L.BC_NEUMANN
            return (kind, self._forcing(bc), 0, np.ones(M), np.ones(M), np.zeros(M))
        raise TypeError(f"unsupported upper boundary condition {bc!r}")

    def _bottom(self, bc):
        M = self.M
        if isinstance(bc, GeothermalHeatFlux):
            return (L.BC_NEUMANN, ("constant", 0.0), -_col(bc.Qgeo, M))   # boundaryvalue = -Qgeo (heat_bc.jl:113)
        if isinstance(bc, ConstantBC):
            kind = L.BC_DIRICHLET if bc.kind == Dirichlet else L.BC_NEUMANN
            return (kind, ("constant", 0.0), _col(bc.value, M))
        if isinstance(bc, PeriodicBC):
            kind = L.BC_DIRICHLET if bc.kind == Dirichlet else L.BC_NEUMANN
            return (kind, self._forcing(bc), None)
        raise TypeError(f"unsupported lower boundary condition {bc!r}")

    # -- SFCC tables (one per distinct (layer, parameter class))
    def fc_dict(self, l, c):
        return dict(kind=int(self.fc_kind[l]), alpha=float(self.vg_alpha[l, c]), n=float(self.vg_n[l, c]),
                    theta_res=float(self.theta_res[l, c]), Tm=float(self.Tm[l, c]), beta=float(self.beta[l, c]),
                    omega=float(self.omega[l, c]))

    def build_tables(self):
        if self.tables is not None:
            return self.tables
        tabs, tmap, cache = [], np.zeros((self.n_layers, self.M), dtype=np.int32), {}
        for l in range(self.n_layers):
            if self.fc_kind[l] == L.FC_FREEWATER or self.fc_solver[l] != L.SOLVER_LUT:
                continue
            solver = self.fcsolvers[l]
            for c in range(self.M):
                key = (l if solver.table is not None else -1, float(self.por[l, c]), float(self.sat[l, c]), float(self.org[l, c]),
                       tuple(sorted(self.fc_dict(l, c).items())), solver.Tmin, solver.errtol, solver.Tmax, solver.extrap)
                if key not in cache:
                    if solver.table is not None:
                        Hk, Tk = (np.asarray(a, dtype=np.float64) for a in solver.table)
                    else:
                        thwi = self.por[l, c] * self.sat[l, c]
                        Hk, Tk = hm.presolver_table(self.fc_dict(l, c), self.tp, float(self.por[l, c]),
                                                    float(thwi / self.por[l, c]), float(self.org[l, c]),
                                                    solver.Tmin, solver.errtol, solver.Tmax)
                    cache[key] = len(tabs)
                    tabs.append((Hk, Tk, L.EXTRAP_LINE if solver.extrap == "line" else L.EXTRAP_FLAT))
                    if len(tabs) > 4096:
                        raise ValueError("too many distinct SFCC parameter classes for LUT mode; use SFCCNewtonSolver")
                tmap[l, c] = cache[key]
        self.tables = (tabs, tmap)
        return self.tables

    # -- diagnostics on the host for initial conditions (vectorised over columns)
    def _cellwise(self, a):
        return a[self.cell_layer, :]          # [N, M]

    def freezethaw_host(self, H):
        """H -> (T, θw, C, kc) for all cells/columns (used by ThermalSteadyStateInit only)."""
        tp = self.tp
        por, sat, org = self._cellwise(self.por), self._cellwise(self.sat), self._cellwise(self.org)
        thwi, thm, tho = hm.soil_fractions(por, sat, org)
        T = np.empty_like(H); thw = np.empty_like(H); C = np.empty_like(H)
        kinds = self.fc_kind[self.cell_layer]
        fw = kinds == L.FC_FREEWATER
        if fw.any():
            Hf = H[fw]; th = thwi[fw]
            thtot = np.maximum(th, 1e-8)
            Lth = tp.L * thtot
            with np.errstate(invalid="ignore", over="ignore"):
                w = np.where(Hf > Lth, thtot, np.where(Hf > 0.0, (Hf / Lth) * thtot, 0.0))
                Cf = hm.heatcapacity(tp, w, th, thm[fw], tho[fw])
                Lt = tp.L * th
                Tf = np.where(Hf < 0.0, Hf / Cf, np.where(Hf >= Lt, (Hf - Lt) / Cf, 0.0))
            T[fw], thw[fw], C[fw] = Tf, w, Cf
        if (~fw).any():
            tabs, tmap = self.build_tables()
            for i in np.nonzero(~fw)[0]:
                l = self.cell_layer[i]
                for c in range(self.M):
                    if self.fc_solver[l] == L.SOLVER_LUT:
                        Hk, Tk, ex = tabs[tmap[l, c]]
                        Ti = float(hm.lut_eval(Hk, Tk, H[i, c], ex == L.EXTRAP_LINE))
                    else:
                        Ti = hm.sfcc_newton(self.fc_dict(l, c), tp, self.por[l, c], thwi[i, c] / self.por[l, c], self.org[l, c], H[i, c], 0.0)
                    w, _ = hm.sfcc_theta(int(self.fc_kind[l]), Ti, thwi[i, c] / por[i, c], por[i, c], self.vg_alpha[l, c],
                                         self.vg_n[l, c], self.theta_res[l, c], self.Tm[l, c], self.beta[l, c], self.omega[l, c])
                    T[i, c], thw[i, c] = Ti, w
                    C[i, c] = hm.sfcc_heatcap(tp, por[i, c], org[i, c], w, thwi[i, c], por[i, c])
        kc = hm.thermalconductivity(tp, thw, thwi, thm, tho)
        return T, thw, C, kc

    def enthalpy_from_T(self, T):
        """initialize_soil_heat! (Soils/soil_heat.jl:38-63): θw, C from T; H = T C + L θw.  T is [N, M]."""
        tp = self.tp
        por, sat, org = self._cellwise(self.por), self._cellwise(self.sat), self._cellwise(self.org)
        thwi, thm, tho = hm.soil_fractions(por, sat, org)
        kinds = self.fc_kind[self.cell_layer][:, None] * np.ones((1, self.M), dtype=np.int32)
        thw = np.where(T > 0.0, thwi, 0.0)
        for k in (L.FC_DALLAMICO, L.FC_PAINTERKARRA):
            m = kinds == k
            if m.any():
                a = lambda x: self._cellwise(x)[m]
                w, _ = hm.sfcc_theta(k, T[m], sat[m], por[m], a(self.vg_alpha), a(self.vg_n), a(self.theta_res), a(self.Tm),
                                     a(self.beta), a(self.omega))
                thw[m] = w
        C = hm.heatcapacity(tp, thw, thwi, thm, tho)
        with np.errstate(invalid="ignore", over="ignore"):
            return T * C + tp.L * thw


def SoilHeatTile(heatop, upperbc, lowerbc, soilprofile, inputs, *inits, grid=DefaultGrid_5cm, fcsolver=None,
                 ncolumns=1, heat=None):
    """models.jl:6-24 -- one Ground layer per soil-profile entry, shared heat operator."""
    heat = heat or HeatBalance(op=heatop if heatop in ("H", "T", "implicit") else "H")
    layers = [(z, Ground(para, heat=heat, fcsolver=fcsolver or SFCCPreSolver())) for z, para in soilprofile]
    strat = Stratigraphy((grid.edges[0], Top(upperbc)), layers, (grid.edges[-1], Bottom(lowerbc)))
    return Tile(strat, grid, inputs, *inits, ncolumns=ncolumns)


def initialcondition(tile: Tile, tspan=None):
    """initialcondition!(tile, tspan) (Tiles/tile.jl:187-200): returns u0 = H[N, M]."""
    N, M = tile.grid.ncells, tile.M
    z = tile.grid.cells
    init = tile.inits[0] if tile.inits else ("T", Profile((0.0, 0.0)))
    if isinstance(init, ThermalSteadyStateInit):
        return _steadystate(tile, init)
    _, prof = init
    zk = [p[0] for p in prof]
    vk = np.array([np.broadcast_to(np.asarray(p[1], dtype=np.float64), (M,)) for p in prof])   # [nk, M]
    T = np.empty((N, M))
    if vk.shape[0] == 1:
        T[:] = vk[0][None, :]
    else:
        for c in range(M) if np.ptp(vk, axis=1).any() else [0]:
            T[:, c] = hm.interp_profile(zk, vk[:, c], z)
        if not np.ptp(vk, axis=1).any():
            T[:] = T[:, :1]
    return T if getattr(tile, "tform", False) else tile.enthalpy_from_T(T)


def _steadystate(tile: Tile, init: ThermalSteadyStateInit):
    """Heat/heat_init.jl:58-85, layer by layer with the absolute cell depth z (as in the reference)."""
    N, M = tile.grid.ncells, tile.M
    z = tile.grid.cells[:, None]
    T = np.zeros((N, M)); kc = np.zeros((N, M)); H = np.zeros((N, M))
    Tup = np.full(M, float(init.T0))
    for l in range(tile.n_layers):
        idx = np.nonzero(tile.cell_layer == l)[0]
        if idx.size == 0:
            continue
        sl = slice(idx[0], idx[-1] + 1)
        it, dT = 1, np.inf
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            while it < init.maxiters and dT > init.thresh:
                Tn = Tup[None, :] + z[sl] * init.Qgeo / kc[sl]
                dT = np.max(np.abs(Tn - T[sl]))
                Tfull = T.copy(); Tfull[sl] = Tn
                H[sl] = tile.enthalpy_from_T(Tfull)[sl]
                Tl, _, _, kl = tile.freezethaw_host(H)
                T[sl], kc[sl] = Tl[sl], kl[sl]
                it += 1
            Tn = Tup[None, :] + z[sl] * init.Qgeo / kc[sl]
        Tfull = T.copy(); Tfull[sl] = Tn
        H[sl] = tile.enthalpy_from_T(Tfull)[sl]
        Tl, _, _, kl = tile.freezethaw_host(H)
        T[sl], kc[sl] = Tl[sl], kl[sl]
        Tup = T[idx[-1]].copy()
    return tile.enthalpy_from_T(T)
