"""Coupled heat + salt diffusion on saline sediment columns (reference: src/Physics/Salt, example
examples/heat_sfcc_salt_constantbc.jl).  Host-side mirror: SalineGround / SaltMassBalance / SaltGradient /
SedimentCompactionInitializer, flattened to a batched SalineTile, plus the engine over the C ABI."""
import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

from . import _lib as L
from .engine import _EngineBase
from .grids import Grid, DefaultGrid_10cm
from .hostmath import ThermalProperties
from .model import DallAmicoSalt, VanGenuchten, _col


@dataclass
class SaltProperties:          # salt_types.jl:1-4
    tau: float = 1.5
    ds0: float = 8.0e-10


@dataclass
class SaltMassBalance:         # salt_types.jl:10-14: CFL(0.5, MaxDelta(5e-3))
    prop: SaltProperties = field(default_factory=SaltProperties)
    courant_number: float = 0.5
    dc_max: float = 5.0e-3


@dataclass
class SaltGradient:            # salt_bc.jl:8-11
    benthicSalt: Any = 900.0
    surfaceState: int = 0


@dataclass
class SedimentCompactionInitializer:   # salt_init.jl:1-7
    porosityZero: float = 0.4
    porosityMin: float = 0.03


def compaction(grid: Grid, porosityZero=0.4, porosityMin=0.03):
    """salt_init.jl:9-17 compaction!"""
    z = grid.cells - grid.cells.min()
    bd0 = 1.0 / ((porosityZero + 0.6845) / 1.8)
    bd = bd0 + 0.0037 * z ** 0.766
    return np.maximum(1.80 * bd ** (-1.0) - 0.6845, porosityMin)


class SalineTile:
    """M saline-sediment columns: one SalineGround layer, SaltHeatBC(SaltGradient, upper temperature) on top, GeothermalHeatFlux at
    the bottom.  Every numeric parameter may be a scalar or an array of length `ncolumns` (ensemble members): benthic salt
    concentration, saturation, organic fraction, van Genuchten alpha / n, porosityZero of the compaction initializer.  The upper
    temperature is a constant (ConstantTemperature), a PeriodicBC or a forcing table ("table", ts, ys)."""

    def __init__(self, grid=DefaultGrid_10cm, sfcc=None, sat=1.0, org=0.0, T_ub=0.0, salt_bc=None, Qgeo=0.053,
                 salt=None, porinit=None, T0=-2.0, c0=0.0, ncolumns=1, prop=None):
        self.grid, self.M = grid, int(ncolumns)
        M = self.M
        self.sfcc = sfcc or DallAmicoSalt(swrc=VanGenuchten(alpha=4.06, n=2.03))
        self.sat_c, self.org_c = _col(sat, M), _col(org, M)
        self.alpha_c, self.n_c = _col(self.sfcc.swrc.alpha, M), _col(self.sfcc.swrc.n, M)
        self.thres_c, self.Tm_c = _col(self.sfcc.swrc.theta_res, M), _col(self.sfcc.Tm, M)
        self.Qgeo = float(Qgeo)
        if isinstance(T_ub, tuple):
            self.top_forcing, self.T_ub = T_ub, 0.0
        elif hasattr(T_ub, "period"):       # PeriodicBC
            self.top_forcing, self.T_ub = ("periodic", T_ub.period, T_ub.amplitude, T_ub.phaseshift, T_ub.level), 0.0
        else:
            self.top_forcing, self.T_ub = ("constant", float(T_ub)), float(T_ub)
        self.salt_bc = salt_bc or SaltGradient()
        self.salt = salt or SaltMassBalance()
        self.tp = prop or ThermalProperties()
        pi = porinit or SedimentCompactionInitializer()
        p0 = _col(pi.porosityZero, M)
        self.por_per_column = bool(np.ptp(p0) > 0)
        self.por_cm = np.stack([compaction(grid, float(p0[c]), pi.porosityMin) for c in range(M)], axis=1) if self.por_per_column else None
        self.por_cell = compaction(grid, float(p0[0]), pi.porosityMin)
        self.benthic = _col(self.salt_bc.benthicSalt, self.M)
        self.T0, self.c0 = T0, c0
        self.per_column_soil = any(np.ptp(a) > 0 for a in (self.sat_c, self.org_c, self.alpha_c, self.n_c, self.thres_c, self.Tm_c))

    # scalar views of column 0 (single-class tiles)
    sat = property(lambda self: float(self.sat_c[0]))
    org = property(lambda self: float(self.org_c[0]))

    def por_of(self, c):
        return self.por_cm[:, c] if self.por_per_column else self.por_cell

    def initialcondition(self):
        N, M = self.grid.ncells, self.M
        u = np.empty((2, N, M))
        u[0] = self.T0
        u[1] = self.c0
        return u


class SaltEngine(_EngineBase):
    def __init__(self, tile: SalineTile, dt0=60.0, dtmin=1.0, dtmax=3600.0, device=0, columns=None):
        self.lib = L.load()
        self.tile = tile
        cols = np.arange(tile.M) if columns is None else np.asarray(columns)
        self.columns = cols
        self.N, self.M, self.nvar = tile.grid.ncells, int(cols.size), 2
        cfg = L.Config(abi_version=L.ABI_VERSION, device=device, n_cells=self.N, n_layers=1, n_columns=self.M,
                       physics=L.PHYSICS_HEAT_SALT, stepper=L.STEPPER_EULER, limiter=L.LIMITER_CFL, top_kind=L.BC_DIRICHLET,
                       bot_kind=L.BC_NEUMANN, use_nfactor=0, maxdelta=float("inf"), courant=float(tile.salt.courant_number),
                       dt0=float(dt0), dtmin=float(dtmin), dtmax=float(dtmax), lite_miniters=2, lite_maxiters=1000, lite_tol=1e-3)
        h = C.c_void_p()
        rc = self.lib.cgb_create(C.byref(cfg), C.byref(h))
        if rc != L.OK:
            raise L.CgbError(rc, self.lib.cgb_last_error(None).decode())
        self.h = h
        e = L.as_f64(tile.grid.edges)
        self._check(self.lib.cgb_set_grid(self.h, L.dptr(e)))
        if tile.por_per_column:
            self._param("por", np.ascontiguousarray(tile.por_cm[:, cols]), L.PER_CELL_COLUMN)
        else:
            self._param("por", tile.por_cell, L.PER_CELL)
        for name, arr in (("sat", tile.sat_c), ("org", tile.org_c), ("vg_alpha", tile.alpha_c), ("vg_n", tile.n_c),
                          ("theta_res", tile.thres_c), ("Tm", tile.Tm_c)):
            if np.ptp(arr) > 0:
                self._param(name, arr[cols], L.PER_COLUMN)
            else:
                self._param(name, [float(arr[0])], L.PER_GLOBAL)
        for name, v in (("tau", tile.salt.prop.tau), ("ds0", tile.salt.prop.ds0), ("salt_dc_max", tile.salt.dc_max),
                        ("salt_surface_state", float(tile.salt_bc.surfaceState))):
            self._param(name, [float(v)], L.PER_GLOBAL)
        for k, v in tile.tp.as_dict().items():
            self._param(k, [v], L.PER_GLOBAL)
        self._param("benthic_salt", tile.benthic[cols], L.PER_COLUMN)
        self._param("bot_value", np.full(self.M, -tile.Qgeo), L.PER_COLUMN)
        self._check(self.lib.cgb_set_freezecurve(self.h, 0, L.FC_DALLAMICOSALT, L.SOLVER_NEWTON))
        f = tile.top_forcing
        if f[0] == "table":
            ts, ys = L.as_f64(f[1]), L.as_f64(f[2])
            self._check(self.lib.cgb_set_forcing_table(self.h, L.TOP, L.dptr(ts), L.dptr(ys), ts.size))
        elif f[0] == "periodic":
            self._check(self.lib.cgb_set_forcing_periodic(self.h, L.TOP, *[float(x) for x in f[1:]]))
        else:
            self._check(self.lib.cgb_set_forcing_constant(self.h, L.TOP, tile.T_ub))

    def set_state(self, u, t0):
        u = L.as_f64(u)
        if u.shape != (2, self.N, self.M):
            raise ValueError(f"state must have shape {(2, self.N, self.M)}")
        self._check(self.lib.cgb_set_state(self.h, L.dptr(u), float(t0)))

    def get_state(self, with_time=False):
        u = np.empty((2, self.N, self.M))
        if not with_time:
            self._check(self.lib.cgb_get_state(self.h, L.dptr(u), None, None))
            return u
        t = np.empty(self.M); dt = np.empty(self.M)
        self._check(self.lib.cgb_get_state(self.h, L.dptr(u), L.dptr(t), L.dptr(dt)))
        return u, t, dt

    def rhs(self, t):
        du = np.empty((2, self.N, self.M))
        self._check(self.lib.cgb_rhs(self.h, float(t), L.dptr(du)))
        return du

    def solve_saveat(self, tsave, savevars=("T", "c", "thw")):
        """out[nsave, nvars, N, M]: the prognostic T, c and cell diagnostics (thw, C, dHdT, kc, H, dT, dc ...) at every save time."""
        tsave = L.as_f64(tsave)
        out = np.empty((tsave.size, len(savevars), self.N, self.M))
        self._check(self.lib.cgb_solve_saveat(self.h, L.dptr(tsave), tsave.size, ",".join(savevars).encode(), L.dptr(out)))
        return out

    def diag(self, var, t):
        n = self.N + 1 if var in ("k", "jH", "ds", "jc") else self.N
        out = np.empty((n, self.M))
        self._check(self.lib.cgb_get_diag(self.h, var.encode(), float(t), L.dptr(out)))
        return out
