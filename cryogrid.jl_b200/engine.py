"""Engine: owns one libcgb200 handle (one GPU, one shard of columns) for a batched Tile."""
import ctypes as C

import numpy as np

from . import _lib as L


# Saved-variable names.  "T"/"θw" follow the reference verbatim: its savefunc returns the diagnostic CACHE of the last RHS
# evaluation, i.e. the state one step before the saved H (include/cgb200.h, cgb_solve_saveat); "*_now" are recomputed from
# the saved state itself.
SAVE_NAMES = {"T": "T", "H": "H", "θw": "thw", "thw": "thw", "T_now": "T_now", "θw_now": "thw_now", "thw_now": "thw_now"}


class _EngineBase:
    """Methods shared by the heat and heat+salt engines (everything that does not depend on the state shape)."""

    def _check(self, rc):
        if rc != L.OK:
            raise L.CgbError(rc, self.lib.cgb_last_error(self.h).decode())

    def _param(self, name, arr, per):
        a = L.as_f64(arr)
        self._check(self.lib.cgb_set_param(self.h, name.encode(), L.dptr(a), a.size, per))

    def status(self):
        st = np.empty(self.M, dtype=np.int32); steps = np.empty(self.M, dtype=np.int64)
        self._check(self.lib.cgb_get_status(self.h, st.ctypes.data_as(C.POINTER(C.c_int32)), steps.ctypes.data_as(C.POINTER(C.c_int64))))
        return st, steps

    def stats(self):
        s = L.Stats()
        self._check(self.lib.cgb_get_stats(self.h, C.byref(s)))
        return {f[0]: getattr(s, f[0]) for f in L.Stats._fields_ if f[0] != "_pad"}

    def restore_state(self, u, t, dt, steps=None):
        """Checkpoint restore (counterpart of get_state(with_time=True)): per-column time and next step size."""
        u = L.as_f64(u)
        if u.size != getattr(self, "nvar", 1) * self.N * self.M:
            raise ValueError(f"state has {u.size} elements, expected {getattr(self, 'nvar', 1) * self.N * self.M}")
        t = L.as_f64(np.broadcast_to(t, (self.M,))); dt = L.as_f64(np.broadcast_to(dt, (self.M,)))
        st = None if steps is None else np.ascontiguousarray(steps, dtype=np.int64)
        self._check(self.lib.cgb_restore_state(self.h, L.dptr(u), L.dptr(t), L.dptr(dt),
                                               None if st is None else st.ctypes.data_as(C.POINTER(C.c_int64))))

    def advance_to(self, t_target, sync=True):
        self._check(self.lib.cgb_advance_to(self.h, float(t_target)))
        if sync:
            self._check(self.lib.cgb_synchronize(self.h))

    def synchronize(self):
        self._check(self.lib.cgb_synchronize(self.h))

    def device_pointers(self):
        u, t, dt = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.cgb_device_state(self.h, C.byref(u), C.byref(t), C.byref(dt)))
        return u.value, t.value, dt.value

    def profile_enable(self, on=True):
        self._check(self.lib.cgb_profile_enable(self.h, int(on)))

    def profile_read(self):
        n = C.c_int64(0); ms = C.c_double(0.0)
        self._check(self.lib.cgb_profile_read(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.cgb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine(_EngineBase):
    """Thin, explicit wrapper over the C ABI. All arrays crossing it are host NumPy arrays, column-fastest."""

    def __init__(self, tile, stepper="euler", dt0=None, dtmin=1.0, dtmax=None, device=0,
                 miniters=2, maxiters=1000, tolerance=1e-3, columns=None, device_tables=None):
        """device_tables: build the SFCC presolver tables on the device (cgb_build_sfcc_tables) instead of the host loop of
        Tile.build_tables(); default (None): when the tile has more than 16 distinct soil classes and no caller-built table."""
        self.device_tables = device_tables
        self.lib = L.load()
        self.tile = tile
        cols = np.arange(tile.M) if columns is None else np.asarray(columns)
        self.columns = cols
        M = int(cols.size)
        N = tile.grid.ncells
        self.N, self.M = N, M
        lite = stepper in ("lite", "implicit", L.STEPPER_LITE_IMPLICIT)
        if lite != bool(tile.implicit):
            raise ValueError("LiteImplicitEuler needs HeatBalance(EnthalpyImplicit()) and CGEuler needs Diffusion1D(:H)")
        if dt0 is None:
            dt0 = 86400.0 if lite else 60.0          # cglite_types.jl:29 ; basic_solvers.jl:22
        if dtmax is None:
            dtmax = dt0 if lite else 3600.0          # cglite_types.jl:30 ; integrator.jl:55
        lim = tile.heat.dtlim
        is_cfl = hasattr(lim, "courant_number")
        cfg = L.Config(
            abi_version=L.ABI_VERSION, device=device, n_cells=N, n_layers=tile.n_layers, n_columns=M,
            physics=L.PHYSICS_HEAT_T if getattr(tile, "tform", False) else L.PHYSICS_HEAT, stepper=L.STEPPER_LITE_IMPLICIT if lite else L.STEPPER_EULER,
            limiter=L.LIMITER_CFL if is_cfl else L.LIMITER_MAXDELTA, top_kind=tile.top_kind, bot_kind=tile.bot_kind,
            use_nfactor=tile.use_nfactor,
            maxdelta=float(lim.maxdelta.dmax if is_cfl else lim.dmax), courant=float(lim.courant_number if is_cfl else 0.5),
            dt0=float(dt0), dtmin=float(dtmin), dtmax=float(dtmax),
            lite_miniters=int(miniters), lite_maxiters=int(maxiters), lite_tol=float(tolerance),
        )
        self.cfg = cfg
        h = C.c_void_p()
        rc = self.lib.cgb_create(C.byref(cfg), C.byref(h))
        if rc != L.OK:
            raise L.CgbError(rc, self.lib.cgb_last_error(None).decode())
        self.h = h
        self._push_description()

    # ---- helpers
    def _push_description(self):
        t, cols = self.tile, self.columns
        e = L.as_f64(t.grid.edges)
        self._check(self.lib.cgb_set_grid(self.h, L.dptr(e)))
        cl = np.ascontiguousarray(t.cell_layer, dtype=np.int32)
        self._check(self.lib.cgb_set_layers(self.h, cl.ctypes.data_as(C.POINTER(C.c_int32))))
        for name in ("por", "sat", "org", "vg_alpha", "vg_n", "theta_res", "Tm", "beta", "omega"):
            self._param(name, getattr(t, name)[:, cols], L.PER_COLUMN_LAYER)
        for k, v in t.tp.as_dict().items():
            self._param(k, [v], L.PER_GLOBAL)
        self._param("nf", t.nf[cols], L.PER_COLUMN)
        self._param("nt", t.nt[cols], L.PER_COLUMN)
        self._param("top_offset", t.top_offset[cols], L.PER_COLUMN)
        if t.bot_value is not None:
            self._param("bot_value", t.bot_value[cols], L.PER_COLUMN)
        for which, f in ((L.TOP, t.top_forcing), (L.BOTTOM, t.bot_forcing)):
            if f[0] == "table":
                ts, ys = L.as_f64(f[1]), L.as_f64(f[2])
                self._check(self.lib.cgb_set_forcing_table(self.h, which, L.dptr(ts), L.dptr(ys), ts.size))
            elif f[0] == "periodic":
                self._check(self.lib.cgb_set_forcing_periodic(self.h, which, *[float(x) for x in f[1:]]))
            else:
                self._check(self.lib.cgb_set_forcing_constant(self.h, which, float(f[1])))
        any_lut = False
        for l in range(t.n_layers):
            self._check(self.lib.cgb_set_freezecurve(self.h, l, int(t.fc_kind[l]), int(t.fc_solver[l])))
            any_lut |= t.fc_kind[l] != L.FC_FREEWATER and t.fc_solver[l] == L.SOLVER_LUT and not getattr(t, "tform", False)
        if any_lut and self._use_device_tables():
            sv = next(s for l, s in enumerate(t.fcsolvers) if t.fc_kind[l] != L.FC_FREEWATER and t.fc_solver[l] == L.SOLVER_LUT)
            n = C.c_int32(0)
            self._check(self.lib.cgb_build_sfcc_tables(self.h, float(sv.Tmin), float(sv.errtol), float(sv.Tmax),
                                                       L.EXTRAP_LINE if sv.extrap == "line" else L.EXTRAP_FLAT, C.byref(n)))
            self.n_device_tables = n.value
        elif any_lut:
            tabs, tmap = t.build_tables()
            for i, (Hk, Tk, ex) in enumerate(tabs):
                Hk, Tk = L.as_f64(Hk), L.as_f64(Tk)
                self._check(self.lib.cgb_set_sfcc_table(self.h, i, L.dptr(Hk), L.dptr(Tk), Hk.size, ex))
            m = np.ascontiguousarray(tmap[:, cols], dtype=np.int32)
            self._check(self.lib.cgb_set_sfcc_table_map(self.h, m.ctypes.data_as(C.POINTER(C.c_int32))))

    def _use_device_tables(self):
        t = self.tile
        if any(getattr(s, "table", None) is not None for s in t.fcsolvers):
            return False                                    # caller-built knots (e.g. sampled from the Julia interpolant) win
        if self.device_tables is not None:
            return bool(self.device_tables)
        varying = sum(int(np.ptp(getattr(t, name)[:, self.columns], axis=1).any()) for name in ("por", "sat", "org", "vg_alpha", "vg_n"))
        return varying > 0 and self.M > 16

    def get_sfcc_table(self, table_id):
        nk = C.c_int64(0)
        self._check(self.lib.cgb_get_sfcc_table(self.h, int(table_id), None, None, 0, C.byref(nk)))
        Hk = np.empty(nk.value); Tk = np.empty(nk.value)
        self._check(self.lib.cgb_get_sfcc_table(self.h, int(table_id), L.dptr(Hk), L.dptr(Tk), nk.value, C.byref(nk)))
        return Hk, Tk

    # ---- state
    def set_state(self, u, t0):
        u = L.as_f64(u)
        if u.shape != (self.N, self.M):
            raise ValueError(f"state must have shape {(self.N, self.M)}, got {u.shape}")
        self._check(self.lib.cgb_set_state(self.h, L.dptr(u), float(t0)))

    def initialcondition(self, t0=0.0):
        """initialcondition!(tile, tspan) on the device, from the tile's initializer (InterpInitializer on :T or
        ThermalSteadyStateInit).  Equivalent to set_state(model.initialcondition(tile), t0) without the host loop."""
        from .model import ThermalSteadyStateInit, Profile
        init = self.tile.inits[0] if self.tile.inits else ("T", Profile((0.0, 0.0)))
        if isinstance(init, ThermalSteadyStateInit):
            T0 = L.as_f64(np.broadcast_to(np.asarray(init.T0, dtype=np.float64), (self.tile.M,))[self.columns]
                          if np.ndim(init.T0) else [float(init.T0)])
            per = L.PER_COLUMN if np.ndim(init.T0) else L.PER_GLOBAL
            self._check(self.lib.cgb_init_steadystate(self.h, L.dptr(T0), per, float(init.Qgeo), int(init.maxiters), float(init.thresh), float(t0)))
            return
        _, prof = init
        zk = L.as_f64([p[0] for p in prof])
        percol = any(np.ndim(p[1]) for p in prof)
        if percol:
            vk = L.as_f64(np.stack([np.broadcast_to(np.asarray(p[1], dtype=np.float64), (self.tile.M,))[self.columns] for p in prof]))
        else:
            vk = L.as_f64([float(p[1]) for p in prof])
        self._check(self.lib.cgb_init_interp(self.h, L.dptr(zk), L.dptr(vk), zk.size, L.PER_COLUMN if percol else L.PER_GLOBAL, float(t0)))

    def get_state(self, with_time=False):
        u = np.empty((self.N, self.M))
        if not with_time:
            self._check(self.lib.cgb_get_state(self.h, L.dptr(u), None, None))
            return u
        t = np.empty(self.M); dt = np.empty(self.M)
        self._check(self.lib.cgb_get_state(self.h, L.dptr(u), L.dptr(t), L.dptr(dt)))
        return u, t, dt

    # ---- hot path
    def rhs(self, t):
        du = np.empty((self.N, self.M))
        self._check(self.lib.cgb_rhs(self.h, float(t), L.dptr(du)))
        return du

    def diag(self, var, t, out=None):
        n = self.N + 1 if var in ("k", "jH") else self.N
        if out is None:
            out = np.empty((n, self.M))
        assert out.shape == (n, self.M) and out.flags.c_contiguous and out.dtype == np.float64
        self._check(self.lib.cgb_get_diag(self.h, var.encode(), float(t), L.dptr(out)))
        return out

    def solve_saveat(self, tsave, savevars=("T",), out=None):
        tsave = L.as_f64(tsave)
        names = SAVE_NAMES
        vars_c = ",".join(names[v] for v in savevars)
        shape = (tsave.size, len(savevars), self.N, self.M)
        if out is None:
            out = np.empty(shape)
        assert out.shape == shape and out.flags.c_contiguous and out.dtype == np.float64
        self._check(self.lib.cgb_solve_saveat(self.h, L.dptr(tsave), tsave.size, vars_c.encode(), L.dptr(out)))
        return out

    def solve_saveat_select(self, tsave, savevars=("T",), cells=(), thawdepth=False, out=None, thaw_out=None):
        """Device-side output path: keep only the cells `cells` of every saved variable (out[nsave, nvars, nsel, M]) and,
        optionally, the sub-grid thaw depth per column and save (thaw[nsave, M])."""
        tsave = L.as_f64(tsave)
        names = SAVE_NAMES
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        vars_c = ",".join(names[v] for v in savevars) if cells.size else "T"
        shape = (tsave.size, len(savevars), cells.size, self.M)
        if out is None and cells.size:
            out = np.empty(shape)
        if thawdepth and thaw_out is None:
            thaw_out = np.empty((tsave.size, self.M))
        self._check(self.lib.cgb_solve_saveat_select(
            self.h, L.dptr(tsave), tsave.size, vars_c.encode(), cells.ctypes.data_as(C.POINTER(C.c_int32)), int(cells.size),
            L.dptr(out) if cells.size else None, L.dptr(thaw_out) if thawdepth else None))
        return out, thaw_out

    def solve_saveat_permafrost(self, tsave, Tmelt=0.0, zaa_threshold=0.5, magt_upper=0.0, magt_lower=10.0):
        """Integrate through `tsave` and reduce the saved T on the device to the permafrost diagnostics of Diagnostics/permafrost.jl:
        returns dict(alt, pf_table, pf_base, zaa [M each], magt [nsave, M])."""
        tsave = L.as_f64(tsave)
        out = {k: np.empty(self.M) for k in ("alt", "pf_table", "pf_base", "zaa")}
        out["magt"] = np.empty((tsave.size, self.M))
        self._check(self.lib.cgb_solve_saveat_permafrost(self.h, L.dptr(tsave), tsave.size, float(Tmelt), float(zaa_threshold), float(magt_upper),
                                                         float(magt_lower), L.dptr(out["alt"]), L.dptr(out["pf_table"]), L.dptr(out["pf_base"]),
                                                         L.dptr(out["zaa"]), L.dptr(out["magt"])))
        return out

    def selftest_sfcc_curve(self, T):
        """θw(T) from the curve table of the SFCC single-class fast path (raises CgbError(UNSUPPORTED) if the tile does not qualify)."""
        T = L.as_f64(T); out = np.empty_like(T)
        self._check(self.lib.cgb_selftest_sfcc_curve(self.h, L.dptr(T), T.size, L.dptr(out)))
        return out

    def ensemble_partial(self, var, t):
        s = np.empty(self.N); q = np.empty(self.N)
        self._check(self.lib.cgb_ensemble_partial(self.h, var.encode(), float(t), L.dptr(s), L.dptr(q)))
        return s, q

    def ensemble_partial_device(self, var, t, out_ptr):
        self._check(self.lib.cgb_ensemble_partial_device(self.h, var.encode(), float(t), C.c_void_p(out_ptr)))
