"""ctypes binding of oracle/libcg_oracle.so -- the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
It consumes the same batched `Tile` description as the CUDA engine and evaluates ONE column at a time with
the plain-C restatement of the reference (oracle/cg_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcg_oracle.so")
FAST_LIB_PATH = os.path.join(HERE, "libcg_oracle_fast.so")   # -O3 -march=native timing copy (bench.py cpu legs only)

_D = C.POINTER(C.c_double)
_I32 = C.POINTER(C.c_int32)
_I64 = C.POINTER(C.c_int64)


class Thermal(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("kh_w", "kh_i", "kh_a", "kh_m", "kh_o", "ch_w", "ch_i", "ch_a", "ch_m", "ch_o", "L")]


class FreezeCurve(C.Structure):
    _fields_ = [("kind", C.c_int32), ("solver", C.c_int32), ("alpha", C.c_double), ("n", C.c_double), ("theta_res", C.c_double),
                ("Tm", C.c_double), ("beta", C.c_double), ("omega", C.c_double), ("g", C.c_double), ("Lsl", C.c_double),
                ("R", C.c_double), ("rho_w", C.c_double), ("nk", C.c_int32), ("extrap", C.c_int32), ("Hk", _D), ("Tk", _D)]


class Forcing(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("value", C.c_double), ("period", C.c_double), ("amplitude", C.c_double),
                ("phase", C.c_double), ("level", C.c_double), ("t", _D), ("y", _D)]


class Column(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_layers", C.c_int32), ("edges", _D), ("cell_layer", _I32), ("por", _D), ("sat", _D),
                ("org", _D), ("fc", C.POINTER(FreezeCurve)), ("tp", Thermal), ("op", C.c_int32), ("top_kind", C.c_int32),
                ("use_nfactor", C.c_int32), ("bot_kind", C.c_int32), ("top", Forcing), ("bot", Forcing), ("nf", C.c_double),
                ("nt", C.c_double), ("limiter", C.c_int32), ("_pad", C.c_int32), ("maxdelta", C.c_double), ("courant", C.c_double)]


class Work(C.Structure):
    _names = ("T", "thw", "C", "dHdT", "dthwdT", "kc", "k", "jH", "dH", "an", "as_", "ap", "bp", "zc", "dk", "dxc", "dT")
    _fields_ = [(n, _D) for n in _names]


class SaltColumn(C.Structure):
    _fields_ = [("n", C.c_int32), ("surface_state", C.c_int32), ("edges", _D), ("por", _D), ("sat", C.c_double), ("org", C.c_double),
                ("fc", FreezeCurve), ("tp", Thermal), ("tau", C.c_double), ("ds0", C.c_double), ("T_ub", C.c_double),
                ("benthic_salt", C.c_double), ("bot_value", C.c_double), ("courant", C.c_double), ("dc_max", C.c_double),
                ("heat_courant", C.c_double), ("use_top_forcing", C.c_int32), ("_pad", C.c_int32), ("top", Forcing)]


class SaltWork(C.Structure):
    _names = ("thw", "dthwdT", "dthwdc", "C", "dHdT", "H", "kc", "ds_mid", "k", "ds", "jH", "jc", "dH", "dc_F", "dT", "dc", "zc", "dk", "dxc")
    _fields_ = [(n, _D) for n in _names]


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("cg_oracle.c", "cg_oracle.h", "Makefile")]
    stale = lambda p: not os.path.exists(p) or os.path.getmtime(p) < max(os.path.getmtime(s) for s in src)
    if force or stale(LIB_PATH) or stale(FAST_LIB_PATH):
        r = subprocess.run(["make", "-C", HERE, "-B"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None
_use_fast = False


def use_timing_copy(on=True):
    """bench.py only: bind the -O3 -march=native timing copy instead of the parity oracle (must be called before lib())."""
    global _use_fast, _lib
    _use_fast, _lib = bool(on), None


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(FAST_LIB_PATH if _use_fast else LIB_PATH)
        l.cg_flux.restype = C.c_double; l.cg_flux.argtypes = [C.c_double] * 4
        l.cg_divergence.restype = C.c_double; l.cg_divergence.argtypes = [C.c_double] * 3
        l.cg_harmonicmean.restype = C.c_double; l.cg_harmonicmean.argtypes = [C.c_double] * 4
        l.cg_grid.argtypes = [_D, C.c_int, _D, _D, _D]
        l.cg_flux_vec.argtypes = [_D, _D, _D, _D, C.c_int]
        l.cg_divergence_vec.argtypes = [_D, _D, _D, C.c_int]
        l.cg_nonlineardiffusion.argtypes = [_D, _D, _D, _D, _D, _D, C.c_int]
        l.cg_tdma_solve.argtypes = [_D, _D, _D, _D, _D, C.c_int]
        l.cg_freewater.argtypes = [C.c_double, C.c_double, C.c_double, _D]
        l.cg_enthalpyinv_freewater.restype = C.c_double; l.cg_enthalpyinv_freewater.argtypes = [C.c_double] * 4
        l.cg_vg_theta.restype = C.c_double; l.cg_vg_theta.argtypes = [C.c_double] * 5 + [_D]
        l.cg_vg_psi.restype = C.c_double; l.cg_vg_psi.argtypes = [C.c_double] * 5
        l.cg_sfcc.restype = C.c_double; l.cg_sfcc.argtypes = [C.POINTER(FreezeCurve)] + [C.c_double] * 4 + [_D, _D]
        l.cg_sfcc_heatcap.restype = C.c_double; l.cg_sfcc_heatcap.argtypes = [C.POINTER(Thermal)] + [C.c_double] * 5
        l.cg_presolver_build.restype = C.c_int
        l.cg_presolver_build.argtypes = [C.POINTER(FreezeCurve), C.POINTER(Thermal)] + [C.c_double] * 6 + [_D, _D, C.c_int]
        l.cg_lut_eval.restype = C.c_double; l.cg_lut_eval.argtypes = [_D, _D, C.c_int, C.c_int, C.c_double]
        l.cg_sfcc_newton.restype = C.c_double
        l.cg_sfcc_newton.argtypes = [C.POINTER(FreezeCurve), C.POINTER(Thermal)] + [C.c_double] * 5
        l.cg_forcing_eval.restype = C.c_double; l.cg_forcing_eval.argtypes = [C.POINTER(Forcing), C.c_double, C.POINTER(C.c_int)]
        l.cg_nfactor.restype = C.c_double; l.cg_nfactor.argtypes = [C.c_double] * 3
        l.cg_rhs.restype = C.c_int; l.cg_rhs.argtypes = [C.POINTER(Column), _D, C.c_double, C.POINTER(Work)]
        l.cg_timestep.restype = C.c_double; l.cg_timestep.argtypes = [C.POINTER(Column), _D, C.POINTER(Work)]
        l.cg_cgeuler_advance.restype = C.c_int
        l.cg_cgeuler_advance.argtypes = [C.POINTER(Column), _D, _D, _D, C.c_double, C.c_double, C.c_double, _I64, C.POINTER(Work)]
        l.cg_cgeuler_advance_trace.restype = C.c_int
        l.cg_cgeuler_advance_trace.argtypes = [C.POINTER(Column), _D, _D, _D, C.c_double, C.c_double, C.c_double, _I64, C.POINTER(Work),
                                               _D, C.c_int64, _I64]
        l.cg_salt_advance_trace.restype = C.c_int
        l.cg_salt_advance_trace.argtypes = [C.POINTER(SaltColumn), _D, _D, _D, _D, C.c_double, C.c_double, C.c_double, _I64,
                                            C.POINTER(SaltWork), _D, C.c_int64]
        l.cg_lite_step.restype = C.c_int
        l.cg_lite_step.argtypes = [C.POINTER(Column), _D, _D, C.c_double, C.c_int, C.c_int, C.c_double, _I64, C.POINTER(Work)]
        l.cg_lite_advance.restype = C.c_int
        l.cg_lite_advance.argtypes = [C.POINTER(Column), _D, _D, _D, C.c_double, C.c_int, C.c_int, C.c_double, _I64, _I64, C.POINTER(Work)]
        l.cg_interp_profile.argtypes = [_D, _D, C.c_int, _D, C.c_int, _D]
        l.cg_init_soil_heat.argtypes = [C.POINTER(Column), _D, _D, C.POINTER(Work)]
        l.cg_steadystate_init.argtypes = [C.POINTER(Column), C.c_double, C.c_double, C.c_int, C.c_double, _D, C.POINTER(Work)]
        l.cg_compaction.argtypes = [_D, C.c_int, C.c_double, C.c_double, _D]
        l.cg_salt_rhs.argtypes = [C.POINTER(SaltColumn), _D, _D, C.POINTER(SaltWork)]
        l.cg_salt_rhs_t.restype = C.c_int; l.cg_salt_rhs_t.argtypes = [C.POINTER(SaltColumn), _D, _D, C.c_double, C.POINTER(SaltWork)]
        l.cg_salt_timestep.restype = C.c_double; l.cg_salt_timestep.argtypes = [C.POINTER(SaltColumn), C.POINTER(SaltWork)]
        l.cg_salt_advance.restype = C.c_int
        l.cg_salt_advance.argtypes = [C.POINTER(SaltColumn), _D, _D, _D, _D, C.c_double, C.c_double, C.c_double, _I64, C.POINTER(SaltWork)]
        l.cg_cgeuler_advance_batch.restype = C.c_int
        l.cg_cgeuler_advance_batch.argtypes = [C.POINTER(Column), C.c_int64, _D, _D, _D, C.c_double, C.c_double, C.c_double, _I64, C.c_int]
        l.cg_lite_advance_batch.restype = C.c_int
        l.cg_lite_advance_batch.argtypes = [C.POINTER(Column), C.c_int64, _D, _D, _D, C.c_double, C.c_int, C.c_int, C.c_double, _I64, _I64, C.c_int]
        l.cg_max_threads.restype = C.c_int
        _lib = l
    return _lib


def dp(a):
    return a.ctypes.data_as(_D)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleWork:
    def __init__(self, n):
        self.n = n
        self.arrays = {}
        self.c = Work()
        for name in Work._names:
            a = np.zeros(n + 1)
            self.arrays[name] = a
            setattr(self.c, name, dp(a))

    def get(self, name):
        name = {"as": "as_"}.get(name, name)
        m = self.n + 1 if name in ("k", "jH") else (self.n - 1 if name == "dxc" else self.n)
        return self.arrays[name][:m].copy()


def thermal_struct(tp):
    return Thermal(**tp.as_dict())


def fc_struct(d, solver=0, table=None, extrap=1, keep=None):
    fc = FreezeCurve(kind=d["kind"], solver=solver, alpha=d["alpha"], n=d["n"], theta_res=d["theta_res"], Tm=d["Tm"],
                     beta=d["beta"], omega=d["omega"], g=9.80665, Lsl=3.34e5, R=8.314459, rho_w=1000.0, nk=0, extrap=extrap)
    if table is not None:
        Hk, Tk = f64(table[0]), f64(table[1])
        if keep is not None:
            keep.extend([Hk, Tk])
        fc.nk = Hk.size; fc.Hk = dp(Hk); fc.Tk = dp(Tk)
    return fc


def forcing_struct(f, keep):
    if f[0] == "table":
        t, y = f64(f[1]), f64(f[2])
        keep.extend([t, y])
        return Forcing(kind=2, n=t.size, t=dp(t), y=dp(y))
    if f[0] == "periodic":
        return Forcing(kind=1, period=f[1], amplitude=f[2], phase=f[3], level=f[4])
    return Forcing(kind=0, value=f[1])


class OracleColumn:
    """One column (index c) of a batched Tile, as the oracle's cg_column."""

    def __init__(self, tile, c=0, oracle_tables=True):
        self.tile, self.c = tile, c
        self.keep = []
        n, nl = tile.grid.ncells, tile.n_layers
        self.n = n
        edges = f64(tile.grid.edges); cl = np.ascontiguousarray(tile.cell_layer, dtype=np.int32)
        por, sat, org = f64(tile.por[:, c]), f64(tile.sat[:, c]), f64(tile.org[:, c])
        self.keep += [edges, cl, por, sat, org]
        fcs = (FreezeCurve * nl)()
        tform = getattr(tile, "tform", False)     # temperature form: no inverse-enthalpy solver, no tables
        tabs = tile.build_tables() if (not tform and any(tile.fc_kind[l] != 0 and tile.fc_solver[l] == 0 for l in range(nl))) else None
        for l in range(nl):
            table, ex = None, 1
            if tabs is not None and tile.fc_kind[l] != 0 and tile.fc_solver[l] == 0:
                Hk, Tk, ex = tabs[0][tabs[1][l, c]]
                table = (Hk, Tk)
            fcs[l] = fc_struct(tile.fc_dict(l, c), int(tile.fc_solver[l]), table, ex, self.keep)
        self.fcs = fcs
        lim = tile.heat.dtlim
        is_cfl = hasattr(lim, "courant_number")
        top = forcing_struct(tile.top_forcing, self.keep)
        # the per-column top offset is folded into the forcing itself
        off = float(tile.top_offset[c])
        if off != 0.0:
            if top.kind == 0:
                top.value += off
            elif top.kind == 1:
                top.level += off
            else:
                y = f64(tile.top_forcing[2]) + off
                self.keep.append(y); top.y = dp(y)
        if tile.bot_value is not None:
            bot = Forcing(kind=0, value=float(tile.bot_value[c]))
        else:
            bot = forcing_struct(tile.bot_forcing, self.keep)
        self.col = Column(
            n=n, n_layers=nl, edges=dp(edges), cell_layer=cl.ctypes.data_as(_I32), por=dp(por), sat=dp(sat), org=dp(org),
            fc=C.cast(fcs, C.POINTER(FreezeCurve)), tp=thermal_struct(tile.tp), op=1 if tile.implicit else (2 if getattr(tile, "tform", False) else 0),
            top_kind=tile.top_kind, use_nfactor=tile.use_nfactor, bot_kind=tile.bot_kind, top=top, bot=bot,
            nf=float(tile.nf[c]), nt=float(tile.nt[c]), limiter=1 if is_cfl else 0,
            maxdelta=float(lim.maxdelta.dmax if is_cfl else lim.dmax), courant=float(lim.courant_number if is_cfl else 0.5),
        )
        self.work = OracleWork(n)

    def rhs(self, H, t):
        H = f64(H)
        st = lib().cg_rhs(C.byref(self.col), dp(H), float(t), C.byref(self.work.c))
        return self.work.get("dT" if self.col.op == 2 else "dH"), st

    def timestep(self, H):
        H = f64(H)
        return lib().cg_timestep(C.byref(self.col), dp(H), C.byref(self.work.c))

    def advance(self, H, t, dt, t_target, dtmin=1.0, dtmax=3600.0):
        H = f64(H).copy(); tt = np.array([t], dtype=np.float64); dd = np.array([dt], dtype=np.float64)
        ns = np.zeros(1, dtype=np.int64)
        st = lib().cg_cgeuler_advance(C.byref(self.col), dp(H), dp(tt), dp(dd), float(t_target), float(dtmin), float(dtmax),
                                      ns.ctypes.data_as(_I64), C.byref(self.work.c))
        return H, tt[0], dd[0], int(ns[0]), st

    def advance_trace(self, H, t, dt, t_target, dtmin=1.0, dtmax=3600.0, cap=1 << 20):
        """Like advance, plus the dt every step was taken with and the number of steps whose next dt the limiter decided."""
        H = f64(H).copy(); tt = np.array([t], dtype=np.float64); dd = np.array([dt], dtype=np.float64)
        ns = np.zeros(1, dtype=np.int64); nl = np.zeros(1, dtype=np.int64)
        tr = np.zeros(cap)
        st = lib().cg_cgeuler_advance_trace(C.byref(self.col), dp(H), dp(tt), dp(dd), float(t_target), float(dtmin), float(dtmax),
                                            ns.ctypes.data_as(_I64), C.byref(self.work.c), dp(tr), cap, nl.ctypes.data_as(_I64))
        return H, tt[0], dd[0], int(ns[0]), st, tr[:min(int(ns[0]), cap)].copy(), int(nl[0])

    def lite_advance(self, H, t, dt, t_target, miniters=2, maxiters=1000, tol=1e-3):
        H = f64(H).copy(); tt = np.array([t], dtype=np.float64); dd = np.array([dt], dtype=np.float64)
        ns = np.zeros(1, dtype=np.int64); it = np.zeros(1, dtype=np.int64)
        st = lib().cg_lite_advance(C.byref(self.col), dp(H), dp(tt), dp(dd), float(t_target), miniters, maxiters, float(tol),
                                   ns.ctypes.data_as(_I64), it.ctypes.data_as(_I64), C.byref(self.work.c))
        return H, tt[0], dd[0], int(ns[0]), int(it[0]), st

    def init_soil_heat(self, T):
        T = f64(T); H = np.zeros(self.n)
        lib().cg_init_soil_heat(C.byref(self.col), dp(T), dp(H), C.byref(self.work.c))
        return H

    def steadystate_init(self, T0, Qgeo, maxiters=100, thresh=1e-6):
        H = np.zeros(self.n)
        lib().cg_steadystate_init(C.byref(self.col), float(T0), float(Qgeo), int(maxiters), float(thresh), dp(H), C.byref(self.work.c))
        return H


def batch_columns(tile, cols):
    """Array of cg_column structs for the OpenMP batch drivers (keeps the Python owners alive)."""
    ocs = [OracleColumn(tile, int(c)) for c in cols]
    arr = (Column * len(ocs))()
    for i, oc in enumerate(ocs):
        arr[i] = oc.col
    return arr, ocs


def cgeuler_advance_batch(tile, cols, H, t, dt, t_target, dtmin=1.0, dtmax=3600.0, nthreads=0, prepared=None):
    """H is [len(cols), n] (column-contiguous).  Returns (status, nsteps[M]); H, t, dt are updated in place."""
    arr, ocs = prepared or batch_columns(tile, cols)
    ns = np.zeros(len(ocs), dtype=np.int64)
    st = lib().cg_cgeuler_advance_batch(arr, len(ocs), dp(H), dp(t), dp(dt), float(t_target), float(dtmin), float(dtmax),
                                        ns.ctypes.data_as(_I64), int(nthreads))
    return st, ns


def lite_advance_batch(tile, cols, H, t, dt, t_target, miniters=2, maxiters=1000, tol=1e-3, nthreads=0, prepared=None):
    arr, ocs = prepared or batch_columns(tile, cols)
    ns = np.zeros(len(ocs), dtype=np.int64); it = np.zeros(len(ocs), dtype=np.int64)
    st = lib().cg_lite_advance_batch(arr, len(ocs), dp(H), dp(t), dp(dt), float(t_target), miniters, maxiters, float(tol),
                                     ns.ctypes.data_as(_I64), it.ctypes.data_as(_I64), int(nthreads))
    return st, ns, it


class OracleSaltColumn:
    """One column of a SalineTile as the oracle's cg_salt_column (src/Physics/Salt restatement)."""

    def __init__(self, tile, c=0):
        self.tile = tile
        n = tile.grid.ncells
        self.n = n
        self.edges = f64(tile.grid.edges); self.por = f64(tile.por_of(c))
        d = dict(kind=3, alpha=float(tile.alpha_c[c]), n=float(tile.n_c[c]), theta_res=float(tile.thres_c[c]),
                 Tm=float(tile.Tm_c[c]), beta=1.0, omega=1.0)
        self.keep = []
        top = forcing_struct(tile.top_forcing, self.keep)
        self.col = SaltColumn(n=n, surface_state=int(tile.salt_bc.surfaceState), edges=dp(self.edges), por=dp(self.por),
                              sat=float(tile.sat_c[c]), org=float(tile.org_c[c]), fc=fc_struct(d, solver=1), tp=thermal_struct(tile.tp),
                              tau=tile.salt.prop.tau, ds0=tile.salt.prop.ds0, T_ub=tile.T_ub, benthic_salt=float(tile.benthic[c]),
                              bot_value=-tile.Qgeo, courant=tile.salt.courant_number, dc_max=tile.salt.dc_max, heat_courant=0.5,
                              use_top_forcing=int(tile.top_forcing[0] != "constant"), top=top)
        self.arrays = {}
        self.w = SaltWork()
        for name in SaltWork._names:
            a = np.zeros(n + 1)
            self.arrays[name] = a
            setattr(self.w, name, dp(a))

    def get(self, name):
        m = self.n + 1 if name in ("k", "ds", "jH", "jc") else self.n
        return self.arrays[name][:m].copy()

    def rhs(self, T, c, t=0.0):
        T, c = f64(T), f64(c)
        lib().cg_salt_rhs_t(C.byref(self.col), dp(T), dp(c), float(t), C.byref(self.w))
        return self.get("dT"), self.get("dc")

    def timestep(self):
        return lib().cg_salt_timestep(C.byref(self.col), C.byref(self.w))

    def advance(self, T, c, t, dt, t_target, dtmin=1.0, dtmax=3600.0):
        T, c = f64(T).copy(), f64(c).copy()
        tt = np.array([t], dtype=np.float64); dd = np.array([dt], dtype=np.float64); ns = np.zeros(1, dtype=np.int64)
        st = lib().cg_salt_advance(C.byref(self.col), dp(T), dp(c), dp(tt), dp(dd), float(t_target), float(dtmin), float(dtmax),
                                   ns.ctypes.data_as(_I64), C.byref(self.w))
        return T, c, tt[0], dd[0], int(ns[0]), st

    def advance_trace(self, T, c, t, dt, t_target, dtmin=1.0, dtmax=3600.0, cap=1 << 20):
        T, c = f64(T).copy(), f64(c).copy()
        tt = np.array([t], dtype=np.float64); dd = np.array([dt], dtype=np.float64); ns = np.zeros(1, dtype=np.int64)
        tr = np.zeros(cap)
        st = lib().cg_salt_advance_trace(C.byref(self.col), dp(T), dp(c), dp(tt), dp(dd), float(t_target), float(dtmin), float(dtmax),
                                         ns.ctypes.data_as(_I64), C.byref(self.w), dp(tr), cap)
        return T, c, tt[0], dd[0], int(ns[0]), st, tr[:min(int(ns[0]), cap)].copy()
