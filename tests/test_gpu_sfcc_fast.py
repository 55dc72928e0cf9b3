"""The SFCC single-class fast path (csrc/cgb_euler_sfcc_fast.cuh: many columns of ONE soil class through the presolver LUT,
MaxDelta limiter -- the north_star target configuration) against the oracle and against the general kernel k_heat_step<.,1>.
CGB_SFCC_FAST=0 (read at every finalize) routes the same problem through the general kernel."""
import ctypes as C
import os

import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg
from cryogrid_jl_b200 import _lib as L
from test_gpu_explicit import _check_advance

pytestmark = pytest.mark.gpu
DAY = 86400.0


class general_kernel:
    def __enter__(self):
        self.old = os.environ.get("CGB_SFCC_FAST")
        os.environ["CGB_SFCC_FAST"] = "0"

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("CGB_SFCC_FAST", None)
        else:
            os.environ["CGB_SFCC_FAST"] = self.old


def _tile(M, kind="dallamico", alpha=2.0, n=1.6, sat=1.0, por=0.5, org=0.0, grid=None, extrap="line"):
    swrc = cg.VanGenuchten(alpha=alpha, n=n)
    fc = cg.PainterKarra(swrc=swrc) if kind == "painterkarra" else cg.DallAmico(swrc=swrc)
    soil = cg.SimpleSoil(por=por, sat=sat, org=org, freezecurve=fc)
    tile = cg.SoilHeatTile("H", cg.PeriodicBC(cg.Dirichlet, common.YEAR, 18.0, 0.0, -10.0), cg.GeothermalHeatFlux(0.053),
                           cg.SoilProfile((0.0, soil)), None, cg.initializer("T", cg.SamoylovDefault_tempprofile),
                           grid=grid or cg.DefaultGrid_5cm, fcsolver=cg.SFCCPreSolver(extrap=extrap), ncolumns=M)
    tile.top_offset = np.random.default_rng(7).normal(0.0, 2.0, M)
    return tile


@pytest.mark.parametrize("kind,alpha,n,sat", [("dallamico", 2.0, 1.6, 1.0), ("painterkarra", 1.0, 2.0, 1.0), ("dallamico", 0.5, 1.3, 1.0),
                                              ("dallamico", 4.06, 3.0, 1.0), ("painterkarra", 2.0, 1.8, 0.7)])
def test_curve_table_reproduces_the_closed_form(orc, kind, alpha, n, sat):
    tile = _tile(2, kind, alpha, n, sat)
    eng = cg.Engine(tile)
    eng.set_state(cg.initialcondition(tile), 0.0)
    rng = np.random.default_rng(0)
    T = np.concatenate([-np.logspace(-12, 2.4, 20000), -rng.uniform(0.0, 60.0, 20000), -np.logspace(-9, 0, 4000) * (1 + 1e-9),
                        rng.uniform(0.0, 30.0, 2000), [0.0, -0.0, 1e-300, -1e-300]])
    got = eng.selftest_sfcc_curve(T)
    eng.close()
    fc = orc.fc_struct(tile.fc_dict(0, 0), solver=1)
    sat_e = float(tile.por[0, 0] * tile.sat[0, 0] / tile.por[0, 0])
    ref = np.array([orc.lib().cg_sfcc(C.byref(fc), float(t), sat_e, float(tile.por[0, 0]), 0.0, None, None) for t in T])
    rel = np.abs(got - ref) / ref
    print(f"curve table vs closed form ({kind}, alpha={alpha}, n={n}, sat={sat}): max rel {rel.max():.2e} at T = {T[rel.argmax()]:.3e}")
    assert rel.max() <= 2e-13


def test_not_eligible_configurations_report_unsupported():
    tile = common.sfcc_tile(ncolumns=4, kind="dallamico", solver="lut", per_column=False)
    tile.por = tile.por + np.linspace(0.0, 0.05, 4)[None, :]          # per-column porosity: more than one class
    eng = cg.Engine(tile)
    eng.set_state(cg.initialcondition(tile), 0.0)
    with pytest.raises(L.CgbError) as e:
        eng.selftest_sfcc_curve(np.array([-1.0]))
    assert e.value.code == L.ERR_UNSUPPORTED
    eng.close()


@pytest.mark.parametrize("kind,alpha,n,sat", [("dallamico", 2.0, 1.6, 1.0), ("painterkarra", 1.0, 2.0, 1.0), ("painterkarra", 2.0, 1.8, 0.7)])
def test_fast_path_matches_oracle_strict(orc, kind, alpha, n, sat):
    # stable scheme (dtmax = 600 s on the 5 cm grid): the north_star gate at every save point, equal step counts
    tile = _tile(6, kind, alpha, n, sat)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=600.0)
    w = _check_advance(orc, tile, eng, H0, 0.0, [3600.0, DAY, 5 * DAY, 20 * DAY], dtmax=600.0, cols=(0, 3, 5))
    print("SFCC fast path vs oracle, 20 days:", w)
    assert eng.selftest_sfcc_curve(np.array([-1.0])).size == 1        # the fast path was the one that ran
    eng.close()


def test_fast_path_matches_general_kernel():
    tile = _tile(64)
    H0 = cg.initialcondition(tile)
    res = {}
    for name in ("fast", "general"):
        ctx = general_kernel() if name == "general" else None
        if ctx:
            ctx.__enter__()
        try:
            eng = cg.Engine(tile, dtmax=600.0)
            eng.set_state(H0, 0.0)
            out = eng.solve_saveat(DAY * np.arange(1, 6), ("T", "θw"))
            H, t, dt = eng.get_state(with_time=True)
            st, steps = eng.status()
            res[name] = (out, H, dt, steps)
            eng.close()
        finally:
            if ctx:
                ctx.__exit__()
    assert np.array_equal(res["fast"][3], res["general"][3])
    assert np.max(np.abs(res["fast"][0][:, 0] - res["general"][0][:, 0])) <= 1e-9
    assert np.max(np.abs(res["fast"][0][:, 1] - res["general"][0][:, 1])) <= 1e-11
    np.testing.assert_allclose(res["fast"][2], res["general"][2], rtol=1e-9)


def test_fast_path_flat_extrapolation_and_small_grids(orc):
    # dtmax below the diffusion limit of the finest cells of each grid (2 cm: ~150 s), so that the strict gate applies
    for name, grid, ex, dtmax in (("10cm", cg.DefaultGrid_10cm, "flat", 600.0), ("1m", cg.DefaultGrid_1m, "line", 3600.0),
                                  ("6 cells", cg.Grid(np.linspace(0.0, 3.0, 7)), "flat", 3600.0), ("2cm", cg.DefaultGrid_2cm, "line", 100.0)):
        tile = _tile(3, "dallamico", 2.0, 1.6, grid=grid, extrap=ex)
        H0 = cg.initialcondition(tile)
        H0[:, 1] += 3.0e8 * (np.arange(H0.shape[0]) % 3 == 0)          # thawed cells far above the last knot
        H0[:, 2] -= 2.0e8 * (np.arange(H0.shape[0]) % 4 == 1)          # and cells far below the first one
        eng = cg.Engine(tile, dtmax=dtmax)
        try:
            w = _check_advance(orc, tile, eng, H0, 0.0, [3600.0, DAY], dtmax=dtmax)
        except AssertionError as e:
            raise AssertionError(f"grid {name}, extrapolation {ex}: {e}") from e
        print(f"grid {name} ({ex}):", w)
        eng.close()


def test_fast_path_large_ragged_batch_saved_outputs_follow_reference_cache(orc):
    # 3001 columns (not a multiple of the 12 warps per CTA); saved T, θw = cache of the LAST RHS evaluation (problem.jl:67-68)
    M = 3001
    tile = _tile(M)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=600.0)
    eng.set_state(H0, 0.0)
    out = eng.solve_saveat(np.array([0.5 * DAY, DAY]), ("T", "θw"))
    st, steps = eng.status()
    assert not st.any()
    for c in (0, 1500, 3000):
        oc = orc.OracleColumn(tile, c)
        Hr, tr, dtr = H0[:, c], 0.0, 60.0
        for k, tt in enumerate((0.5 * DAY, DAY)):
            Hr, tr, dtr, ns, s2 = oc.advance(Hr, tr, dtr, tt, dtmax=600.0)
            assert np.max(np.abs(out[k, 0, :, c] - oc.work.get("T"))) <= 1e-9
            assert np.max(np.abs(out[k, 1, :, c] - oc.work.get("thw"))) <= 1e-11
    eng.close()


# ---- per-column parameters: the Newton kernel with the column's own curve table (csrc/cgb_euler_newton_tab.cuh) -------------
class general_newton:
    def __enter__(self):
        self.old = os.environ.get("CGB_NEWTON_TAB")
        os.environ["CGB_NEWTON_TAB"] = "0"

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("CGB_NEWTON_TAB", None)
        else:
            os.environ["CGB_NEWTON_TAB"] = self.old


@pytest.mark.parametrize("kind", ["dallamico", "painterkarra"])
def test_newton_tab_matches_oracle_strict(orc, kind):
    # cfg4 shape (per-column alpha, n, porosity, organic fraction; one layer), stable dtmax: north_star gate at every save point
    tile = common.sfcc_tile(ncolumns=20, kind=kind, solver="newton", per_column=True, grid=cg.DefaultGrid_5cm, seed=21,
                            top=cg.PeriodicBC(cg.Dirichlet, common.YEAR, 18.0, 0.0, -10.0), bottom=cg.GeothermalHeatFlux(0.053),
                            temp=cg.SamoylovDefault_tempprofile)
    tile.top_offset = np.random.default_rng(3).normal(0.0, 2.0, 20)
    tile.sat[:, 10:] = np.random.default_rng(4).uniform(0.6, 1.0, 10)         # unsaturated columns: psi0 < 0, T* < Tm
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=600.0)
    w = _check_advance(orc, tile, eng, H0, 0.0, [3600.0, DAY, 5 * DAY, 20 * DAY], dtmax=600.0, cols=(0, 7, 12, 19))
    print(f"Newton with per-column curve tables vs oracle ({kind}), 20 days:", w)
    eng.close()


def test_newton_tab_matches_general_newton_kernel_and_saves():
    tile = common.cfg4_tile(96, "newton")
    H0 = cg.initialcondition(tile)
    res = {}
    for name in ("tab", "tab_nocache", "general"):
        ctx = general_newton() if name == "general" else None
        if ctx:
            ctx.__enter__()
        if name == "tab_nocache":
            os.environ["CGB_NT_CACHE"] = "0"
        try:
            eng = cg.Engine(tile, dtmax=600.0)
            eng.set_state(H0, 0.0)
            out = np.concatenate([eng.solve_saveat(DAY * np.arange(1, 3), ("T", "θw")), eng.solve_saveat(DAY * np.arange(3, 4), ("T", "θw"))])
            H, t, dt = eng.get_state(with_time=True)
            st, steps = eng.status()
            assert not st.any()
            res[name] = (out, dt, steps)
            eng.close()
        finally:
            os.environ.pop("CGB_NT_CACHE", None)
            if ctx:
                ctx.__exit__()
    # the tables read back from HBM in the second and third launch are the fitted ones: bit-identical results
    assert np.array_equal(res["tab"][0], res["tab_nocache"][0]) and np.array_equal(res["tab"][1], res["tab_nocache"][1])
    assert np.array_equal(res["tab"][2], res["general"][2])
    assert np.max(np.abs(res["tab"][0][:, 0] - res["general"][0][:, 0])) <= 1e-9
    assert np.max(np.abs(res["tab"][0][:, 1] - res["general"][0][:, 1])) <= 1e-11
    np.testing.assert_allclose(res["tab"][1], res["general"][1], rtol=1e-9)


def test_newton_tab_small_grids_and_extreme_states(orc):
    for name, grid, dtmax in (("6 cells", cg.Grid(np.linspace(0.0, 3.0, 7)), 3600.0), ("1m", cg.DefaultGrid_1m, 3600.0), ("2cm", cg.DefaultGrid_2cm, 100.0)):
        tile = common.sfcc_tile(ncolumns=5, kind="dallamico", solver="newton", per_column=True, grid=grid, seed=2,
                                top=cg.PeriodicBC(cg.Dirichlet, common.YEAR, 18.0, 0.0, -10.0), bottom=cg.GeothermalHeatFlux(0.053),
                                temp=cg.SamoylovDefault_tempprofile)
        H0 = cg.initialcondition(tile)
        H0[:, 1] += 3.0e8 * (np.arange(H0.shape[0]) % 3 == 0)          # thawed cells
        H0[:, 2] -= 2.5e8 * (np.arange(H0.shape[0]) % 4 == 1)          # cells below -64 degC: outside the curve table (closed form)
        eng = cg.Engine(tile, dtmax=dtmax)
        try:
            w = _check_advance(orc, tile, eng, H0, 0.0, [3600.0, DAY], dtmax=dtmax)
        except AssertionError as e:
            raise AssertionError(f"grid {name}: {e}") from e
        print(f"Newton-tab, grid {name}:", w)
        eng.close()
