"""SFCCPreSolver tables built on the device (cgb_build_sfcc_tables, csrc/cgb_presolver.cuh) against the oracle's restatement of
FreezeCurves.initialize! (call site Soils/soil_heat.jl:26-36; default solver Soils/ground.jl:27).  One thread per distinct soil
class; classes are deduplicated on the host.  The device uses the same arithmetic as the host builders (pow, true divisions,
uncontracted), so the knots agree to rounding -- bit-equality is not attainable because CUDA's pow and libm's pow differ in
the last ulp."""
import ctypes as C

import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg

pytestmark = pytest.mark.gpu


def _oracle_table(orc, tile, l, c, Tmin=-60.0, errtol=1e-4, Tmax=0.0, cap=1 << 16):
    fc = orc.fc_struct(tile.fc_dict(l, c), solver=0)
    tp = orc.thermal_struct(tile.tp)
    Hk = np.zeros(cap); Tk = np.zeros(cap)
    por, sat, org = float(tile.por[l, c]), float(tile.sat[l, c]), float(tile.org[l, c])
    nk = orc.lib().cg_presolver_build(C.byref(fc), C.byref(tp), por, (por * sat) / por, org, Tmin, errtol, Tmax, orc.dp(Hk), orc.dp(Tk), cap)
    assert nk > 0
    return Hk[:nk].copy(), Tk[:nk].copy()


def test_device_tables_match_the_oracle_builder_on_random_classes(orc):
    M = 1000
    tile = common.sfcc_tile(ncolumns=M, kind="dallamico", solver="lut", per_column=True, grid=cg.DefaultGrid_1m, seed=11)
    rng = np.random.default_rng(12)
    tile.sat = np.where(rng.random((1, M)) < 0.3, rng.uniform(0.5, 1.0, (1, M)), 1.0)       # some unsaturated classes
    eng = cg.Engine(tile, device_tables=True)
    assert eng.n_device_tables == M
    worst = dict(dH=0.0, dT=0.0, mismatched_counts=0)
    for c in range(0, M, 7):
        Hd, Td = eng.get_sfcc_table(c)
        Ho, To = _oracle_table(orc, tile, 0, c)
        if len(Hd) != len(Ho):
            worst["mismatched_counts"] += 1          # a halving decision within rounding of errtol: possible, must stay rare
            continue
        worst["dH"] = max(worst["dH"], float(np.max(np.abs(Hd - Ho)) / np.max(np.abs(Ho))))     # knots cross H = 0: scale by the table range
        worst["dT"] = max(worst["dT"], float(np.max(np.abs(Td - To))))
    print("device presolver vs oracle builder:", worst)
    assert worst["mismatched_counts"] <= 1
    assert worst["dH"] <= 1e-14 and worst["dT"] <= 1e-10
    eng.close()


def test_device_tables_are_deduplicated_and_drive_the_lut_kernels(orc):
    # 600 columns but only 3 distinct soil classes in 2 layers -> at most 6 tables; the engine then runs the reference's default
    # LUT semantics (cfg4 with the presolver instead of the Newton accuracy mode)
    M = 600
    rng = np.random.default_rng(5)
    cls = rng.integers(0, 3, M)
    por = np.array([0.35, 0.5, 0.65])[cls]; alpha = np.array([0.8, 2.0, 4.0])[cls]; n = np.array([1.4, 1.6, 2.2])[cls]
    top = cg.SimpleSoil(por=por, sat=1.0, org=0.3, freezecurve=cg.DallAmico(swrc=cg.VanGenuchten(alpha=alpha, n=n)))
    bot = cg.SimpleSoil(por=por * 0.8, sat=1.0, org=0.0, freezecurve=cg.DallAmico(swrc=cg.VanGenuchten(alpha=alpha, n=n)))
    tile = cg.SoilHeatTile("H", cg.PeriodicBC(cg.Dirichlet, common.YEAR, 18.0, 0.0, -10.0), cg.GeothermalHeatFlux(0.053),
                           cg.SoilProfile((0.0, top), (2.0, bot)), None, cg.initializer("T", cg.SamoylovDefault_tempprofile),
                           grid=cg.DefaultGrid_10cm, fcsolver=cg.SFCCPreSolver(), ncolumns=M)
    eng = cg.Engine(tile, dtmax=600.0, device_tables=True)
    assert eng.n_device_tables == 6
    H0 = cg.initialcondition(tile)
    eng.set_state(H0, 0.0)
    eng.advance_to(86400.0)
    H, t, dt = eng.get_state(with_time=True)
    st, steps = eng.status()
    assert not st.any()
    Tg = eng.diag("T", 86400.0)
    # the oracle with ITS OWN (host-built) tables: the two table sets agree to rounding, so the trajectories agree to the gate
    for c in (0, 1, 2, 599):
        oc = orc.OracleColumn(tile, c)
        Hr, tr, dtr, ns, s2 = oc.advance(H0[:, c], 0.0, 60.0, 86400.0, dtmax=600.0)
        assert s2 == 0 and steps[c] == ns
        oc.rhs(Hr, 86400.0)
        assert np.max(np.abs(Tg[:, c] - oc.work.get("T"))) <= 1e-6
    eng.close()


def test_device_tables_default_policy():
    # caller-built knots always win; few classes keep the host builder (bit-identical tables on both sides of the parity tests)
    tile = common.sfcc_tile(ncolumns=4, kind="dallamico", solver="lut")
    eng = cg.Engine(tile)
    assert not eng._use_device_tables()
    eng.close()
    tile = common.sfcc_tile(ncolumns=64, kind="dallamico", solver="lut", per_column=True, grid=cg.DefaultGrid_1m)
    eng = cg.Engine(tile)
    assert eng._use_device_tables() and eng.n_device_tables == 64
    eng.close()
