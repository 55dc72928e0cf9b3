"""Host-side logic (no GPU): stratigraphy flattening, initial conditions and the SFCC host math against the oracle."""
import ctypes as C

import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg
from cryogrid_jl_b200 import hostmath as hm


def test_layer_mapping_samoylov():
    tile = common.samoylov_freew_tile(ncolumns=3)
    e = tile.grid.edges
    cl = tile.cell_layer
    assert cl[0] == 0 and cl[-1] == 4 and np.all(np.diff(cl) >= 0)
    for l, z in enumerate([0.0, 0.1, 0.4, 3.0, 10.0]):
        first = np.nonzero(cl == l)[0][0]
        assert e[first] == pytest.approx(z)
    assert tile.por.shape == (5, 3)
    assert np.all(np.abs(tile.por[0] - 0.8) <= 0.05)
    assert np.all(tile.bot_value == -0.053)


def test_initialcondition_interp_matches_oracle(orc):
    tile = common.samoylov_freew_tile(ncolumns=4, seed=3)
    H0 = cg.initialcondition(tile)
    prof = cg.SamoylovDefault_tempprofile
    zk = np.array([p[0] for p in prof]); vk = np.array([p[1] for p in prof])
    T = np.zeros(tile.grid.ncells)
    orc.lib().cg_interp_profile(orc.dp(zk), orc.dp(vk), zk.size, orc.dp(tile.grid.cells), T.size, orc.dp(T))
    for c in range(tile.M):
        Href = orc.OracleColumn(tile, c).init_soil_heat(T)
        np.testing.assert_allclose(H0[:, c], Href, rtol=1e-14, atol=1e-6)


def test_steadystate_init_matches_oracle(orc):
    tile = common.lite_ensemble_tile(ncolumns=5, seed=1234)
    H0 = cg.initialcondition(tile)
    assert np.all(np.isfinite(H0))
    for c in range(tile.M):
        Href = orc.OracleColumn(tile, c).steadystate_init(-15.0, 0.053)
        np.testing.assert_allclose(H0[:, c], Href, rtol=1e-10, atol=1e-3)


@pytest.mark.parametrize("kind", [1, 2, 3])
def test_sfcc_closed_forms_match_oracle(orc, kind):
    rng = np.random.default_rng(kind)
    for _ in range(50):
        d = dict(kind=kind, alpha=rng.uniform(0.5, 5), n=rng.uniform(1.3, 2.5), theta_res=rng.uniform(0, 0.05), Tm=0.0,
                 beta=rng.uniform(0.8, 1.5), omega=rng.uniform(0.6, 1.0))
        fc = orc.fc_struct(d)
        por, sat, T = rng.uniform(0.2, 0.8), rng.uniform(0.5, 1.0), rng.uniform(-15, 2)
        c = rng.uniform(0, 900) if kind == 3 else 0.0
        dT = C.c_double(0); dc = C.c_double(0)
        th = orc.lib().cg_sfcc(C.byref(fc), T, sat, por, c, C.byref(dT), C.byref(dc))
        th2, dT2 = hm.sfcc_theta(kind, np.float64(T), sat, por, d["alpha"], d["n"], d["theta_res"], 0.0, d["beta"], d["omega"], c)
        assert th == pytest.approx(float(th2), rel=1e-12)
        assert dT.value == pytest.approx(float(dT2), rel=1e-10, abs=1e-14)
        assert d["theta_res"] <= th <= por + 1e-15
        # derivative against central differences
        h = 1e-6
        thp = orc.lib().cg_sfcc(C.byref(fc), T + h, sat, por, c, None, None)
        thm = orc.lib().cg_sfcc(C.byref(fc), T - h, sat, por, c, None, None)
        assert dT.value == pytest.approx((thp - thm) / (2 * h), rel=2e-4, abs=1e-7)
        if kind == 3:
            hc = 1e-3
            cp = orc.lib().cg_sfcc(C.byref(fc), T, sat, por, c + hc, None, None)
            cm = orc.lib().cg_sfcc(C.byref(fc), T, sat, por, c - hc, None, None)
            assert dc.value == pytest.approx((cp - cm) / (2 * hc), rel=2e-4, abs=1e-9)


def test_presolver_table_host_vs_oracle(orc):
    d = dict(kind=2, alpha=1.0, n=2.0, theta_res=0.0, Tm=0.0, beta=1.0, omega=1.0)
    tp = cg.ThermalProperties()
    Hk, Tk = hm.presolver_table(d, tp, 0.5, 1.0, 0.0)
    fc = orc.fc_struct(d); th = orc.thermal_struct(tp)
    Hk2 = np.zeros(10000); Tk2 = np.zeros(10000)
    nk = orc.lib().cg_presolver_build(C.byref(fc), C.byref(th), 0.5, 1.0, 0.0, -60.0, 1e-4, 0.0, orc.dp(Hk2), orc.dp(Tk2), 10000)
    assert nk == Hk.size and 10 < nk < 5000
    np.testing.assert_allclose(Hk, Hk2[:nk], rtol=1e-12)
    np.testing.assert_allclose(Tk, Tk2[:nk], rtol=1e-9, atol=1e-11)
    assert Tk[0] == -60.0 and np.all(np.diff(Hk) > 0) and np.all(np.diff(Tk) > 0)
    # the interpolant inverts the enthalpy function to within the error budget
    for T in np.linspace(-30, -0.01, 40):
        H, _ = hm._sfcc_enthalpy(d, tp, 0.5, 1.0, 0.0, T)
        assert abs(orc.lib().cg_lut_eval(orc.dp(Hk), orc.dp(Tk), nk, 1, H) - T) < 5e-4
        assert float(hm.lut_eval(Hk, Tk, H)) == pytest.approx(orc.lib().cg_lut_eval(orc.dp(Hk), orc.dp(Tk), nk, 1, H), abs=1e-12)


def test_sfcc_newton_inverts_enthalpy(orc):
    d = dict(kind=1, alpha=2.0, n=1.6, theta_res=0.0, Tm=0.0, beta=1.0, omega=1.0)
    tp = cg.ThermalProperties()
    fc = orc.fc_struct(d, solver=1); th = orc.thermal_struct(tp)
    for T in (-40.0, -5.0, -0.5, -0.01, 0.0, 3.0):
        H, _ = hm._sfcc_enthalpy(d, tp, 0.45, 0.9, 0.1, T)
        Ts = orc.lib().cg_sfcc_newton(C.byref(fc), C.byref(th), 0.45, 0.9, 0.1, H, -1.0)
        assert Ts == pytest.approx(T, abs=1e-10)
        assert hm.sfcc_newton(d, tp, 0.45, 0.9, 0.1, H, -1.0) == pytest.approx(Ts, abs=1e-11)
