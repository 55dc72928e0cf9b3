"""Pins the CPU oracle (oracle/cg_oracle.c) against every golden vector / known-answer test the reference's
own test-suite holds for the heat-column hot path.  Citations: /root/reference/test/...
None of these tests reads /root/reference at run time; the expected values are restated from the test files.
"""
import ctypes as C

import numpy as np
import pytest
from scipy.optimize import brentq
from scipy.special import erf, erfc

import common
import cryogrid_jl_b200 as cg


def test_flux_and_divergence_stencil(orc):
    # test/Numerics/math_tests.jl:6-21
    x = np.arange(0.0, 1.0001, 0.25)
    xc = (x[:-1] + x[1:]) / 2
    dx = x[1:] - x[:-1]
    dxc = xc[1:] - xc[:-1]
    y = np.array([0.0, 1.0, 0.0])
    k = np.ones(x.size)
    jy = np.zeros(x.size)
    orc.lib().cg_flux_vec(orc.dp(jy), orc.dp(y), orc.dp(dxc), orc.dp(k), y.size)
    assert jy[0] == 0.0 and jy[-1] == 0.0
    assert jy[1] == pytest.approx(-1 / 0.25)
    assert jy[2] == pytest.approx(1 / 0.25)
    div = np.zeros(y.size)
    orc.lib().cg_divergence_vec(orc.dp(div), orc.dp(jy), orc.dp(dx), y.size)
    np.testing.assert_allclose(div, [1 / 0.25 ** 2, -2 / 0.25 ** 2, 1 / 0.25 ** 2], atol=1e-3)   # test/testutils.jl:14-17


def test_nonlineardiffusion_analytic(orc):
    # test/Numerics/math_tests.jl:23-45
    x = np.exp(np.arange(0.0, 0.05 + 1e-12, 0.001))
    xc = (x[:-1] + x[1:]) / 2
    y = xc ** 3 / 6
    kx = np.sin(x)
    d2y = np.sin(xc) * xc + np.cos(xc) * 0.5 * xc ** 2
    dx = x[1:] - x[:-1]
    dxc = xc[1:] - xc[:-1]
    jy = np.zeros(x.size)
    jy[0] = -kx[0] * 0.5 * x[0] ** 2
    jy[-1] = -kx[-1] * 0.5 * x[-1] ** 2
    div = np.zeros(y.size)
    orc.lib().cg_nonlineardiffusion(orc.dp(div), orc.dp(jy), orc.dp(y), orc.dp(dxc), orc.dp(kx), orc.dp(dx), y.size)
    assert np.all(np.isfinite(div))
    np.testing.assert_allclose(div, d2y, atol=0.01)


def test_tdma_solve(orc):
    # test/Numerics/math_tests.jl:54-61
    n = 5
    A = np.diag(2 * np.ones(n)) + np.diag(np.ones(n - 1), 1) + np.diag(np.ones(n - 1), -1)
    b = np.ones(n)
    x_true = np.linalg.solve(A, b)
    a, bb, c, d, x = np.ones(n - 1), 2 * np.ones(n), np.ones(n - 1), b.copy(), np.zeros(n)
    orc.lib().cg_tdma_solve(orc.dp(x), orc.dp(a), orc.dp(bb), orc.dp(c), orc.dp(d), n)
    np.testing.assert_allclose(x, x_true, rtol=1e-12)


def test_tdma_random_diagonally_dominant(orc):
    rng = np.random.default_rng(0)
    for n in (2, 3, 17, 278):
        a, c = -rng.uniform(0.1, 1, n - 1), -rng.uniform(0.1, 1, n - 1)
        b = 2.5 + rng.uniform(0, 1, n)
        d = rng.normal(size=n)
        A = np.diag(b) + np.diag(c, 1) + np.diag(a, -1)
        x = np.zeros(n)
        orc.lib().cg_tdma_solve(orc.dp(x), orc.dp(a.copy()), orc.dp(b.copy()), orc.dp(c.copy()), orc.dp(d.copy()), n)
        np.testing.assert_allclose(x, np.linalg.solve(A, d), rtol=1e-10, atol=1e-12)


def test_grid(orc):
    # test/Numerics/grid_tests.jl:8-35: cells are midpoints, Δ are first differences
    e = np.arange(0.0, 10.0 + 1e-9, 0.5)
    cells, de, dc = np.zeros(e.size - 1), np.zeros(e.size - 1), np.zeros(e.size - 2)
    orc.lib().cg_grid(orc.dp(e), e.size, orc.dp(cells), orc.dp(de), orc.dp(dc))
    np.testing.assert_allclose(cells, (e[:-1] + e[1:]) / 2)
    np.testing.assert_allclose(de, np.diff(e))
    np.testing.assert_allclose(dc, np.diff(cells))
    g = cg.Grid(e)
    np.testing.assert_array_equal(g.cells, cells)
    np.testing.assert_array_equal(g.dedges, de)
    np.testing.assert_array_equal(g.dcells, dc)


def test_default_grid_sizes():
    # default_grids.jl + docs/src/manual/overview.md:49-51 ; SURVEY 8: N = 278 / 218 / 178
    assert cg.DefaultGrid_2cm.ncells == 278
    assert cg.DefaultGrid_5cm.ncells == 218
    assert cg.DefaultGrid_10cm.ncells == 178
    assert cg.DefaultGrid_5cm.edges[3] == 0.15   # Julia float ranges give the correctly rounded decimal
    assert cg.DefaultGrid_5cm.edges[-1] == 1000.0


def _simple_tile(top, bot, Tinit, grid=None, prop=None):
    grid = grid or cg.Grid(np.linspace(0.0, 1.0, 101))
    heat = cg.HeatBalance(op="H", prop=prop or common.unit_soil_props())
    strat = cg.Stratigraphy((0.0, cg.Top(top)), [(0.0, cg.Ground(cg.SimpleSoil(por=0.0, sat=1.0, org=0.0), heat=heat))],
                            (1.0, cg.Bottom(bot)))
    return cg.Tile(strat, grid, None, ncolumns=1), Tinit * np.ones(grid.ncells)


def test_boundary_flux_signs(orc):
    # test/Physics/Heat/heat_conduction_tests.jl:36-54 (fluxes are positive downward)
    bc = cg.ConstantBC(cg.Dirichlet, 0.0)
    for Tinit, top_sign, bot_sign in ((-1.0, +1, -1), (+1.0, -1, +1)):
        tile, T = _simple_tile(bc, bc, Tinit)
        oc = orc.OracleColumn(tile)
        oc.rhs(T, 0.0)               # unit soil: H == T
        jH = oc.work.get("jH")
        assert np.sign(jH[0]) == top_sign
        assert np.sign(jH[-1]) == bot_sign


def test_neumann_boundary_value(orc):
    # heat_conduction_tests.jl:70-77: Neumann flux is returned verbatim (== -1 W/m^2)
    tile, T = _simple_tile(cg.ConstantBC(cg.Neumann, -1.0), cg.ConstantBC(cg.Neumann, 0.0), 0.0)
    T = np.linspace(-23, 27, T.size)
    oc = orc.OracleColumn(tile)
    oc.rhs(T, 0.0)
    assert oc.work.get("jH")[0] == -1.0


def test_inner_edge_signs(orc):
    # heat_conduction_tests.jl:55-69 nonlineardiffusion! with sin profile: dH[1], dH[end] > 0 (and < 0 mirrored)
    x = np.linspace(0.0, 1.0, 101)
    xc = (x[:-1] + x[1:]) / 2
    k = np.linspace(0.5, 1.0, x.size)
    for sgn in (1.0, -1.0):
        T = sgn * np.sin(xc * np.pi)
        jH, dH = np.zeros(x.size), np.zeros(xc.size)
        orc.lib().cg_nonlineardiffusion(orc.dp(dH), orc.dp(jH), orc.dp(T), orc.dp(np.diff(xc)), orc.dp(k), orc.dp(np.diff(x)), xc.size)
        assert np.sign(dH[0]) == sgn and np.sign(dH[-1]) == sgn


def test_nfactor(orc):
    # heat_conduction_tests.jl:81-98: TemperatureBC(+-10, NFactor(nf=.5, nt=.75)) -> 7.5 and -5.0
    assert orc.lib().cg_nfactor(10.0, 0.5, 0.75) * 10.0 == 7.5
    assert orc.lib().cg_nfactor(-10.0, 0.5, 0.75) * -10.0 == -5.0
    for Tair, expect in ((10.0, 7.5), (-10.0, -5.0)):
        tile, T = _simple_tile(cg.TemperatureBC(Tair, cg.NFactor(nf=0.5, nt=0.75)), cg.ConstantBC(cg.Neumann, 0.0), 0.0)
        oc = orc.OracleColumn(tile)
        oc.rhs(T, 0.0)
        # jH[1] = -k (T[1] - T_ub) / (dk/2) with k = 1, T = 0, dk = 0.01
        assert oc.work.get("jH")[0] == pytest.approx(expect / 0.005, rel=1e-14)


def test_fourier_solution_explicit(orc):
    # heat_conduction_tests.jl:100-140: exp(-4π²t) sin(2πx), grid 0:0.01:1, k = 1, Dirichlet 0/0, Euler dt = 1e-5, t in [0, 0.5]
    tile = common.fourier_tile()
    xc = tile.grid.cells
    oc = orc.OracleColumn(tile)
    H = np.sin(2 * np.pi * xc)
    t, dt = 0.0, 1e-5
    saves = np.arange(1, 51) * 0.01
    errs = []
    for ts in saves:
        H, t, dt, ns, st = oc.advance(H, t, dt, ts, dtmin=1e-5, dtmax=1e-5)
        assert st == 0
        assert t == pytest.approx(ts, abs=1e-12)
        errs.append(np.abs(H - np.exp(-ts * 4 * np.pi ** 2) * np.sin(2 * np.pi * xc)))
    errs = np.array(errs)
    assert np.max(np.mean(errs, axis=0)) < 1e-5       # norm(mean(abs.(...), dims=1), Inf) < 1e-5 K
    assert np.max(errs[-1]) < 1e-5


def test_forcing_interpolation(orc):
    # test/IO/forcing_tests.jl:5-25
    rng = np.random.default_rng(7)
    ts = 1577836800.0 + 86400.0 * np.arange(61)      # 2020-01-01 .. 2020-03-01 daily
    ys = rng.normal(size=ts.size)
    keep = []
    f = orc.forcing_struct(("table", ts, ys), keep)
    st = C.c_int(0)
    ev = lambda t: orc.lib().cg_forcing_eval(C.byref(f), float(t), C.byref(st))
    assert ev(ts[0]) == pytest.approx(ys[0])
    assert ev(ts[-1]) == pytest.approx(ys[-1])
    assert ev((ts[0] + ts[1]) / 2) == pytest.approx((ys[0] + ys[1]) / 2)
    assert st.value == 0
    assert np.isnan(ev(-1.0)) and st.value == 4       # BoundsError outside the range
    st.value = 0
    assert np.isnan(ev(ts[-1] + 1.0)) and st.value == 4
    prov = cg.Interpolated1D(ts, test=ys)
    assert prov("test", ts[3]) == pytest.approx(ys[3])
    with pytest.raises(IndexError):
        prov("test", -1.0)


def test_lite_implicit_periodic_analytic(orc):
    # test/Solvers/cglite_implicit_tests.jl:14-46: |T - analytic| < 0.01 K in the top 10 m over 5 years of daily steps
    P, A, T0 = common.YEAR, 1.0, 1.0
    heat = cg.HeatBalance(op="implicit")
    soil = cg.Ground(cg.SimpleSoil(por=0.0, org=0.0), heat=heat)
    strat = cg.Stratigraphy((0.0, cg.Top(cg.PeriodicBC(cg.Dirichlet, P, 1.0, 0.0, T0))), [(0.0, soil)],
                            (1000.0, cg.Bottom(cg.ConstantBC(cg.Neumann, 0.0))))
    tile = cg.Tile(strat, cg.DefaultGrid_2cm, None, ncolumns=1)
    alpha = heat.prop.kh_m / heat.prop.ch_m
    Tan = lambda z, t: T0 + A * np.exp(-z * np.sqrt(np.pi / (alpha * P))) * np.sin(2 * np.pi * t / P - z * np.sqrt(np.pi / (alpha * P)))
    z = tile.grid.cells
    oc = orc.OracleColumn(tile)
    H = oc.init_soil_heat(Tan(z, 0.0))
    t, dt = 0.0, 86400.0
    top10 = z <= 10.0
    worst = 0.0
    for day in range(1, 5 * 365 + 1):
        H, t, dt, ns, it, st = oc.lite_advance(H, t, dt, day * 86400.0)
        assert st == 0 and ns == 1
        oc.rhs(H, t)
        T = oc.work.get("T")
        worst = max(worst, np.max(np.abs(T[top10] - Tan(z[top10], t))))
    assert worst < 0.01


def _stefan_lambda(k_s, k_l, c_s, c_l, T_s, T_l, thwi, T_m=0.0, rho=1000.0, Lf=3.34e5):
    # src/Physics/Heat/analytic/stefan_analytic.jl:56-65 (Hu et al. 1996, eq. 13)
    a_l, a_s = k_l / c_l, k_s / c_s
    St_l = c_l * (T_l - T_m) / (thwi * Lf * rho)
    St_s = c_s * (T_m - T_s) / (thwi * Lf * rho)
    f = lambda lam: (St_l / (np.exp(lam ** 2) * erf(lam))
                     - St_s * np.sqrt(a_s) / (np.sqrt(a_l) * np.exp(a_l * lam ** 2 / a_s) * erfc(lam * np.sqrt(a_l / a_s)))
                     - lam * np.sqrt(np.pi))
    return brentq(f, 1e-6, 5.0), a_l


def test_lite_implicit_stefan(orc):
    # test/Solvers/cglite_implicit_tests.jl:49-88: thaw depth vs Neumann's two-phase Stefan solution, < 0.01 m
    heat = cg.HeatBalance(op="implicit")
    soil = cg.Ground(cg.SimpleSoil(por=0.3, sat=1.0, org=0.0), heat=heat)
    strat = cg.Stratigraphy((0.0, cg.Top(cg.ConstantTemperature(1.0))), [(0.0, soil)], (1000.0, cg.Bottom(cg.ConstantBC(cg.Neumann, 0.0))))
    tile = cg.Tile(strat, cg.DefaultGrid_2cm, None, ncolumns=1)
    oc = orc.OracleColumn(tile)
    H = oc.init_soil_heat(-1.0 * np.ones(tile.grid.ncells))
    tp = heat.prop
    thm = 0.7
    k_s = (np.sqrt(tp.kh_i) * 0.3 + np.sqrt(tp.kh_m) * thm) ** 2
    k_l = (np.sqrt(tp.kh_w) * 0.3 + np.sqrt(tp.kh_m) * thm) ** 2
    c_s = tp.ch_i * 0.3 + tp.ch_m * thm
    c_l = tp.ch_w * 0.3 + tp.ch_m * thm
    lam, a_l = _stefan_lambda(k_s, k_l, c_s, c_l, -1.0, 1.0, 0.3)
    thick = tile.grid.dedges
    t, dt = 0.0, 86400.0
    worst = 0.0
    for day in range(1, 5 * 365 + 1):
        H, t, dt, ns, it, st = oc.lite_advance(H, t, dt, day * 86400.0)
        assert st == 0
        oc.rhs(H, t)
        td = np.sum(oc.work.get("thw") / 0.3 * thick)
        worst = max(worst, abs(td - 2 * lam * np.sqrt(a_l * t)))
    assert worst < 0.01


def test_energy_conservation_zero_net_flux(orc):
    # examples/heat_sfcc_constantbc.jl:36-57 (as written: FreeWater, 10 W/m^2 in and out): integrate(H) is conserved
    tile = cg.SoilHeatTile("H", cg.ConstantBC(cg.Neumann, 10.0), cg.ConstantBC(cg.Neumann, 10.0), cg.SoilProfile((0.0, cg.SimpleSoil())),
                           None, cg.initializer("T", cg.TemperatureProfile((0.0, -20.0), (1000.0, 10.0))), grid=cg.DefaultGrid_10cm)
    H0 = cg.initialcondition(tile)[:, 0]
    oc = orc.OracleColumn(tile)
    np.testing.assert_allclose(oc.init_soil_heat(np.interp(tile.grid.cells, [0.0, 1000.0], [-20.0, 10.0])), H0, rtol=1e-13)
    H, t, dt, ns, st = oc.advance(H0, 0.0, 60.0, 30 * 86400.0, dtmin=1.0, dtmax=900.0)
    assert st == 0 and t == 30 * 86400.0
    dk = tile.grid.dedges
    Htot0, Htot1 = np.sum(H0 * dk), np.sum(H * dk)
    assert abs(Htot1 - Htot0) <= 1e-9 * abs(Htot0)


def test_freewater_inverse_enthalpy_branches(orc):
    # Heat/heat_conduction.jl:18-57: frozen / phase change / thawed, and dHdT = 1e8 exactly at T == 0
    tile = cg.SoilHeatTile("H", cg.ConstantBC(cg.Neumann, 0.0), cg.ConstantBC(cg.Neumann, 0.0), cg.SoilProfile((0.0, cg.SimpleSoil())),
                           None, cg.initializer("T", 0.0), grid=cg.Grid(np.linspace(0, 1, 5)))
    oc = orc.OracleColumn(tile)
    L, thwi = 3.34e8, 0.5
    H = np.array([-1e6, 0.25 * L * thwi, L * thwi, L * thwi + 1e6])
    oc.rhs(H, 0.0)
    T, thw, dHdT, Cc = oc.work.get("T"), oc.work.get("thw"), oc.work.get("dHdT"), oc.work.get("C")
    assert T[0] < 0 and T[1] == 0.0 and T[2] == 0.0 and T[3] > 0
    np.testing.assert_allclose(thw, [0.0, 0.125, 0.5, 0.5], rtol=1e-15)
    assert dHdT[1] == 1e8 and dHdT[2] == 1e8 and dHdT[0] == Cc[0] and dHdT[3] == Cc[3]
    assert T[3] == pytest.approx(1e6 / Cc[3])
