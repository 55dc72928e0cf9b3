"""GPU parity tests of the implicit (LiteImplicitEuler / CryoGridLite) path against the CPU oracle, plus the
reference's own analytic acceptance tests (test/Solvers/cglite_implicit_tests.jl) run through the CUDA path."""
import os

import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["warp", "thread"])
def variant(request):
    """Both batched-tridiagonal variants: partition/PCR per warp (default) and the reference-order Thomas per thread."""
    old = os.environ.get("CGB_LITE_VARIANT")
    os.environ["CGB_LITE_VARIANT"] = request.param
    yield request.param
    if old is None:
        del os.environ["CGB_LITE_VARIANT"]
    else:
        os.environ["CGB_LITE_VARIANT"] = old


def test_lite_diag_coefficients(orc):
    tile = common.lite_ensemble_tile(ncolumns=11, seed=1234)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, stepper="lite")
    eng.set_state(H0, 0.0)
    t = 40 * 86400.0
    got = {v: eng.diag(v, t) for v in ("T", "thw", "dHdT", "kc", "k", "DT_an", "DT_as", "DT_ap", "DT_bp")}
    for c in range(tile.M):
        oc = orc.OracleColumn(tile, c)
        oc.rhs(H0[:, c], t)
        for v, name in (("T", "T"), ("thw", "thw"), ("dHdT", "dHdT"), ("kc", "kc"), ("k", "k"), ("DT_an", "an"), ("DT_as", "as"),
                        ("DT_ap", "ap"), ("DT_bp", "bp")):
            ref = oc.work.get(name)
            np.testing.assert_allclose(got[v][:, c], ref, rtol=1e-12, atol=1e-12 * max(1.0, np.max(np.abs(ref))), err_msg=v)
    assert np.all(eng.rhs(t) == 0.0)     # heat_implicit.jl:106
    eng.close()


def test_lite_one_year_ensemble(orc, variant):
    # cfg3 shape at test size: daily steps for one year; <= 1e-6 K and <= 1e-8 θw against the oracle, equal iteration counts
    tile = common.lite_ensemble_tile(ncolumns=9, seed=1234)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, stepper="lite")
    eng.set_state(H0, 0.0)
    days = [1, 2, 30, 120, 200, 365]
    ref = {c: (H0[:, c].copy(), 0.0, 86400.0, 0, 0) for c in range(tile.M)}
    for d in days:
        tt = d * 86400.0
        eng.advance_to(tt)
        H, t, dt = eng.get_state(with_time=True)
        st, steps = eng.status()
        assert not st.any()
        Tg, thwg = eng.diag("T", tt), eng.diag("thw", tt)
        for c in range(tile.M):
            oc = orc.OracleColumn(tile, c)
            Hr, tr, dtr, ns, it = ref[c]
            Hr, tr, dtr, n2, i2, s2 = oc.lite_advance(Hr, tr, dtr, tt)
            ref[c] = (Hr, tr, dtr, ns + n2, it + i2)
            assert s2 == 0 and t[c] == tr == tt and steps[c] == ns + n2 == d
            oc.rhs(Hr, tt)
            assert np.max(np.abs(Tg[:, c] - oc.work.get("T"))) <= 1e-6
            assert np.max(np.abs(thwg[:, c] - oc.work.get("thw"))) <= 1e-8
    total_iters = sum(v[4] for v in ref.values())
    got = eng.stats()["lite_iterations"]
    if variant == "thread":
        assert got == total_iters          # same operation order as the reference: identical Picard trajectories
    else:
        assert abs(got - total_iters) <= max(1, 1e-3 * total_iters)   # rounding may flip a borderline convergence test
    eng.close()


def test_lite_periodic_analytic_on_gpu(variant):
    # test/Solvers/cglite_implicit_tests.jl:14-46 through the CUDA path
    P, A, T0 = common.YEAR, 1.0, 1.0
    heat = cg.HeatBalance(op="implicit")
    soil = cg.Ground(cg.SimpleSoil(por=0.0, org=0.0), heat=heat)
    strat = cg.Stratigraphy((0.0, cg.Top(cg.PeriodicBC(cg.Dirichlet, P, 1.0, 0.0, T0))), [(0.0, soil)],
                            (1000.0, cg.Bottom(cg.ConstantBC(cg.Neumann, 0.0))))
    tile = cg.Tile(strat, cg.DefaultGrid_2cm, None, ncolumns=2)
    alpha = heat.prop.kh_m / heat.prop.ch_m
    Tan = lambda z, t: T0 + A * np.exp(-z * np.sqrt(np.pi / (alpha * P))) * np.sin(2 * np.pi * t / P - z * np.sqrt(np.pi / (alpha * P)))
    z = tile.grid.cells
    H0 = tile.enthalpy_from_T(np.repeat(Tan(z, 0.0)[:, None], 2, axis=1))
    prob = cg.CryoGridProblem(tile, H0, (0.0, 5 * P), saveat=86400.0, savevars=("T",))
    sol = cg.solve(prob, cg.LiteImplicitEuler(), dt=86400.0)
    out = cg.CryoGridOutput(sol)
    top10 = z <= 10.0
    err = out.T[:, top10, 0] - Tan(z[top10][None, :], sol.t[:, None])
    assert sol.retcode == "Success" and out.T.shape[0] == 5 * 365 + 1
    assert np.max(np.abs(err)) < 0.01


def test_lite_stefan_on_gpu(variant):
    # test/Solvers/cglite_implicit_tests.jl:49-88 through the CUDA path (thaw depth vs Neumann solution < 0.01 m)
    from test_oracle_golden import _stefan_lambda
    heat = cg.HeatBalance(op="implicit")
    soil = cg.Ground(cg.SimpleSoil(por=0.3, sat=1.0, org=0.0), heat=heat)
    strat = cg.Stratigraphy((0.0, cg.Top(cg.ConstantTemperature(1.0))), [(0.0, soil)], (1000.0, cg.Bottom(cg.ConstantBC(cg.Neumann, 0.0))))
    tile = cg.Tile(strat, cg.DefaultGrid_2cm, None, cg.initializer("T", -1.0), ncolumns=1)
    H0 = cg.initialcondition(tile)
    prob = cg.CryoGridProblem(tile, H0, (0.0, 5 * common.YEAR), saveat=86400.0, savevars=("T", "θw"))
    sol = cg.solve(prob, cg.LiteImplicitEuler(), dt=86400.0)
    out = cg.CryoGridOutput(sol)
    tp = heat.prop
    k_s = (np.sqrt(tp.kh_i) * 0.3 + np.sqrt(tp.kh_m) * 0.7) ** 2
    k_l = (np.sqrt(tp.kh_w) * 0.3 + np.sqrt(tp.kh_m) * 0.7) ** 2
    lam, a_l = _stefan_lambda(k_s, k_l, tp.ch_i * 0.3 + tp.ch_m * 0.7, tp.ch_w * 0.3 + tp.ch_m * 0.7, -1.0, 1.0, 0.3)
    td = np.sum(out.vars["θw"][:, :, 0] / 0.3 * tile.grid.dedges[None, :], axis=1)
    assert np.max(np.abs(td - 2 * lam * np.sqrt(a_l * sol.t))) < 0.01


def test_lite_maxiters_status(variant):
    tile = common.lite_ensemble_tile(ncolumns=4, seed=7)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, stepper="lite", maxiters=1, tolerance=1e-12, miniters=1)
    eng.set_state(H0, 0.0)
    eng.advance_to(30 * 86400.0)
    st, steps = eng.status()
    assert np.all(st & 1) and np.all(steps == 30)     # ReturnCode.MaxIters, cglite_step.jl:74-77
    eng.close()


@pytest.mark.parametrize("grid_name", ["DefaultGrid_1m", "DefaultGrid_20cm", "DefaultGrid_5cm"])
def test_lite_warp_vs_thread_other_grids(orc, grid_name):
    # different cells-per-lane instantiations of the warp kernel against the oracle (Dirichlet bottom: heat_implicit.jl:83 quirk)
    grid = getattr(cg, grid_name)
    heat = cg.HeatBalance(op="implicit")
    layers = [(0.0, cg.Ground(cg.SimpleSoil(por=np.array([0.6, 0.4, 0.5]), org=0.3), heat=heat)),
              (1.0, cg.Ground(cg.SimpleSoil(por=0.35), heat=heat)), (20.0, cg.Ground(cg.SimpleSoil(por=0.1), heat=heat))]
    strat = cg.Stratigraphy((0.0, cg.Top(cg.PeriodicBC(cg.Dirichlet, common.YEAR, 12.0, 0.5, -3.0))), layers,
                            (1000.0, cg.Bottom(cg.ConstantBC(cg.Dirichlet, 9.0))))
    tile = cg.Tile(strat, grid, None, cg.initializer("T", cg.TemperatureProfile((0.0, -4.0), (10.0, -2.0), (1000.0, 9.0))), ncolumns=3)
    H0 = cg.initialcondition(tile)
    os.environ["CGB_LITE_VARIANT"] = "warp"
    try:
        eng = cg.Engine(tile, stepper="lite")
        eng.set_state(H0, 0.0)
        eng.advance_to(60 * 86400.0)
        T = eng.diag("T", 60 * 86400.0)
        st, steps = eng.status()
        assert not st.any() and np.all(steps == 60)
        for c in range(tile.M):
            oc = orc.OracleColumn(tile, c)
            Hr, *_ = oc.lite_advance(H0[:, c], 0.0, 86400.0, 60 * 86400.0)
            oc.rhs(Hr, 60 * 86400.0)
            assert np.max(np.abs(T[:, c] - oc.work.get("T"))) <= 1e-6
        eng.close()
    finally:
        del os.environ["CGB_LITE_VARIANT"]


def test_device_steadystate_init(orc):
    # ThermalSteadyStateInit on the device vs the oracle (Heat/heat_init.jl:58-85)
    tile = common.lite_ensemble_tile(ncolumns=37, seed=1234)
    eng = cg.Engine(tile, stepper="lite")
    eng.initialcondition(0.0)
    H = eng.get_state()
    assert np.all(np.isfinite(H))
    for c in range(0, tile.M, 6):
        Href = orc.OracleColumn(tile, c).steadystate_init(-15.0, 0.053)
        np.testing.assert_allclose(H[:, c], Href, rtol=1e-9, atol=1e-2)
    eng.advance_to(86400.0)
    assert not eng.status()[0].any()
    eng.close()


def test_lite_saved_diagnostics_follow_reference_cache(orc, variant):
    # saved T/θw = cache of the last tile evaluation (start of the last Picard iteration), cglite_step.jl:45-62
    tile = common.lite_ensemble_tile(ncolumns=5, seed=11)
    H0 = cg.initialcondition(tile)
    prob = cg.CryoGridProblem(tile, H0, (0.0, 20 * 86400.0), saveat=86400.0, savevars=("T", "θw", "H"))
    sol = cg.solve(prob, cg.LiteImplicitEuler())
    out = cg.CryoGridOutput(sol)
    for c in range(tile.M):
        oc = orc.OracleColumn(tile, c)
        Hr, tr, dtr = H0[:, c].copy(), 0.0, 86400.0
        for s in range(1, 21):
            Hr, tr, dtr, *_ = oc.lite_advance(Hr, tr, dtr, s * 86400.0)
            assert np.max(np.abs(out.T[s][:, c] - oc.work.get("T"))) <= 1e-6
            assert np.max(np.abs(out.vars["θw"][s][:, c] - oc.work.get("thw"))) <= 1e-8
            np.testing.assert_allclose(out.H[s][:, c], Hr, rtol=1e-10, atol=1e-2)


def test_lite_multi_target_saveat_vs_oracle(orc):
    # device-side saveat of the warp kernel (several daily saves per launch) against the oracle's work arrays, and
    # against the one-launch-per-save path
    tile = common.lite_ensemble_tile(ncolumns=7, seed=12)
    H0 = cg.initialcondition(tile)
    days = 86400.0 * np.arange(1, 20)
    a = cg.Engine(tile, stepper="lite"); a.set_state(H0, 0.0)
    multi = a.solve_saveat(days, ("T", "θw"))                        # chunked path
    b = cg.Engine(tile, stepper="lite"); b.set_state(H0, 0.0)
    single = b.solve_saveat(days, ("T", "θw", "H"))                  # H saved -> one launch per save
    np.testing.assert_array_equal(multi[:, 0], single[:, 0])
    np.testing.assert_array_equal(multi[:, 1], single[:, 1])
    np.testing.assert_array_equal(a.get_state(), b.get_state())
    assert a.stats()["kernel_launches"] < b.stats()["kernel_launches"]
    for c in (0, 6):
        oc = orc.OracleColumn(tile, c)
        Hr, tr, dtr = H0[:, c].copy(), 0.0, 86400.0
        for s, t in enumerate(days):
            Hr, tr, dtr, *_ = oc.lite_advance(Hr, tr, dtr, t)
            assert np.max(np.abs(multi[s, 0][:, c] - oc.work.get("T"))) <= 1e-6
            assert np.max(np.abs(multi[s, 1][:, c] - oc.work.get("thw"))) <= 1e-8
    a.close(); b.close()
