"""GPU parity tests of the explicit (CGEuler) path: CUDA engine (through the C ABI) vs the CPU oracle on the same
seeded inputs.  Tolerances (BASELINE.json north_star): relative 1e-10 on a single RHS evaluation (scaled as in
SURVEY 8d), <= 1e-6 K temperature and <= 1e-8 liquid-water fraction after the integration, equal step counts."""
import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg
from cryogrid_jl_b200 import _lib as L

pytestmark = pytest.mark.gpu

RHS_TOL = 1e-10
T_TOL = 1e-6
THW_TOL = 1e-8


def _oracle_diag(orc, tile, c, H, t):
    oc = orc.OracleColumn(tile, c)
    dH, st = oc.rhs(H, t)
    return oc, dH


def _check_rhs(orc, tile, eng, H0, t):
    dH = eng.rhs(t)
    dk = tile.grid.dedges
    diags = {v: eng.diag(v, t) for v in ("T", "thw", "C", "kc", "k", "jH", "dHdT")}
    worst = 0.0
    for c in range(tile.M):
        oc, dref = _oracle_diag(orc, tile, c, H0[:, c], t)
        worst = max(worst, common.rhs_error(dH[:, c], dref, oc.work.get("jH"), dk))
        np.testing.assert_allclose(diags["T"][:, c], oc.work.get("T"), rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(diags["thw"][:, c], oc.work.get("thw"), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(diags["C"][:, c], oc.work.get("C"), rtol=1e-12)
        np.testing.assert_allclose(diags["kc"][:, c], oc.work.get("kc"), rtol=1e-12)
        np.testing.assert_allclose(diags["k"][:, c], oc.work.get("k"), rtol=1e-12)
        np.testing.assert_allclose(diags["dHdT"][:, c], oc.work.get("dHdT"), rtol=1e-10)
        jref = oc.work.get("jH")
        np.testing.assert_allclose(diags["jH"][:, c], jref, rtol=1e-11, atol=1e-11 * np.max(np.abs(jref)))
    assert worst < RHS_TOL, worst
    return worst


def _check_advance(orc, tile, eng, H0, t0, targets, dt0=60.0, dtmin=1.0, dtmax=3600.0, cols=None, envelope=False):
    """Advance the engine and the oracle through `targets` and compare at every save point.

    Strict mode (envelope=False): equal step counts, |dT| <= 1e-6 K, |dθw| <= 1e-8 -- the north_star gate.  It holds
    whenever the explicit scheme is stable (dtmax below the diffusion limit of the finest cells).
    Envelope mode: CGEuler with dtmax=3600 s on a 5 cm grid is *unstable* in the deep fine cells whenever MaxDelta lets
    dt exceed ~800 s (DESIGN.md "sensitivity"): the reference's own trajectory then amplifies a 1-ulp perturbation
    of u0 to ~0.1 K within days.  There the engine must stay within the oracle's own perturbation envelope."""
    eng.set_state(H0, t0)
    cols = list(range(tile.M) if cols is None else cols)
    ref = {c: (H0[:, c].copy(), t0, dt0, 0) for c in cols}
    rng = np.random.default_rng(99)
    pert = {c: [(H0[:, c] * (1.0 + s * 2.2e-16 * rng.integers(1, 4, H0.shape[0])), t0, dt0) for s in (1.0, -1.0)] for c in cols} if envelope else {}
    worst = {"T": 0.0, "thw": 0.0}
    for tt in targets:
        eng.advance_to(tt)
        H, t, dt = eng.get_state(with_time=True)
        st, steps = eng.status()
        Tg, thwg = eng.diag("T", tt), eng.diag("thw", tt)
        assert not st.any()
        for c in cols:
            oc = orc.OracleColumn(tile, c)
            Hr, tr, dtr, ns = ref[c]
            Hr, tr, dtr, n2, s2 = oc.advance(Hr, tr, dtr, tt, dtmin=dtmin, dtmax=dtmax)
            ref[c] = (Hr, tr, dtr, ns + n2)
            assert s2 == 0
            assert t[c] == tr == tt
            oc.rhs(Hr, tt)
            Tr, thr = oc.work.get("T"), oc.work.get("thw")
            eT, eth = np.max(np.abs(Tg[:, c] - Tr)), np.max(np.abs(thwg[:, c] - thr))
            if not envelope:
                assert steps[c] == ns + n2, (c, steps[c], ns + n2)
                assert eT <= T_TOL, (tt, c, eT)
                assert eth <= THW_TOL, (tt, c, eth)
                assert dt[c] == pytest.approx(dtr, rel=1e-9)
            else:
                envT, envth = 0.0, 0.0
                for k, (Hp, tp, dtp) in enumerate(pert[c]):
                    Hp, tp, dtp, _, _ = oc.advance(Hp, tp, dtp, tt, dtmin=dtmin, dtmax=dtmax)
                    pert[c][k] = (Hp, tp, dtp)
                    oc.rhs(Hp, tt)
                    envT = max(envT, np.max(np.abs(oc.work.get("T") - Tr)))
                    envth = max(envth, np.max(np.abs(oc.work.get("thw") - thr)))
                assert abs(steps[c] - (ns + n2)) <= max(5, 0.1 * steps[c])
                assert eT <= max(T_TOL, 100.0 * envT) and eT < 0.5, (tt, c, eT, envT)
                assert eth <= max(THW_TOL, 100.0 * envth) and eth < 0.05, (tt, c, eth, envth)
            worst["T"] = max(worst["T"], eT); worst["thw"] = max(worst["thw"], eth)
    return worst


def test_rhs_freewater_samoylov(orc):
    tile = common.samoylov_freew_tile(ncolumns=37, seed=1)       # ragged: not a multiple of 4 or 32
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    eng.set_state(H0, 0.0)
    _check_rhs(orc, tile, eng, H0, 0.0)
    _check_rhs(orc, tile, eng, H0, 100 * 86400.0 + 1234.5)      # summer forcing, between knots
    eng.close()


def test_rhs_freewater_mixed_phase(orc):
    # random enthalpies spanning frozen / phase-change / thawed cells
    tile = common.samoylov_freew_tile(ncolumns=16, seed=2)
    rng = np.random.default_rng(5)
    H0 = rng.uniform(-8e7, 3.5e8, size=(tile.grid.ncells, tile.M))
    H0[::7] = 0.0
    eng = cg.Engine(tile)
    eng.set_state(H0, 0.0)
    _check_rhs(orc, tile, eng, H0, 3600.0 * 5)
    eng.close()


@pytest.mark.parametrize("grid_name", ["DefaultGrid_10cm", "DefaultGrid_2cm", "DefaultGrid_20cm", "DefaultGrid_1m"])
def test_rhs_other_grids(orc, grid_name):
    tile = common.samoylov_freew_tile(ncolumns=5, seed=3, grid=getattr(cg, grid_name))
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    eng.set_state(H0, 0.0)
    _check_rhs(orc, tile, eng, H0, 7 * 86400.0)
    eng.close()


def test_advance_freewater_30_days_stable(orc):
    # cfg2 columns with dtmax = 600 s (below the diffusion limit of the 5 cm cells): strict gate at every save point
    tile = common.samoylov_freew_tile(ncolumns=12, seed=1)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=600.0)
    _check_advance(orc, tile, eng, H0, 0.0, [86400.0 * d for d in (1, 2, 3, 10, 30)], dtmax=600.0)
    eng.close()


def test_advance_freewater_one_year_stable(orc):
    # the north_star gate: <= 1e-6 K and <= 1e-8 liquid water after one simulated year, equal step counts
    # (4 columns; the oracle needs a few seconds each)
    tile = common.samoylov_freew_tile(ncolumns=4, seed=11)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=600.0)
    worst = _check_advance(orc, tile, eng, H0, 0.0, [86400.0 * d for d in (100, 200, 300, 365)], dtmax=600.0)
    print("one-year parity (stable cfg2):", worst)
    eng.close()


def test_advance_freewater_cfg2_as_specified(orc):
    # cfg2 exactly as specified (dtmax = 3600 s): day 1 is still strict; later the reference's own trajectory is
    # chaotic at ~3.5 m depth, so parity is checked against the oracle's 1-ulp perturbation envelope
    tile = common.samoylov_freew_tile(ncolumns=3, seed=1)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    _check_advance(orc, tile, eng, H0, 0.0, [86400.0 * d for d in (1, 2, 10, 30, 60)], envelope=True)
    eng.close()


def test_fourier_solution_on_gpu(orc):
    # test/Physics/Heat/heat_conduction_tests.jl:100-140 through the CUDA path
    tile = common.fourier_tile(ncolumns=3)
    xc = tile.grid.cells
    H0 = np.repeat(np.sin(2 * np.pi * xc)[:, None], 3, axis=1)
    eng = cg.Engine(tile, dt0=1e-5, dtmin=1e-5, dtmax=1e-5)
    eng.set_state(H0, 0.0)
    errs = []
    for s in range(1, 51):
        eng.advance_to(s * 0.01)
        H = eng.get_state()
        errs.append(np.abs(H[:, 0] - np.exp(-s * 0.01 * 4 * np.pi ** 2) * np.sin(2 * np.pi * xc)))
    errs = np.array(errs)
    assert np.max(np.mean(errs, axis=0)) < 1e-5
    oc = orc.OracleColumn(tile, 0)
    Hr, *_ = oc.advance(H0[:, 0], 0.0, 1e-5, 0.5, dtmin=1e-5, dtmax=1e-5)
    np.testing.assert_allclose(H[:, 1], Hr, rtol=0, atol=1e-11)
    eng.close()


def test_neumann_and_dirichlet_boundaries(orc):
    # Neumann top + Dirichlet bottom, periodic top
    grid = cg.DefaultGrid_20cm
    for top, bot in ((cg.ConstantBC(cg.Neumann, 10.0), cg.ConstantBC(cg.Dirichlet, 2.0)),
                     (cg.PeriodicBC(cg.Dirichlet, common.YEAR, 12.0, 0.3, -4.0), cg.ConstantBC(cg.Neumann, -0.05))):
        tile = cg.SoilHeatTile("H", top, bot, cg.SoilProfile((0.0, cg.SimpleSoil(por=0.4, org=0.1)), (2.0, cg.SimpleSoil(por=0.2))), None,
                               cg.initializer("T", cg.TemperatureProfile((0.0, -5.0), (1000.0, 8.0))), grid=grid, ncolumns=3)
        H0 = cg.initialcondition(tile)
        eng = cg.Engine(tile)
        eng.set_state(H0, 0.0)
        _check_rhs(orc, tile, eng, H0, 0.25 * common.YEAR)
        _check_advance(orc, tile, eng, H0, 0.0, [86400.0, 5 * 86400.0])   # 20 cm grid: stable at dtmax = 3600 s
        eng.close()


def test_cfl_limiter(orc):
    tile = common.samoylov_freew_tile(ncolumns=3, seed=8, grid=cg.DefaultGrid_20cm)
    tile.heat.dtlim = cg.CFL(0.5, cg.MaxDelta(5e4))
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    _check_advance(orc, tile, eng, H0, 0.0, [86400.0, 3 * 86400.0])
    eng.close()


def test_forcing_out_of_range_sets_status(orc):
    tile = common.samoylov_freew_tile(ncolumns=2, forcing_days=2)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    eng.set_state(H0, 0.0)
    eng.advance_to(5 * 86400.0)           # forcing ends after 2 days -> BoundsError in the reference
    st, _ = eng.status()
    assert np.all(st & L.STATUS_FORCING_RANGE)
    assert eng.stats()["status_or"] & L.STATUS_FORCING_RANGE
    eng.close()


def test_api_errors():
    tile = common.samoylov_freew_tile(ncolumns=2)
    eng = cg.Engine(tile)
    with pytest.raises(L.CgbError):
        eng.advance_to(10.0)              # no state yet
    with pytest.raises(L.CgbError):
        eng._param("no_such_param", [1.0], L.PER_GLOBAL)
    with pytest.raises(L.CgbError):
        eng._param("por", [1.0, 2.0, 3.0], L.PER_LAYER)
    with pytest.raises(L.CgbError):
        eng.set_state(cg.initialcondition(tile), 0.0); eng.diag("nope", 0.0)
    eng.close()


def test_idempotent_and_checkpoint(orc):
    tile = common.samoylov_freew_tile(ncolumns=6, seed=4)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    eng.set_state(H0, 0.0)
    eng.advance_to(86400.0)
    H1, t1, dt1 = eng.get_state(with_time=True)
    eng.advance_to(86400.0)               # already there: nothing happens
    H2, t2, dt2 = eng.get_state(with_time=True)
    np.testing.assert_array_equal(H1, H2); np.testing.assert_array_equal(t1, t2); np.testing.assert_array_equal(dt1, dt2)
    eng.close()


def test_solve_saveat_matches_stepwise(orc):
    tile = common.samoylov_freew_tile(ncolumns=9, seed=6)
    H0 = cg.initialcondition(tile)
    prob = cg.CryoGridProblem(tile, H0, (0.0, 5 * 86400.0), saveat=86400.0, savevars=("T", "H", "θw", "T_now", "θw_now"))
    sol = cg.solve(prob, cg.CGEuler(), dtmax=600.0)
    out = cg.CryoGridOutput(sol)
    assert out.T.shape == (6, tile.grid.ncells, 9) and sol.retcode == "Success"
    eng = cg.Engine(tile, dtmax=600.0)
    eng.set_state(H0, 0.0)
    np.testing.assert_array_equal(out.H[0], H0)
    np.testing.assert_array_equal(out.T[0], out.vars["T_now"][0])          # initial save: the cache IS the state
    for s in range(1, 6):
        eng.advance_to(s * 86400.0)
        np.testing.assert_array_equal(out.H[s], eng.get_state())
        np.testing.assert_array_equal(out.vars["T_now"][s], eng.diag("T", s * 86400.0))
        np.testing.assert_array_equal(out.vars["θw_now"][s], eng.diag("thw", s * 86400.0))
    eng.close()
    # "T"/"θw" are what the reference saves: the diagnostic cache of the last RHS evaluation, one step behind H
    # (Numerics/statevars.jl:60-106, basic_solvers.jl:61-66) -- exactly what the oracle's work arrays hold after an advance
    for c in (0, 4, 8):
        oc = orc.OracleColumn(tile, c)
        Hr, tr, dtr = H0[:, c].copy(), 0.0, 60.0
        for s in range(1, 6):
            Hr, tr, dtr, _, _ = oc.advance(Hr, tr, dtr, s * 86400.0, dtmax=600.0)
            assert np.max(np.abs(out.T[s][:, c] - oc.work.get("T"))) <= T_TOL
            assert np.max(np.abs(out.vars["θw"][s][:, c] - oc.work.get("thw"))) <= THW_TOL
            np.testing.assert_allclose(out.H[s][:, c], Hr, rtol=1e-12, atol=1e-3)
    assert np.max(np.abs(out.T[3] - out.vars["T_now"][3])) > 1e-9                # ... and it does differ from T(H saved)


def test_energy_conservation_large_batch():
    # size-independent property at a large batch: zero net boundary flux => integral of H conserved per column
    M = 4096
    rng = np.random.default_rng(0)
    soil = cg.SimpleSoil(por=rng.uniform(0.3, 0.7, M), org=rng.uniform(0, 0.3, M))
    tile = cg.SoilHeatTile("H", cg.ConstantBC(cg.Neumann, 10.0), cg.ConstantBC(cg.Neumann, 10.0), cg.SoilProfile((0.0, soil)), None,
                           cg.initializer("T", cg.TemperatureProfile((0.0, -20.0), (1000.0, 10.0))), grid=cg.DefaultGrid_5cm, ncolumns=M)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=900.0)
    eng.set_state(H0, 0.0)
    eng.advance_to(10 * 86400.0)
    H = eng.get_state()
    dk = tile.grid.dedges[:, None]
    e0, e1 = np.sum(H0 * dk, axis=0), np.sum(H * dk, axis=0)
    assert np.max(np.abs(e1 - e0) / np.abs(e0)) < 1e-9
    st = eng.stats()
    assert st["status_or"] == 0 and st["cell_steps"] == st["column_steps"] * tile.grid.ncells and st["min_steps"] > 900
    eng.close()


def test_ensemble_partial_sums():
    tile = common.samoylov_freew_tile(ncolumns=1000, seed=9)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    eng.set_state(H0, 0.0)
    s, q = eng.ensemble_partial("T", 0.0)
    T = eng.diag("T", 0.0)
    np.testing.assert_allclose(s, T.sum(axis=1), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(q, (T * T).sum(axis=1), rtol=1e-12, atol=1e-9)
    eng.close()


def test_reciprocal_sequences_accuracy():
    # the kernels replace IEEE division by MUFU.RCP64H + Newton; check the error against 1/x over the operand ranges
    # that occur (heat capacities ~1e6, conductivity denominators ~1e-2..1e2, |dH| ~1e-6..1e6)
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(1e5, 5e6, 20000), 10.0 ** rng.uniform(-8, 8, 20000), 1.0 + rng.uniform(0, 1, 20000)])
    lib = L.load()
    raw, n2, c3 = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    assert lib.cgb_selftest_rcp(L.dptr(x), x.size, L.dptr(raw), L.dptr(n2), L.dptr(c3)) == 0
    ref = 1.0 / x
    e_raw, e_n2, e_c3 = (np.max(np.abs(v - ref) / ref) for v in (raw, n2, c3))
    print(f"rcp rel. error: raw {e_raw:.3e} (2^{np.log2(e_raw):.1f}), newton2 {e_n2:.3e}, cubic {e_c3:.3e}")
    assert e_n2 < 3e-16          # <= ~1.3 ulp
    assert e_raw < 2.0 ** -16


def test_device_initialcondition_interp(orc):
    # InterpInitializer + initialize_soil_heat! on the device vs the host mirror and the oracle
    for tile in (common.samoylov_freew_tile(ncolumns=9, seed=3), common.sfcc_tile(ncolumns=5, kind="painterkarra", solver="lut")):
        H_host = cg.initialcondition(tile)
        eng = cg.Engine(tile)
        eng.initialcondition(0.0)
        H_dev, t, dt = eng.get_state(with_time=True)
        np.testing.assert_allclose(H_dev, H_host, rtol=1e-13, atol=1e-6)
        assert np.all(t == 0.0) and np.all(dt == 60.0)
        eng.advance_to(3600.0)            # the state is usable right away
        eng.close()
    # per-column temperature profiles
    tile = common.samoylov_freew_tile(ncolumns=4, seed=5)
    Tsurf = np.array([-5.0, -1.0, 0.5, 3.0])
    tile.inits = (cg.initializer("T", cg.TemperatureProfile((0.0, Tsurf), (10.0, -2.0), (1000.0, 10.0))),)
    eng = cg.Engine(tile)
    eng.initialcondition(0.0)
    np.testing.assert_allclose(eng.get_state(), cg.initialcondition(tile), rtol=1e-13, atol=1e-6)
    eng.close()


def _thawdepth_ref(T, z):
    # Diagnostics.thawdepth (Diagnostics/permafrost.jl:38-59) restated
    fr = np.nonzero(T <= 0.0)[0]; th = np.nonzero(T > 0.0)[0]
    if th.size and fr.size and th[0] < fr[0]:
        return (0.0 - T[th[0]]) * (z[fr[0]] - z[th[0]]) / (T[fr[0]] - T[th[0]]) + z[th[0]]
    return z[-1] if fr.size == 0 else z[0]


def test_device_output_path_select_and_thawdepth():
    # selected depths + thaw depth computed on the device == post-processing of the full output
    tile = common.samoylov_freew_tile(ncolumns=33, seed=7)
    H0 = cg.initialcondition(tile)
    days = np.array([120.0, 150.0, 180.0, 200.0]) * 86400.0
    eng = cg.Engine(tile, dtmax=600.0)
    eng.set_state(H0, 0.0)
    full = eng.solve_saveat(days, ("T", "θw"))
    z = tile.grid.cells
    cells = np.array([np.argmin(np.abs(z - d)) for d in (0.01, 0.1, 0.5, 1.0, 2.0, 10.0, 20.0)], dtype=np.int32)
    eng.set_state(H0, 0.0)
    sel, thaw = eng.solve_saveat_select(days, ("T", "θw"), cells, thawdepth=True)
    np.testing.assert_array_equal(sel, full[:, :, cells, :])
    ref = np.array([[_thawdepth_ref(full[s, 0, :, c], z) for c in range(tile.M)] for s in range(days.size)])
    np.testing.assert_allclose(thaw, ref, rtol=1e-13, atol=1e-13)
    assert thaw[0].max() > 0.1            # Tair peaks at day 91: the active layer is open at day 120
    eng.set_state(H0, 0.0)
    _, thaw2 = eng.solve_saveat_select(days, (), (), thawdepth=True)
    np.testing.assert_array_equal(thaw2, thaw)
    eng.close()
