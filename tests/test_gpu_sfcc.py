"""GPU parity tests of the SFCC freeze curves (PainterKarra / DallAmico via the presolver LUT or on-device Newton)
against the CPU oracle.  Parity at the FreezeCurves.jl boundary is "unpinned" (SURVEY 8c): both sides restate the
published closed forms; the LUT knots are handed to both sides identically (cgb_set_sfcc_table)."""
import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg
from test_gpu_explicit import _check_rhs, _check_advance

pytestmark = pytest.mark.gpu


def _samoylov_sfcc_tile(M, solver="lut", seed=1):
    # SamoylovDefault-like 5-layer SFCC profile with explicit van Genuchten parameters (models.jl:26-45 passes preset
    # names whose (α, n) values live in FreezeCurves.jl; configs pass α, n explicitly -- SURVEY 8c)
    rng = np.random.default_rng(seed)
    spec = [(0.0, 0.80, 0.75, 4.0, 1.6), (0.1, 0.80, 0.25, 4.0, 1.6), (0.4, 0.55, 0.25, 2.0, 1.4), (3.0, 0.50, 0.0, 2.0, 1.4),
            (10.0, 0.30, 0.0, 4.06, 2.03)]
    fcsolver = cg.SFCCPreSolver() if solver == "lut" else cg.SFCCNewtonSolver()
    prof = [(z, cg.SimpleSoil(por=por, sat=1.0, org=org, freezecurve=cg.PainterKarra(swrc=cg.VanGenuchten(alpha=a, n=n))))
            for z, por, org, a, n in spec]
    ts = 3600.0 * np.arange(400 * 24 + 1)
    Tair = -10.0 + 18.0 * np.sin(2.0 * np.pi * ts / common.YEAR)
    forcings = cg.Interpolated1D(ts, Tair=Tair)
    nf = rng.uniform(0.4, 0.8, M)
    return cg.SoilHeatTile("H", cg.TemperatureBC(cg.Input("Tair"), cg.NFactor(nf=nf, nt=0.9)), cg.GeothermalHeatFlux(0.053),
                           cg.SoilProfile(*prof), forcings, cg.initializer("T", cg.SamoylovDefault_tempprofile),
                           grid=cg.DefaultGrid_10cm, fcsolver=fcsolver, ncolumns=M)


@pytest.mark.parametrize("kind", ["painterkarra", "dallamico"])
def test_rhs_sfcc_lut_single_layer(orc, kind):
    tile = common.sfcc_tile(ncolumns=6, kind=kind, solver="lut")
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=900.0)
    eng.set_state(H0, 0.0)
    _check_rhs(orc, tile, eng, H0, 0.0)
    # random enthalpies across the whole table range and beyond (extrapolation on both sides)
    rng = np.random.default_rng(3)
    H1 = rng.uniform(-2.5e8, 2.2e8, size=H0.shape)
    eng.set_state(H1, 0.0)
    _check_rhs(orc, tile, eng, H1, 0.0)
    eng.close()


def test_rhs_sfcc_lut_multilayer(orc):
    tile = _samoylov_sfcc_tile(7, "lut")
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    eng.set_state(H0, 0.0)
    _check_rhs(orc, tile, eng, H0, 50 * 86400.0)
    eng.close()


def test_rhs_sfcc_newton_per_column_parameters(orc):
    # cfg4 shape: per-column α, n, por, org -> no shared LUT; on-device Newton vs the oracle's tightly converged Newton
    tile = common.sfcc_tile(ncolumns=24, kind="dallamico", solver="newton", per_column=True, grid=cg.DefaultGrid_5cm,
                            top=cg.ConstantBC(cg.Dirichlet, -8.0), bottom=cg.GeothermalHeatFlux(0.053))
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile)
    eng.set_state(H0, 0.0)
    _check_rhs(orc, tile, eng, H0, 0.0)
    eng.close()


def test_advance_sfcc_lut_cfg1b(orc):
    # examples/heat_sfcc_constantbc.jl with the SFCC actually wired in (cfg1b): 10 W/m^2 in and out, dtmax = 900 s
    tile = common.sfcc_tile(ncolumns=3, kind="painterkarra", solver="lut")
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=900.0)
    _check_advance(orc, tile, eng, H0, 0.0, [3600.0, 86400.0, 10 * 86400.0, 30 * 86400.0], dtmax=900.0)
    H = eng.get_state()
    dk = tile.grid.dedges[:, None]
    e0, e1 = np.sum(H0 * dk, axis=0), np.sum(H * dk, axis=0)
    assert np.max(np.abs(e1 - e0) / np.abs(e0)) < 1e-9        # energy conservation (zero net flux)
    eng.close()


def test_advance_sfcc_multilayer_30_days(orc):
    tile = _samoylov_sfcc_tile(4, "lut", seed=2)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=900.0)      # below the diffusion limit of the 10 cm cells: strict gate
    _check_advance(orc, tile, eng, H0, 0.0, [86400.0, 10 * 86400.0, 30 * 86400.0], dtmax=900.0)
    eng.close()


def test_advance_sfcc_newton_10_days(orc):
    tile = common.sfcc_tile(ncolumns=5, kind="dallamico", solver="newton", per_column=True, grid=cg.DefaultGrid_10cm,
                            top=cg.PeriodicBC(cg.Dirichlet, common.YEAR, 15.0, 0.0, -6.0), bottom=cg.GeothermalHeatFlux(0.053),
                            temp=cg.TemperatureProfile((0.0, -6.0), (20.0, -4.0), (1000.0, 12.0)))
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=900.0)
    _check_advance(orc, tile, eng, H0, 0.0, [86400.0, 10 * 86400.0], dtmax=900.0)
    eng.close()


def test_caller_supplied_table_is_used(orc):
    # parity mode: the caller (e.g. the Julia glue) hands over its own presolver knots
    Hk = np.array([-3e8, -1e8, -1e7, 0.0, 1.67e8, 4e8])
    Tk = np.array([-80.0, -20.0, -1.0, -0.05, 0.0, 60.0])
    soil = cg.SimpleSoil(por=0.5, freezecurve=cg.PainterKarra(swrc=cg.VanGenuchten(alpha=1.0, n=2.0)))
    tile = cg.SoilHeatTile("H", cg.ConstantBC(cg.Neumann, 0.0), cg.ConstantBC(cg.Neumann, 0.0), cg.SoilProfile((0.0, soil)), None,
                           cg.initializer("T", -5.0), grid=cg.DefaultGrid_1m, fcsolver=cg.SFCCPreSolver(table=(Hk, Tk), extrap="flat"),
                           ncolumns=2)
    rng = np.random.default_rng(0)
    H = rng.uniform(-4e8, 5e8, size=(tile.grid.ncells, 2))
    eng = cg.Engine(tile)
    eng.set_state(H, 0.0)
    T = eng.diag("T", 0.0)
    np.testing.assert_allclose(T, np.interp(H, Hk, Tk), rtol=1e-13, atol=1e-13)
    _check_rhs(orc, tile, eng, H, 0.0)
    eng.close()


def test_exp_log_replacements_accuracy():
    # the van Genuchten powers use exp_fast / log_pos (cgb_common.cuh) instead of the CUDA library's exp / log / log1p;
    # check them against libm over the operand ranges that occur (x = α|ψ| from 1e-300 to 1e12, exponents to ±700)
    from cryogrid_jl_b200 import _lib as L
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-700.0, 700.0, 40000), rng.uniform(-2.0, 2.0, 20000), 10.0 ** rng.uniform(-300, 300, 20000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 20000), 1.0 + 10.0 ** rng.uniform(-16, 0, 20000)])
    ex, lg = np.empty_like(x), np.empty_like(x)
    assert L.load().cgb_selftest_explog(L.dptr(x), x.size, L.dptr(ex), L.dptr(lg)) == 0
    a = np.clip(x, -708.0, 709.0)
    e_exp = np.max(np.abs(ex - np.exp(a)) / np.exp(a))
    ref = np.log(np.abs(x))
    # log: relative error where |log| is not tiny, absolute error (in ulps of 1) near x = 1
    e_log = np.max(np.abs(lg - ref) / np.maximum(np.abs(ref), 1e-3))
    near1 = np.abs(ref) < 1e-3
    e_log1 = np.max(np.abs(lg[near1] - ref[near1]) / np.maximum(np.abs(ref[near1]), 1e-17))
    print(f"exp_fast rel. error {e_exp:.3e}, log_pos rel. error {e_log:.3e}, near 1: {e_log1:.3e}")
    assert e_exp < 4.5e-16
    assert e_log < 4.5e-16
    assert e_log1 < 6e-16
    # clamping: arguments beyond the range stay finite and monotone
    assert np.all(np.isfinite(ex)) and np.all(ex > 0.0)


def test_exp_log_table_variants_accuracy():
    # the STEPPING kernels evaluate the van Genuchten powers with table-driven exp / log (exp_tab / log_tab: 64-entry shared-memory
    # tables).  exp: relative error; log: ABSOLUTE error in units of max(1, |log x|) -- its results only ever feed exp(n log x) and
    # exp(-m log(1 + x^n)), where the absolute error of the logarithm is what matters.
    from cryogrid_jl_b200 import _lib as L
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-700.0, 700.0, 40000), rng.uniform(-2.0, 2.0, 20000), 10.0 ** rng.uniform(-300, 300, 20000),
                        1.0 + rng.uniform(-1e-3, 1e-3, 20000), 1.0 + 10.0 ** rng.uniform(-16, 0, 20000), 2.0 ** rng.integers(-1000, 1000, 2000)])
    ex, lg = np.empty_like(x), np.empty_like(x)
    assert L.load().cgb_selftest_explog_tab(L.dptr(x), x.size, L.dptr(ex), L.dptr(lg)) == 0
    a = np.clip(x, -708.0, 709.0)
    e_exp = np.max(np.abs(ex - np.exp(a)) / np.exp(a))
    ref = np.log(np.abs(x))
    e_log = np.max(np.abs(lg - ref) / np.maximum(np.abs(ref), 1.0))
    print(f"exp_tab rel. error {e_exp:.3e}, log_tab abs. error (units of max(1, |log|)) {e_log:.3e}")
    assert e_exp < 7e-16 and e_log < 5e-16
    assert np.all(np.isfinite(ex)) and np.all(ex > 0.0)


@pytest.mark.parametrize("solver", ["lut", "newton"])
def test_advance_sfcc_cfl_limiter(orc, solver):
    # CFL limiter on SFCC columns: the stepping kernel evaluates dH/dT only in this configuration (LUT mode)
    tile = _samoylov_sfcc_tile(3, solver, seed=5)
    tile.heat.dtlim = cg.CFL(0.5, cg.MaxDelta(5e4))
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=900.0)
    _check_advance(orc, tile, eng, H0, 0.0, [86400.0, 5 * 86400.0], dtmax=900.0)
    eng.close()


@pytest.mark.parametrize("solver", ["lut", "newton"])
def test_multi_target_saveat_equals_single_launches(solver):
    # device-side saveat of the general stepping kernel: several save times per launch (chunked path of cgb_solve_saveat)
    # must give exactly what one launch per save time gives (saved T/θw = cache of the last RHS evaluation of each interval)
    tile = _samoylov_sfcc_tile(5, solver, seed=7)
    H0 = cg.initialcondition(tile)
    days = 86400.0 * np.arange(1, 7)
    a = cg.Engine(tile, dtmax=900.0); a.set_state(H0, 0.0)
    multi = a.solve_saveat(days, ("T", "θw"))                       # chunked: no H, no *_now      -> [nsave, 2, N, M]
    b = cg.Engine(tile, dtmax=900.0); b.set_state(H0, 0.0)
    for s, t in enumerate(days):
        one = b.solve_saveat([t], ("T", "θw", "H"))                 # H in the list -> one launch per save
        if solver == "lut":
            np.testing.assert_array_equal(multi[s, 0], one[0, 0])
            np.testing.assert_array_equal(multi[s, 1], one[0, 1])
        else:
            # the on-device Newton stops at |ΔT| <= 1e-9 from a seed that a fresh launch resets (H/c_w) and a running
            # column carries over (first-order predictor): equal within the solver tolerance, not bitwise
            assert np.max(np.abs(multi[s, 0] - one[0, 0])) <= 1e-6
            assert np.max(np.abs(multi[s, 1] - one[0, 1])) <= 1e-8
        np.testing.assert_array_equal(one[0, 2], b.get_state())
    if solver == "lut":
        np.testing.assert_array_equal(a.get_state(), b.get_state())
    else:
        np.testing.assert_allclose(a.get_state(), b.get_state(), rtol=1e-9, atol=1.0)
    assert a.stats()["kernel_launches"] < b.stats()["kernel_launches"]
    a.close(); b.close()


@pytest.mark.parametrize("solver", ["lut", "newton"])
def test_advance_sfcc_one_year(orc, solver):
    # the north_star gate on SFCC columns: <= 1e-6 K and <= 1e-8 θw against the oracle after ONE simulated year (a full
    # freeze-thaw cycle: every cell of the active layer crosses the kink of its freeze curve twice), equal step counts.
    # dtmax = 900 s is below the diffusion limit of the 10 cm cells (DESIGN.md "sensitivity finding").
    tile = common.sfcc_tile(ncolumns=2, kind="dallamico", solver=solver, alpha=2.0, n=1.6, grid=cg.DefaultGrid_10cm,
                            top=cg.PeriodicBC(cg.Dirichlet, common.YEAR, 15.0, 0.0, -4.0), bottom=cg.GeothermalHeatFlux(0.053),
                            temp=cg.TemperatureProfile((0.0, -4.0), (20.0, -3.0), (1000.0, 12.0)))
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=900.0)
    worst = _check_advance(orc, tile, eng, H0, 0.0, [86400.0 * d for d in (120, 240, 365)], dtmax=900.0, cols=[0])
    print(f"one year SFCC ({solver}): max |dT| {worst['T']:.2e} K, max |dθw| {worst['thw']:.2e}")
    eng.close()
