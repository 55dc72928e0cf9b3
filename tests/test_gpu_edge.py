"""Edge cases through the C ABI on the GPU: smallest and largest grids the kernels are built for, a single column,
many layers, ragged column counts -- every cells-per-lane instantiation of every kernel family against the oracle."""
import os

import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg
from cryogrid_jl_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _grid(n):
    # geometric spacing from 2 cm at the top, n cells
    dz = 0.02 * 1.02 ** np.arange(n)
    return cg.Grid(np.concatenate([[0.0], np.cumsum(dz)]))


def _tile(n, M, op="H", nlayers=3, seed=0):
    rng = np.random.default_rng(seed)
    grid = _grid(n)
    heat = cg.HeatBalance(op=op)
    zs = np.linspace(0.0, grid.edges[-1], nlayers + 1)[:-1]
    layers = [(z, cg.Ground(cg.SimpleSoil(por=rng.uniform(0.2, 0.7, M), org=rng.uniform(0.0, 0.4)), heat=heat)) for z in zs]
    strat = cg.Stratigraphy((0.0, cg.Top(cg.TemperatureBC(cg.PeriodicBC(cg.Dirichlet, 86400.0 * 20, 9.0, 0.0, -1.0), cg.NFactor(0.7, 0.9)))),
                            layers, (grid.edges[-1], cg.Bottom(cg.GeothermalHeatFlux(0.05))))
    return cg.Tile(strat, grid, None, cg.initializer("T", cg.TemperatureProfile((0.0, -3.0), (grid.edges[-1], 2.0))), ncolumns=M)


@pytest.mark.parametrize("n", [2, 3, 31, 32, 33, 64, 65, 100, 160, 193, 224, 225, 256, 290, 320, 384, 385, 512])
def test_explicit_all_grid_sizes(orc, n):
    M = 5
    tile = _tile(n, M, nlayers=min(3, n))
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=30.0, dtmin=0.5)        # dtmax below the diffusion limit of the 2 cm cells
    eng.set_state(H0, 0.0)
    dH = eng.rhs(100.0)
    tt = 3600.0
    eng.advance_to(tt)
    H, t, dt = eng.get_state(with_time=True)
    st, steps = eng.status()
    assert not st.any() and np.all(t == tt)
    for c in range(M):
        oc = orc.OracleColumn(tile, c)
        dref, _ = oc.rhs(H0[:, c], 100.0)
        assert common.rhs_error(dH[:, c], dref, oc.work.get("jH"), tile.grid.dedges) < 1e-10
        Hr, tr, dtr, ns, s2 = oc.advance(H0[:, c], 0.0, 60.0, tt, dtmin=0.5, dtmax=30.0)
        assert steps[c] == ns
        np.testing.assert_allclose(H[:, c], Hr, rtol=1e-11, atol=1e-4)
    eng.close()


@pytest.mark.parametrize("n", [2, 33, 100, 224, 290, 512])
@pytest.mark.parametrize("variant", ["warp", "thread"])
def test_lite_all_grid_sizes(orc, n, variant):
    os.environ["CGB_LITE_VARIANT"] = variant
    try:
        M = 4
        tile = _tile(n, M, op="implicit", nlayers=min(3, n), seed=1)
        H0 = cg.initialcondition(tile)
        eng = cg.Engine(tile, stepper="lite", dt0=3600.0)
        eng.set_state(H0, 0.0)
        eng.advance_to(40 * 3600.0)
        H = eng.get_state()
        st, steps = eng.status()
        assert not st.any() and np.all(steps == 40)
        for c in range(M):
            Hr, *_ = orc.OracleColumn(tile, c).lite_advance(H0[:, c], 0.0, 3600.0, 40 * 3600.0)
            np.testing.assert_allclose(H[:, c], Hr, rtol=1e-9, atol=1e-2)
        eng.close()
    finally:
        del os.environ["CGB_LITE_VARIANT"]


def test_too_many_cells_is_refused():
    import ctypes as C
    lib = L.load()
    cfg = L.Config(abi_version=L.ABI_VERSION, device=0, n_cells=513, n_layers=1, n_columns=4, maxdelta=5e4, courant=0.5,
                   dt0=60.0, dtmin=1.0, dtmax=3600.0, lite_miniters=2, lite_maxiters=1000, lite_tol=1e-3)
    h = C.c_void_p()
    assert lib.cgb_create(C.byref(cfg), C.byref(h)) == L.ERR_UNSUPPORTED
    assert b"512" in lib.cgb_last_error(None)


def test_many_layers_and_single_column(orc):
    tile = _tile(128, 1, nlayers=64, seed=3)           # many layers, one column
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=30.0)
    eng.set_state(H0, 0.0)
    eng.advance_to(1800.0)
    H = eng.get_state()
    Hr, *_ = orc.OracleColumn(tile, 0).advance(H0[:, 0], 0.0, 60.0, 1800.0, dtmax=30.0)
    np.testing.assert_allclose(H[:, 0], Hr, rtol=1e-11, atol=1e-4)
    eng.close()


def test_large_ragged_batch_is_deterministic():
    # 100 003 columns (not a multiple of anything): two runs agree bit for bit although columns are scheduled dynamically
    M = 100003
    tile = common.samoylov_freew_tile(ncolumns=M, seed=2, forcing_days=3)
    H0 = cg.initialcondition(tile)
    res = []
    for _ in range(2):
        eng = cg.Engine(tile)
        eng.set_state(H0, 0.0)
        eng.advance_to(0.5 * 86400.0)
        res.append(eng.get_state())
        s = eng.stats()
        assert s["status_or"] == 0 and s["n_columns"] == M
        eng.close()
    np.testing.assert_array_equal(res[0], res[1])


def test_heterogeneous_per_cell_parameters(orc):
    # Heterogeneous{SimpleSoil} -- one parameter set per cell (n_layers == n_cells), explicit and implicit
    grid = cg.DefaultGrid_20cm
    prof = cg.SoilProfile((0.0, cg.SimpleSoil(por=np.array([0.8, 0.6]), org=0.7)), (1.0, cg.SimpleSoil(por=0.5, org=0.2)),
                          (10.0, cg.SimpleSoil(por=0.3, org=0.0)), (1000.0, cg.SimpleSoil(por=0.05, org=0.0)))
    for op, stepper in (("H", "euler"), ("implicit", "lite")):
        heat = cg.HeatBalance(op=op)
        strat = cg.Stratigraphy((0.0, cg.Top(cg.TemperatureBC(cg.PeriodicBC(cg.Dirichlet, common.YEAR, 15.0, 0.0, -5.0)))),
                                [(0.0, cg.Ground(cg.Heterogeneous(prof), heat=heat))], (1000.0, cg.Bottom(cg.GeothermalHeatFlux(0.053))))
        tile = cg.Tile(strat, grid, None, cg.initializer("T", cg.TemperatureProfile((0.0, -6.0), (1000.0, 10.0))), ncolumns=2)
        assert tile.n_layers == grid.ncells and np.array_equal(tile.cell_layer, np.arange(grid.ncells))
        assert tile.por[0, 0] == pytest.approx(np.interp(grid.cells[0], [0.0, 1.0], [0.8, 0.5]))
        H0 = cg.initialcondition(tile)
        eng = cg.Engine(tile, stepper=stepper)
        eng.set_state(H0, 0.0)
        tt = 5 * 86400.0
        eng.advance_to(tt)
        H = eng.get_state()
        assert not eng.status()[0].any()
        for c in range(2):
            oc = orc.OracleColumn(tile, c)
            if stepper == "euler":
                Hr, *_ = oc.advance(H0[:, c], 0.0, 60.0, tt)
            else:
                Hr, *_ = oc.lite_advance(H0[:, c], 0.0, 86400.0, tt)
            np.testing.assert_allclose(H[:, c], Hr, rtol=1e-9, atol=1e-2)
        eng.close()


@pytest.mark.parametrize("n", [2, 31, 33, 65, 100, 225, 290, 385, 512])
@pytest.mark.parametrize("kind", ["newton", "lut", "mixed"])
def test_sfcc_all_grid_sizes(orc, n, kind):
    # every cells-per-lane instantiation of the general stepping kernel: SFCC + Newton (interleaved cell->lane mapping),
    # SFCC + presolver LUT (contiguous mapping) and a FreeWater / SFCC-Newton stratigraphy (generic mode, interleaved)
    M = 3
    grid = _grid(n)
    fc = cg.DallAmico(swrc=cg.VanGenuchten(alpha=1.5, n=1.7))
    top_soil = cg.SimpleSoil(por=0.55, sat=1.0, org=0.1, freezecurve=fc) if kind != "mixed" else cg.SimpleSoil(por=0.55, org=0.1)
    deep_soil = cg.SimpleSoil(por=0.35, sat=1.0, org=0.0, freezecurve=cg.PainterKarra(swrc=cg.VanGenuchten(alpha=0.8, n=2.2)))
    zmid = grid.edges[max(1, n // 3)]
    prof = [(0.0, top_soil)] + ([(zmid, deep_soil)] if n >= 3 else [])
    fcsolver = cg.SFCCPreSolver() if kind == "lut" else cg.SFCCNewtonSolver()
    tile = cg.SoilHeatTile("H", cg.PeriodicBC(cg.Dirichlet, 86400.0 * 20, 9.0, 0.0, -1.0), cg.GeothermalHeatFlux(0.05), cg.SoilProfile(*prof), None,
                           cg.initializer("T", cg.TemperatureProfile((0.0, -3.0), (grid.edges[-1], 1.0))), grid=grid, fcsolver=fcsolver, ncolumns=M)
    H0 = cg.initialcondition(tile)
    eng = cg.Engine(tile, dtmax=30.0, dtmin=0.5)
    eng.set_state(H0, 0.0)
    dH = eng.rhs(100.0)
    tt = 1800.0
    eng.advance_to(tt)
    H, t, dt = eng.get_state(with_time=True)
    st, steps = eng.status()
    assert not st.any() and np.all(t == tt)
    for c in (0, M - 1):
        oc = orc.OracleColumn(tile, c)
        dref, _ = oc.rhs(H0[:, c], 100.0)
        assert common.rhs_error(dH[:, c], dref, oc.work.get("jH"), tile.grid.dedges) < 1e-10
        Hr, tr, dtr, ns, s2 = oc.advance(H0[:, c], 0.0, 60.0, tt, dtmin=0.5, dtmax=30.0)
        assert steps[c] == ns
        np.testing.assert_allclose(H[:, c], Hr, rtol=1e-9, atol=1e-2)
    eng.close()
