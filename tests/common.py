"""Shared problem builders for the tests (same descriptions go to the oracle and to the CUDA engine)."""
import numpy as np

import cryogrid_jl_b200 as cg

YEAR = 365 * 24 * 3600.0


def unit_soil_props():
    """Thermal properties that make T == H and k == 1 for a por=0 soil (used by the reference's
    TestGroundLayer-style tests, test/Physics/Heat/heat_conduction_tests.jl:100-140)."""
    return cg.ThermalProperties(kh_w=0.0, kh_i=0.0, kh_a=0.0, kh_m=1.0, kh_o=0.0, ch_w=0.0, ch_i=0.0, ch_a=0.0, ch_m=1.0, ch_o=0.0)


def fourier_tile(ncolumns=1):
    grid = cg.Grid(np.linspace(0.0, 1.0, 101))
    heat = cg.HeatBalance(op="H", prop=unit_soil_props())
    soil = cg.SimpleSoil(por=0.0, sat=1.0, org=0.0)
    strat = cg.Stratigraphy((0.0, cg.Top(cg.ConstantBC(cg.Dirichlet, 0.0))), [(0.0, cg.Ground(soil, heat=heat))],
                            (1.0, cg.Bottom(cg.ConstantBC(cg.Dirichlet, 0.0))))
    return cg.Tile(strat, grid, None, ncolumns=ncolumns)


def samoylov_freew_tile(ncolumns=1, seed=1, jitter=0.05, grid=None, forcing_days=400, t0=0.0):
    """BASELINE cfg2: 5-layer FreeWater profile of examples/heat_freeW_samoylov.jl:17-23 with per-column
    porosity jitter U(+-jitter), synthetic sinusoidal Tair sampled hourly, NFactor(0.6, 0.9),
    GeothermalHeatFlux(0.053), SamoylovDefault temperature profile."""
    grid = grid or cg.DefaultGrid_5cm
    rng = np.random.default_rng(seed)
    M = ncolumns
    base = [(0.0, 0.80, 0.75), (0.1, 0.80, 0.25), (0.4, 0.55, 0.25), (3.0, 0.50, 0.0), (10.0, 0.30, 0.0)]
    prof = []
    for z, por, org in base:
        p = por + (rng.uniform(-jitter, jitter, size=M) if jitter > 0 else np.zeros(M))
        prof.append((z, cg.SimpleSoil(por=p, sat=1.0, org=org)))
    ts = t0 + 3600.0 * np.arange(forcing_days * 24 + 1)
    Tair = -10.0 + 18.0 * np.sin(2.0 * np.pi * (ts - t0) / YEAR)
    forcings = cg.Interpolated1D(ts, Tair=Tair)
    tile = cg.SoilHeatTile("H", cg.TemperatureBC(cg.Input("Tair"), cg.NFactor(nf=0.6, nt=0.9)), cg.GeothermalHeatFlux(0.053),
                           cg.SoilProfile(*prof), forcings, cg.initializer("T", cg.SamoylovDefault_tempprofile), grid=grid,
                           ncolumns=M)
    return tile


def sfcc_tile(ncolumns=1, kind="painterkarra", solver="lut", alpha=1.0, n=2.0, grid=None, per_column=False, seed=4,
              top=None, bottom=None, temp=None, por=0.5, org=0.0):
    """cfg1b-style single-layer SFCC column (examples/heat_sfcc_constantbc.jl) or per-column parameters (cfg4)."""
    grid = grid or cg.DefaultGrid_10cm
    M = ncolumns
    rng = np.random.default_rng(seed)
    if per_column:
        alpha = rng.uniform(0.5, 5.0, M); n = rng.uniform(1.3, 2.5, M); por = rng.uniform(0.3, 0.7, M); org = rng.uniform(0.0, 0.5, M)
    swrc = cg.VanGenuchten(alpha=alpha, n=n)
    fc = cg.PainterKarra(swrc=swrc) if kind == "painterkarra" else cg.DallAmico(swrc=swrc)
    fcsolver = cg.SFCCPreSolver() if solver == "lut" else cg.SFCCNewtonSolver()
    soil = cg.SimpleSoil(por=por, sat=1.0, org=org, freezecurve=fc)
    top = top or cg.ConstantBC(cg.Neumann, 10.0)
    bottom = bottom or cg.ConstantBC(cg.Neumann, 10.0)
    temp = temp or cg.TemperatureProfile((0.0, -20.0), (1000.0, 10.0))
    return cg.SoilHeatTile("H", top, bottom, cg.SoilProfile((0.0, soil)), None, cg.initializer("T", temp), grid=grid,
                           fcsolver=fcsolver, ncolumns=M)


def lite_ensemble_tile(ncolumns=1, seed=1234, grid=None, forcing_days=800, t0=0.0, fixed=False):
    """BASELINE cfg3: 6-layer profile + priors of examples/cglite_parameter_ensembles.jl:21-38, EnthalpyImplicit,
    FreeWater, ThermalSteadyStateInit(T0=-15), synthetic sinusoidal Tair."""
    grid = grid or cg.DefaultGrid_2cm
    rng = np.random.default_rng(seed)
    M = ncolumns
    heat = cg.HeatBalance(op="implicit")
    spec = [(0.0, (0.65, 0.95), 0.80, 0.75), (0.1, (0.65, 0.95), 0.80, 0.25), (0.4, (0.35, 0.75), 0.55, 0.25),
            (3.0, (0.30, 0.70), 0.50, 0.0), (10.0, (0.10, 0.50), 0.30, 0.0), (100.0, None, 0.05, 0.0)]
    layers = []
    for z, pr, por0, org in spec:
        por = por0 if (pr is None or fixed) else rng.uniform(pr[0], pr[1], M)
        layers.append((z, cg.Ground(cg.SimpleSoil(por=por, sat=1.0, org=org), heat=heat)))
    nf = 0.5 if fixed else rng.beta(1, 1, M)
    nt = 0.9 if fixed else rng.beta(1, 1, M)
    ts = t0 + 3600.0 * 24 * np.arange(forcing_days + 1)
    Tair = -10.0 + 18.0 * np.sin(2.0 * np.pi * (ts - t0) / YEAR)
    forcings = cg.Interpolated1D(ts, Tair=Tair)
    strat = cg.Stratigraphy((-2.0, cg.Top(cg.TemperatureBC(cg.Input("Tair"), cg.NFactor(nf=nf, nt=nt)))), layers,
                            (1000.0, cg.Bottom(cg.GeothermalHeatFlux(0.053))))
    return cg.Tile(strat, grid, forcings, cg.ThermalSteadyStateInit(T0=-15.0), ncolumns=M)


def rhs_error(dH, dH_ref, jH_ref, dk):
    """SURVEY 8d parity metric: |dH - ref| / max(|ref|, max_j |jH| / dk) per cell."""
    scale = np.maximum(np.abs(dH_ref), np.max(np.abs(jH_ref)) / dk)
    return np.max(np.abs(dH - dH_ref) / scale)
