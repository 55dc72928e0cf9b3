"""Synthetic problem builders for the BASELINE.json configurations (SURVEY.md 8d "Synthetic inputs"): the same
descriptions go to the CUDA engine (bench.py, benchmarks/) and to the CPU oracle (tests/).  Pure host-side model
construction through the public API of this package -- nothing here touches the device or the oracle."""
import numpy as np

from . import (ThermalProperties, Grid, HeatBalance, SimpleSoil, Stratigraphy, Top, Bottom, ConstantBC, Dirichlet, Neumann, Ground,
               Tile, DefaultGrid_5cm, DefaultGrid_10cm, DefaultGrid_2cm, Interpolated1D, SoilHeatTile, TemperatureBC, Input, NFactor,
               GeothermalHeatFlux, SoilProfile, initializer, SamoylovDefault_tempprofile, VanGenuchten, PainterKarra, DallAmico,
               SFCCPreSolver, SFCCNewtonSolver, TemperatureProfile, ThermalSteadyStateInit, PeriodicBC, SalineTile, SaltGradient)

YEAR = 365 * 24 * 3600.0
DAY = 86400.0

def unit_soil_props():
    """Thermal properties that make T == H and k == 1 for a por=0 soil (used by the reference's
    TestGroundLayer-style tests, test/Physics/Heat/heat_conduction_tests.jl:100-140)."""
    return ThermalProperties(kh_w=0.0, kh_i=0.0, kh_a=0.0, kh_m=1.0, kh_o=0.0, ch_w=0.0, ch_i=0.0, ch_a=0.0, ch_m=1.0, ch_o=0.0)


def fourier_tile(ncolumns=1):
    grid = Grid(np.linspace(0.0, 1.0, 101))
    heat = HeatBalance(op="H", prop=unit_soil_props())
    soil = SimpleSoil(por=0.0, sat=1.0, org=0.0)
    strat = Stratigraphy((0.0, Top(ConstantBC(Dirichlet, 0.0))), [(0.0, Ground(soil, heat=heat))],
                            (1.0, Bottom(ConstantBC(Dirichlet, 0.0))))
    return Tile(strat, grid, None, ncolumns=ncolumns)


def samoylov_freew_tile(ncolumns=1, seed=1, jitter=0.05, grid=None, forcing_days=400, t0=0.0):
    """BASELINE cfg2: 5-layer FreeWater profile of examples/heat_freeW_samoylov.jl:17-23 with per-column
    porosity jitter U(+-jitter), synthetic sinusoidal Tair sampled hourly, NFactor(0.6, 0.9),
    GeothermalHeatFlux(0.053), SamoylovDefault temperature profile."""
    grid = grid or DefaultGrid_5cm
    rng = np.random.default_rng(seed)
    M = ncolumns
    base = [(0.0, 0.80, 0.75), (0.1, 0.80, 0.25), (0.4, 0.55, 0.25), (3.0, 0.50, 0.0), (10.0, 0.30, 0.0)]
    prof = []
    for z, por, org in base:
        p = por + (rng.uniform(-jitter, jitter, size=M) if jitter > 0 else np.zeros(M))
        prof.append((z, SimpleSoil(por=p, sat=1.0, org=org)))
    ts = t0 + 3600.0 * np.arange(forcing_days * 24 + 1)
    Tair = -10.0 + 18.0 * np.sin(2.0 * np.pi * (ts - t0) / YEAR)
    forcings = Interpolated1D(ts, Tair=Tair)
    tile = SoilHeatTile("H", TemperatureBC(Input("Tair"), NFactor(nf=0.6, nt=0.9)), GeothermalHeatFlux(0.053),
                           SoilProfile(*prof), forcings, initializer("T", SamoylovDefault_tempprofile), grid=grid,
                           ncolumns=M)
    return tile


def sfcc_tile(ncolumns=1, kind="painterkarra", solver="lut", alpha=1.0, n=2.0, grid=None, per_column=False, seed=4,
              top=None, bottom=None, temp=None, por=0.5, org=0.0):
    """cfg1b-style single-layer SFCC column (examples/heat_sfcc_constantbc.jl) or per-column parameters (cfg4)."""
    grid = grid or DefaultGrid_10cm
    M = ncolumns
    rng = np.random.default_rng(seed)
    if per_column:
        alpha = rng.uniform(0.5, 5.0, M); n = rng.uniform(1.3, 2.5, M); por = rng.uniform(0.3, 0.7, M); org = rng.uniform(0.0, 0.5, M)
    swrc = VanGenuchten(alpha=alpha, n=n)
    fc = PainterKarra(swrc=swrc) if kind == "painterkarra" else DallAmico(swrc=swrc)
    fcsolver = SFCCPreSolver() if solver == "lut" else SFCCNewtonSolver()
    soil = SimpleSoil(por=por, sat=1.0, org=org, freezecurve=fc)
    top = top or ConstantBC(Neumann, 10.0)
    bottom = bottom or ConstantBC(Neumann, 10.0)
    temp = temp or TemperatureProfile((0.0, -20.0), (1000.0, 10.0))
    return SoilHeatTile("H", top, bottom, SoilProfile((0.0, soil)), None, initializer("T", temp), grid=grid,
                           fcsolver=fcsolver, ncolumns=M)


def lite_ensemble_tile(ncolumns=1, seed=1234, grid=None, forcing_days=800, t0=0.0, fixed=False):
    """BASELINE cfg3: 6-layer profile + priors of examples/cglite_parameter_ensembles.jl:21-38, EnthalpyImplicit,
    FreeWater, ThermalSteadyStateInit(T0=-15), synthetic sinusoidal Tair."""
    grid = grid or DefaultGrid_2cm
    rng = np.random.default_rng(seed)
    M = ncolumns
    heat = HeatBalance(op="implicit")
    spec = [(0.0, (0.65, 0.95), 0.80, 0.75), (0.1, (0.65, 0.95), 0.80, 0.25), (0.4, (0.35, 0.75), 0.55, 0.25),
            (3.0, (0.30, 0.70), 0.50, 0.0), (10.0, (0.10, 0.50), 0.30, 0.0), (100.0, None, 0.05, 0.0)]
    layers = []
    for z, pr, por0, org in spec:
        por = por0 if (pr is None or fixed) else rng.uniform(pr[0], pr[1], M)
        layers.append((z, Ground(SimpleSoil(por=por, sat=1.0, org=org), heat=heat)))
    nf = 0.5 if fixed else rng.beta(1, 1, M)
    nt = 0.9 if fixed else rng.beta(1, 1, M)
    ts = t0 + 3600.0 * 24 * np.arange(forcing_days + 1)
    Tair = -10.0 + 18.0 * np.sin(2.0 * np.pi * (ts - t0) / YEAR)
    forcings = Interpolated1D(ts, Tair=Tair)
    strat = Stratigraphy((-2.0, Top(TemperatureBC(Input("Tair"), NFactor(nf=nf, nt=nt)))), layers,
                            (1000.0, Bottom(GeothermalHeatFlux(0.053))))
    return Tile(strat, grid, forcings, ThermalSteadyStateInit(T0=-15.0), ncolumns=M)


def cfg4_tile(ncolumns, variant="newton", seed=4, grid=None):
    """BASELINE cfg4 family on DefaultGrid_5cm: sinusoidal surface temperature (PeriodicBC) with a per-column level shift
    N(0, 2 K), GeothermalHeatFlux(0.053), SamoylovDefault temperature profile.
      "newton": DallAmico, per-column alpha, n, porosity, organic fraction, on-device Newton (cfg4 as specified)
      "lut":    DallAmico alpha = 2, n = 1.6 through the presolver LUT (cfg4')
      "lut_n2": PainterKarra alpha = 1, n = 2 through the presolver LUT (cfg4'', the curve of examples/heat_sfcc_constantbc.jl)"""
    grid = grid or DefaultGrid_5cm
    kw = dict(grid=grid, top=PeriodicBC(Dirichlet, YEAR, 18.0, 0.0, -10.0), bottom=GeothermalHeatFlux(0.053),
              temp=SamoylovDefault_tempprofile)
    if variant == "newton":
        tile = sfcc_tile(ncolumns=ncolumns, kind="dallamico", solver="newton", per_column=True, seed=seed, **kw)
    elif variant == "lut":
        tile = sfcc_tile(ncolumns=ncolumns, kind="dallamico", solver="lut", alpha=2.0, n=1.6, **kw)
    elif variant == "lut_n2":
        tile = sfcc_tile(ncolumns=ncolumns, kind="painterkarra", solver="lut", alpha=1.0, n=2.0, **kw)
    else:
        raise ValueError(variant)
    tile.top_offset = np.random.default_rng(seed + 1000).normal(0.0, 2.0, ncolumns)
    return tile


def cfg5_tile(ncolumns, seed=5):
    """BASELINE cfg5: examples/heat_sfcc_salt_constantbc.jl with per-column benthic salt U(600, 1000)."""
    rng = np.random.default_rng(seed)
    benthic = rng.uniform(600.0, 1000.0, ncolumns) if ncolumns > 1 else 900.0
    return SalineTile(grid=DefaultGrid_10cm, salt_bc=SaltGradient(benthic, 0), ncolumns=ncolumns)


def rhs_error(dH, dH_ref, jH_ref, dk):
    """SURVEY 8d parity metric of one RHS evaluation: |dH - ref| / max(|ref|, max_j |jH| / dk) per cell, maximum over cells."""
    scale = np.maximum(np.abs(dH_ref), np.max(np.abs(jH_ref)) / dk)
    return float(np.max(np.abs(dH - dH_ref) / scale))
