"""cryogrid.jl_b200 -- B200-native batched soil-column engine behind CryoGrid.jl's solver API.

Only the heat-column hot path: model description mirror (model.py), solver front-end (solvers.py), and the
ctypes binding (engine.py, _lib.py) of the hand-written CUDA library csrc/ -> lib/libcgb200.so.
"""
from . import _lib, hostmath
from .grids import (Grid, DefaultGrid_1cm, DefaultGrid_2cm, DefaultGrid_5cm, DefaultGrid_10cm, DefaultGrid_20cm,
                    DefaultGrid_50cm, DefaultGrid_1m)
from .hostmath import ThermalProperties
from .model import (FreeWater, VanGenuchten, DallAmico, PainterKarra, DallAmicoSalt, SFCCPreSolver, SFCCNewtonSolver,
                    SimpleSoil, Heterogeneous, Profile, SoilProfile, TemperatureProfile, Input, Interpolated1D, NFactor, TemperatureBC,
                    ConstantBC, PeriodicBC, ConstantTemperature, ConstantFlux, GeothermalHeatFlux, MaxDelta, CFL,
                    HeatBalance, EnthalpyImplicit, Ground, Top, Bottom, Stratigraphy, ThermalSteadyStateInit, initializer,
                    Tile, SoilHeatTile, initialcondition, Dirichlet, Neumann)
from .engine import Engine
from .salt import (SaltProperties, SaltMassBalance, SaltGradient, SedimentCompactionInitializer, compaction, SalineTile,
                   SaltEngine)
from .solvers import CGEuler, LiteImplicitEuler, CryoGridProblem, CryoGridSolution, CryoGridOutput, solve

SamoylovDefault_tempprofile = TemperatureProfile(   # models.jl:35-44
    (0.0, -1.0), (0.2, -1.0), (2.0, -1.0), (5.0, -3.0), (9.5, -6.0), (25.0, -9.0), (100.0, -9.0), (1000.0, 10.2))

from . import presets  # noqa: E402  (needs the names above)
