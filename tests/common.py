"""Shared helpers for the tests: the synthetic problem builders live in the package (cryogrid_jl_b200.presets), so that
bench.py and benchmarks/ do not depend on the test tree; the same descriptions go to the oracle and to the CUDA engine."""
import numpy as np

import cryogrid_jl_b200 as cg
from cryogrid_jl_b200.presets import (YEAR, DAY, unit_soil_props, fourier_tile, samoylov_freew_tile, sfcc_tile,  # noqa: F401
                                      lite_ensemble_tile, cfg4_tile, cfg5_tile, rhs_error)
