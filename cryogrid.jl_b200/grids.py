"""Grid and the preset grids of the reference (src/Numerics/grid.jl:17-39, src/default_grids.jl)."""
from fractions import Fraction

import numpy as np


class Grid:
    """1-D grid defined by its edges; cells are midpoints (Numerics/grid.jl:24-33)."""

    def __init__(self, edges):
        e = np.asarray(edges, dtype=np.float64)
        if e.ndim != 1 or e.size < 3:
            raise ValueError("Grid needs at least 3 edges")
        if not np.all(np.diff(e) > 0):
            raise AssertionError("grid values should be in ascending order")
        self.edges = e.copy()
        self.cells = (e[:-1] + e[1:]) / 2.0
        self.dedges = e[1:] - e[:-1]            # Δ(grid): cell thicknesses
        self.dcells = self.cells[1:] - self.cells[:-1]  # Δ(cells(grid)): midpoint distances

    @property
    def ncells(self):
        return self.cells.size

    def __len__(self):
        return self.edges.size

    def __repr__(self):
        return f"Grid({self.edges[0]}..{self.edges[-1]}) of length {self.edges.size}"


def _jlrange(a, s, b):
    """Julia float range a:s:b -- elements are the correctly rounded values of the exact decimals."""
    A, S, B = Fraction(a), Fraction(s), Fraction(b)
    n = int((B - A) // S)
    return [float(A + i * S) for i in range(n + 1)]


_TAIL = _jlrange("21", "1", "30") + _jlrange("35", "5", "50") + _jlrange("60", "10", "100") + _jlrange("200", "100", "1000")

# src/default_grids.jl:1-79
DefaultGrid_1cm = Grid(_jlrange("0", "0.01", "0.5") + _jlrange("0.55", "0.02", "2") + _jlrange("2.05", "0.05", "4.0")
                       + _jlrange("4.1", "0.1", "10") + _jlrange("10.2", "0.2", "20") + _TAIL)
DefaultGrid_2cm = Grid(_jlrange("0", "0.02", "2") + _jlrange("2.05", "0.05", "4.0") + _jlrange("4.1", "0.1", "10")
                       + _jlrange("10.2", "0.2", "20") + _TAIL)
DefaultGrid_5cm = Grid(_jlrange("0.0", "0.05", "4.0") + _jlrange("4.1", "0.1", "10") + _jlrange("10.2", "0.2", "20") + _TAIL)
DefaultGrid_10cm = Grid(_jlrange("0.0", "0.1", "10") + _jlrange("10.2", "0.2", "20") + _TAIL)
DefaultGrid_20cm = Grid(_jlrange("0.0", "0.2", "20") + _TAIL)
DefaultGrid_50cm = Grid(_jlrange("0.0", "0.5", "20") + _TAIL)
DefaultGrid_1m = Grid(_jlrange("0.0", "1", "30") + _jlrange("35", "5", "50") + _jlrange("60", "10", "100")
                      + _jlrange("200", "100", "1000"))
