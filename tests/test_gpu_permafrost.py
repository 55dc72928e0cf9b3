"""Device-side permafrost diagnostics (cgb_solve_saveat_permafrost) against a NumPy restatement of Diagnostics/permafrost.jl applied
to the full saved temperature field of the same run (cgb_solve_saveat)."""
import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg

pytestmark = pytest.mark.gpu
DAY = 86400.0


def _thawdepth(T, z, Tmelt=0.0):                      # permafrost.jl:38-59, one time step
    fr = np.nonzero(T <= Tmelt)[0]; th = np.nonzero(T > Tmelt)[0]
    if th.size and fr.size and th[0] < fr[0]:
        i1, i2 = th[0], fr[0]
        return (Tmelt - T[i1]) * (z[i2] - z[i1]) / (T[i2] - T[i1]) + z[i1]       # solve_depth_linear, Diagnostics.jl:16
    return z[-1] if not fr.size else z[0]


def _reference(Tsave, z, Tmelt=0.0, thr=0.5, up=0.0, lo=10.0):
    """Tsave[nsave, N] of one column -> (alt, table, base, zaa, magt[nsave]) as permafrost.jl computes them for one year."""
    alt = max(_thawdepth(T, z, Tmelt) for T in Tsave)                               # :66-72
    amax, amin = Tsave.max(axis=0), Tsave.min(axis=0)
    mask = amax < Tmelt
    table = z[int(np.argmax(mask))]                                                 # :80-86 (argmax of all-false = first index)
    N = z.size
    r = int(np.argmax(mask[::-1])) + 1                                              # 1-based position from the bottom
    base = z[max(N - r, 1) - 1]                                                     # :95-101, verbatim index arithmetic
    A = amax - amin
    lo_i = np.nonzero(A < thr)[0]
    if lo_i.size and lo_i[0] > 0:
        i2 = lo_i[0]; i1 = i2 - 1
        zaa = (thr - A[i1]) * (z[i2] - z[i1]) / (A[i2] - A[i1]) + z[i1]             # :7-30
    elif lo_i.size:
        zaa = z[0]
    else:
        zaa = np.inf
    sel = (z >= up) & (z <= lo)
    magt = Tsave[:, sel].mean(axis=1)                                                # :108-113
    return alt, table, base, zaa, magt


def test_permafrost_diagnostics_match_the_reference_functions():
    M = 24
    tile = common.samoylov_freew_tile(ncolumns=M, seed=3, jitter=0.05)
    tile.top_offset = np.linspace(-6.0, 9.0, M)          # from cold permafrost to columns that thaw deeply / have no permafrost table near the top
    H0 = cg.initialcondition(tile)
    tsave = DAY * np.arange(5, 366, 5)
    eng = cg.Engine(tile, dtmax=600.0)
    eng.set_state(H0, 0.0)
    full = eng.solve_saveat(tsave, ("T",))[:, 0]          # [nsave, N, M]
    eng.set_state(H0, 0.0)
    red = eng.solve_saveat_permafrost(tsave, zaa_threshold=0.5, magt_upper=0.0, magt_lower=10.0)
    eng.close()
    z = tile.grid.cells
    for c in range(M):
        alt, table, base, zaa, magt = _reference(full[:, :, c], z)
        assert red["alt"][c] == pytest.approx(alt, rel=1e-13), c
        assert red["pf_table"][c] == table and red["pf_base"][c] == base, c
        assert (np.isinf(zaa) and np.isinf(red["zaa"][c])) or red["zaa"][c] == pytest.approx(zaa, rel=1e-12), c
        np.testing.assert_allclose(red["magt"][:, c], magt, rtol=1e-13, atol=1e-13)
    assert red["alt"].max() > 0.3 and red["alt"].min() < red["alt"].max()       # the ensemble spans different thaw depths
