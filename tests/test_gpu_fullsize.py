"""BASELINE.json's FULL batch sizes (cfg2 10^4, cfg3 / cfg4 / cfg5 10^5 columns) through properties that do not need the oracle on
every column:

  * duplicated members -- the second half of every ensemble is a random permutation of the first half (same parameters, same
    forcing offset, same initial state): a column's trajectory must not depend on where it sits in the batch, which warp or SM
    picked it from the work-stealing cursor, or what its neighbours do.  Gate: BIT-IDENTICAL state, time, dt and step count;
  * spot checks -- a handful of columns, drawn at random from the whole batch, against the oracle at the strict north_star
    gate and tighter (equal step counts, |dT| <= 1e-9 K, |dθw| <= 1e-11; measured 1e-13);
  * every column status 0 and cell_steps == column_steps x N.

Each test integrates the full batch for a few simulated days (seconds of GPU time)."""
import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg

pytestmark = pytest.mark.gpu
DAY = 86400.0
T_TOL, THW_TOL = 1e-9, 1e-11          # measured 1e-13; the north_star gate is 1e-6 K / 1e-8


def _perm(M, seed):
    half = M // 2
    return half, np.random.default_rng(seed).permutation(half)


def _dup(a, half, perm):
    a = np.asarray(a)
    if a.ndim == 0 or a.shape[-1] != 2 * half:
        return a
    a = a.copy()
    a[..., half:] = a[..., :half][..., perm]
    return a


def _check_duplicates(eng, half, perm):
    H, t, dt = eng.get_state(with_time=True)
    st, steps = eng.status()
    assert not st.any()
    assert np.array_equal(H[..., half:], H[..., :half][..., perm])
    assert np.array_equal(t[half:], t[:half][perm]) and np.array_equal(dt[half:], dt[:half][perm])
    assert np.array_equal(steps[half:], steps[:half][perm])
    s = eng.stats()
    assert s["status_or"] == 0 and s["cell_steps"] == s["column_steps"] * s["n_cells"]
    return H, t, dt, steps


def _spot_check(orc, tile, eng, H0, H, steps, targets, cols, dtmax):
    """`targets`: the times the engine was advanced to, in order (dt is clipped at every one of them: the oracle must stop there too)."""
    t_end = targets[-1]
    Tg, thwg = eng.diag("T", t_end), eng.diag("thw", t_end)
    worst = 0.0
    for c in cols:
        oc = orc.OracleColumn(tile, int(c))
        Hr, tr, dtr, ns = H0[:, c].copy(), 0.0, 60.0, 0
        for tt in targets:
            Hr, tr, dtr, n1, st = oc.advance(Hr, tr, dtr, tt, dtmax=dtmax)
            assert st == 0
            ns += n1
        assert ns == steps[c], (c, ns, steps[c])
        oc.rhs(Hr, t_end)
        eT, eth = np.max(np.abs(Tg[:, c] - oc.work.get("T"))), np.max(np.abs(thwg[:, c] - oc.work.get("thw")))
        assert eT <= T_TOL and eth <= THW_TOL, (c, eT, eth)
        worst = max(worst, eT)
    return worst


@pytest.mark.parametrize("variant", ["newton", "lut", "lut_n2"])
def test_fullsize_cfg4_sfcc(orc, variant):
    # cfg4 (per-column parameters, Newton on per-column curve tables) / cfg4' / cfg4'' (one soil class, SFCC fast path): 10^5 columns
    M = 100_000
    half, perm = _perm(M, 11)
    tile = common.cfg4_tile(M, variant)
    tile.top_offset = _dup(tile.top_offset, half, perm)
    _dup_parameters(tile, M, half, perm)
    H0 = cg.initialcondition(tile)
    H0 = _dup(H0, half, perm)
    eng = cg.Engine(tile, dtmax=600.0)
    eng.set_state(H0, 0.0)
    t_end = 2 * DAY
    eng.advance_to(DAY)
    eng.advance_to(t_end)                      # second launch: per-column curve tables come back from HBM
    H, t, dt, steps = _check_duplicates(eng, half, perm)
    cols = np.random.default_rng(5).integers(0, M, 5)
    w = _spot_check(orc, tile, eng, H0, H, steps, (DAY, t_end), cols, dtmax=600.0)
    print(f"cfg4 {variant}, {M} columns x {tile.grid.ncells} cells, 2 days: duplicates bit-identical, 5 random columns vs oracle {w:.2e} K, "
          f"{eng.stats()['cell_steps']:.3e} cell-steps")
    eng.close()


def _parameter_owners(tile):
    """Every object reachable from the tile that may hold per-column parameter arrays (soil, freeze curve, SWRC)."""
    seen, stack, out = set(), [tile], []
    while stack:
        o = stack.pop()
        if id(o) in seen or isinstance(o, (np.ndarray, str, bytes, int, float, bool, type(None))):
            continue
        seen.add(id(o))
        if isinstance(o, (list, tuple)):
            stack.extend(o)
            continue
        if isinstance(o, dict):
            stack.extend(o.values())
            continue
        if hasattr(o, "__dict__") and type(o).__module__.startswith("cryogrid_jl_b200"):
            out.append(o)
            stack.extend(vars(o).values())
    return out


def _dup_parameters(tile, M, half, perm):
    n = 0
    for obj in _parameter_owners(tile):
        for name, val in list(vars(obj).items()):
            if isinstance(val, np.ndarray) and val.ndim >= 1 and val.shape[-1] == M:
                object.__setattr__(obj, name, _dup(val, half, perm))
                n += 1
    return n


def test_fullsize_cfg2_freewater(orc):
    # cfg2 as specified: 10^4 FreeWater columns on DefaultGrid_5cm (the bench.py workload), 5 days
    M = 10_000
    half, perm = _perm(M, 12)
    tile = common.samoylov_freew_tile(ncolumns=M, seed=1)
    _dup_parameters(tile, M, half, perm)
    H0 = _dup(cg.initialcondition(tile), half, perm)
    eng = cg.Engine(tile, dtmax=600.0)
    eng.set_state(H0, 0.0)
    t_end = 5 * DAY
    eng.advance_to(t_end)
    H, t, dt, steps = _check_duplicates(eng, half, perm)
    cols = np.random.default_rng(6).integers(0, M, 5)
    w = _spot_check(orc, tile, eng, H0, H, steps, (t_end,), cols, dtmax=600.0)
    print(f"cfg2, {M} columns, 5 days: duplicates bit-identical, 5 random columns vs oracle {w:.2e} K")
    eng.close()


def test_fullsize_cfg3_lite(orc):
    # cfg3: 10^5-member LiteImplicitEuler ensemble (6 layers, per-member porosities and n-factors), 3 daily steps
    M = 100_000
    half, perm = _perm(M, 13)
    tile = common.lite_ensemble_tile(ncolumns=M, seed=1234, forcing_days=10)
    _dup_parameters(tile, M, half, perm)
    eng = cg.Engine(tile, stepper="lite")
    eng.initialcondition(0.0)                  # steady-state initial condition on the device
    H0 = eng.get_state()
    assert np.array_equal(H0[:, half:], H0[:, :half][:, perm])
    t_end = 3 * DAY
    eng.advance_to(t_end)
    H, t, dt, steps = _check_duplicates(eng, half, perm)
    assert (steps == 3).all()
    worst = 0.0
    for c in np.random.default_rng(7).integers(0, M, 4):
        Href, *_ = orc.OracleColumn(tile, int(c)).lite_advance(H0[:, c].copy(), 0.0, DAY, t_end)
        worst = max(worst, np.max(np.abs(H[:, c] - Href) / (np.abs(Href) + 1e3)))
    assert worst < 1e-8, worst
    print(f"cfg3, {M} columns x {tile.grid.ncells} cells, 3 daily steps: duplicates bit-identical, 4 random columns vs oracle {worst:.2e} (relative H)")
    eng.close()


def test_fullsize_cfg4_newton_two_kernels_agree():
    # 10^5 columns with per-column parameters through BOTH Newton kernels: the per-column curve tables (k_heat_newton_tab) against the
    # closed form with table-driven exp / log (k_heat_step<.,2>, CGB_NEWTON_TAB=0).  Equal step counts for every column, and the
    # temperatures within 1e-10 K (measured 6e-13 K: the two share the stopping rule, not the curve evaluation).
    import os
    M = 100_000
    tile = common.cfg4_tile(M, "newton")
    H0 = cg.initialcondition(tile)
    res = {}
    for name in ("tab", "general"):
        old = os.environ.get("CGB_NEWTON_TAB")
        if name == "general":
            os.environ["CGB_NEWTON_TAB"] = "0"
        try:
            eng = cg.Engine(tile, dtmax=600.0)
            eng.set_state(H0, 0.0)
            eng.advance_to(DAY)
            eng.advance_to(2 * DAY)
            st, steps = eng.status()
            assert not st.any()
            res[name] = (steps.copy(), eng.diag("T", 2 * DAY), eng.diag("thw", 2 * DAY))
            eng.close()
        finally:
            if old is None:
                os.environ.pop("CGB_NEWTON_TAB", None)
            else:
                os.environ["CGB_NEWTON_TAB"] = old
    assert np.array_equal(res["tab"][0], res["general"][0])
    dT = np.max(np.abs(res["tab"][1] - res["general"][1]))
    dth = np.max(np.abs(res["tab"][2] - res["general"][2]))
    print(f"cfg4, {M} columns, both Newton kernels: max |dT| {dT:.2e} K, max |dθw| {dth:.2e}, step counts equal")
    assert dT <= 1e-10 and dth <= 1e-12, (dT, dth)
