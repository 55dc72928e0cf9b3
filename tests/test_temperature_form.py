"""Temperature-based heat conduction, HeatBalance(:T): the prognostic state is T, the freeze curve is
evaluated directly (no inverse-enthalpy solve), du = dT = dH / dHdT, CFL limiter on dHdT / kc.
Reference: Physics/Heat/heat_conduction.jl:110-116 (variables), :147-152 (computeprognostic!), :79-84 (temperatureflux!),
:162-182 (CFL timestep with derivative = dT); Physics/Soils/soil_heat.jl:100-113 (freezethaw! incl. the default-saturation
quirk: sfcc(T; θsat) is called WITHOUT the soil's saturation); Physics/Heat/heat_types.jl:17 (default CFL(0.5, MaxDelta(Inf)))."""
import ctypes as C

import numpy as np
import pytest

import common
import cryogrid_jl_b200 as cg
from cryogrid_jl_b200 import hostmath as hm

DAY = 86400.0


def _tile(M=3, op="T", sat=0.8, grid=None, layers=1):
    rng = np.random.default_rng(3)
    if layers == 1:
        prof = [(0.0, cg.SimpleSoil(por=0.5, sat=sat, org=0.1, freezecurve=cg.PainterKarra(swrc=cg.VanGenuchten(alpha=1.0, n=2.0))))]
    else:
        prof = [(0.0, cg.SimpleSoil(por=0.8 + rng.uniform(-0.05, 0.05, M), sat=1.0, org=0.5, freezecurve=cg.DallAmico(swrc=cg.VanGenuchten(alpha=4.0, n=1.6)))),
                (0.4, cg.SimpleSoil(por=0.55, sat=sat, org=0.25, freezecurve=cg.DallAmico(swrc=cg.VanGenuchten(alpha=2.0, n=1.4)))),
                (3.0, cg.SimpleSoil(por=0.3, sat=1.0, org=0.0, freezecurve=cg.PainterKarra(swrc=cg.VanGenuchten(alpha=4.06, n=2.03))))]
    return cg.SoilHeatTile(op, cg.PeriodicBC(cg.Dirichlet, common.YEAR, 12.0, 0.3, -4.0), cg.GeothermalHeatFlux(0.053), cg.SoilProfile(*prof), None,
                           cg.initializer("T", cg.TemperatureProfile((0.0, -1.0), (10.0, 0.5), (1000.0, 10.0))), grid=grid or cg.DefaultGrid_10cm,
                           ncolumns=M)


def test_temperature_form_oracle_semantics(orc):
    tile = _tile()
    assert tile.tform and isinstance(tile.heat.dtlim, cg.CFL) and np.isinf(tile.heat.dtlim.maxdelta.dmax)   # heat_types.jl:17
    T0 = cg.initialcondition(tile)
    np.testing.assert_allclose(T0[:, 0], np.interp(tile.grid.cells, [0.0, 10.0, 1000.0], [-1.0, 0.5, 10.0]))     # the state IS T
    oc = orc.OracleColumn(tile, 0)
    dT, st = oc.rhs(T0[:, 0], 0.0)
    w = oc.work
    # dT = dH / dHdT (heat_conduction.jl:79-84); dHdT = C + dθw/dT (L + T (c_w - c_i)) (heat_methods.jl:105)
    np.testing.assert_allclose(dT, w.get("dH") / w.get("dHdT"), rtol=1e-15)
    tp = tile.tp
    np.testing.assert_allclose(w.get("dHdT"), w.get("C") + w.get("dthwdT") * (tp.L + T0[:, 0] * (tp.ch_w - tp.ch_i)), rtol=1e-15)
    # default-saturation quirk: θw comes from the curve at saturation 1.0 although the soil has sat = 0.8
    thw1, _ = hm.sfcc_theta(2, T0[:, 0], 1.0, 0.5, 1.0, 2.0)
    thw08, _ = hm.sfcc_theta(2, T0[:, 0], 0.8, 0.5, 1.0, 2.0)
    np.testing.assert_allclose(w.get("thw"), thw1, rtol=1e-13)
    assert np.max(np.abs(thw1 - thw08)) > 1e-3
    # ... while the heat capacity uses the soil's actual water content (regular heatcapacity(), not sfccheatcap)
    thwi, thm, tho = hm.soil_fractions(0.5, 0.8, 0.1)
    np.testing.assert_allclose(w.get("C"), hm.heatcapacity(tp, w.get("thw"), thwi, thm, tho), rtol=1e-15)
    # CFL limiter: min over cells of 0.5 dHdT / kc dx^2 (heat_conduction.jl:162-182)
    dt = oc.timestep(T0[:, 0])
    assert dt == pytest.approx(np.min(0.5 * w.get("dHdT") / w.get("kc") * tile.grid.dedges ** 2), rel=1e-15)
    # the :T and :H forms describe the same physics: one small step moves the enthalpy by the same amount
    Tn, t, dtn, ns, st = oc.advance(T0[:, 0], 0.0, 1.0, 1.0, dtmin=1.0, dtmax=1.0)
    assert ns == 1 and st == 0
    dH_lin = w.get("dHdT") * (Tn - T0[:, 0])
    np.testing.assert_allclose(dH_lin, w.get("dH") * 1.0, rtol=1e-9, atol=1e-6)


def test_freewater_has_no_temperature_form(orc):
    tile = cg.SoilHeatTile("T", cg.ConstantBC(cg.Dirichlet, -5.0), cg.GeothermalHeatFlux(0.053), cg.SoilProfile((0.0, cg.SimpleSoil())), None,
                           cg.initializer("T", cg.TemperatureProfile((0.0, -1.0), (1000.0, 10.0))), grid=cg.DefaultGrid_1m, ncolumns=1)
    oc = orc.OracleColumn(tile, 0)
    dT, _ = oc.rhs(cg.initialcondition(tile)[:, 0], 0.0)
    assert np.all(np.isnan(dT))                      # the reference throws a MethodError here


@pytest.mark.gpu
@pytest.mark.parametrize("layers", [1, 3])
def test_temperature_form_gpu_parity(orc, layers):
    tile = _tile(M=5, layers=layers)
    T0 = cg.initialcondition(tile)
    T0[:, 1] -= 3.0; T0[:, 2] += 1.5            # colder / partly thawed columns
    eng = cg.Engine(tile)
    eng.set_state(T0, 0.0)
    dT = eng.rhs(0.0)
    diags = {v: eng.diag(v, 0.0) for v in ("T", "thw", "C", "dHdT", "dthwdT", "kc", "k", "jH", "dH", "Henth")}
    for c in range(tile.M):
        oc = orc.OracleColumn(tile, c)
        dref, _ = oc.rhs(T0[:, c], 0.0)
        scale = np.maximum(np.abs(dref), np.max(np.abs(oc.work.get("jH"))) / tile.grid.dedges / oc.work.get("dHdT"))
        assert np.max(np.abs(dT[:, c] - dref) / scale) <= 1e-10
        for v in ("T", "thw", "C", "dHdT", "kc", "k"):
            np.testing.assert_allclose(diags[v][:, c], oc.work.get(v), rtol=1e-11, atol=1e-14, err_msg=v)
        np.testing.assert_allclose(diags["dthwdT"][:, c], oc.work.get("dthwdT"), rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(diags["Henth"][:, c], T0[:, c] * oc.work.get("C") + tile.tp.L * oc.work.get("thw"), rtol=1e-11)
    # 30 simulated days with the CFL limiter: north_star gates, equal step counts
    ref = {c: (T0[:, c].copy(), 0.0, 60.0, 0) for c in range(tile.M)}
    for tt in (DAY, 5 * DAY, 30 * DAY):
        eng.advance_to(tt)
        T, t, dt = eng.get_state(with_time=True)
        st, steps = eng.status()
        Wg = eng.diag("thw", tt)
        assert not st.any()
        for c in range(tile.M):
            oc = orc.OracleColumn(tile, c)
            Tr, tr, dtr, ns = ref[c]
            Tr, tr, dtr, n2, s2 = oc.advance(Tr, tr, dtr, tt)
            ref[c] = (Tr, tr, dtr, ns + n2)
            assert s2 == 0 and t[c] == tr == tt and steps[c] == ns + n2, (c, steps[c], ns + n2)
            oc.rhs(Tr, tt)
            assert np.max(np.abs(T[:, c] - Tr)) <= 1e-6 and np.max(np.abs(Wg[:, c] - oc.work.get("thw"))) <= 1e-8
            assert dt[c] == pytest.approx(dtr, rel=1e-9)
    # saved outputs through solve_saveat (device-side saveat) agree with stepwise advancing
    eng2 = cg.Engine(tile)
    eng2.set_state(T0, 0.0)
    out = eng2.solve_saveat(np.array([DAY, 5 * DAY]), ("T_now", "θw_now"))
    eng3 = cg.Engine(tile)
    eng3.set_state(T0, 0.0)
    eng3.advance_to(DAY); eng3.advance_to(5 * DAY)
    np.testing.assert_allclose(out[1, 0], eng3.get_state(), rtol=0, atol=1e-12)
    # on-device initial condition (cgb_init_interp): for the temperature form the interpolated profile IS the state
    eng3.initialcondition(0.0)
    np.testing.assert_allclose(eng3.get_state(), cg.initialcondition(tile), rtol=0, atol=1e-13)
    for e in (eng, eng2, eng3):
        e.close()


@pytest.mark.gpu
def test_temperature_form_rejects_unsupported_configurations():
    from cryogrid_jl_b200 import _lib as L
    tile = cg.SoilHeatTile("T", cg.ConstantBC(cg.Dirichlet, -5.0), cg.GeothermalHeatFlux(0.053), cg.SoilProfile((0.0, cg.SimpleSoil())), None,
                           cg.initializer("T", cg.TemperatureProfile((0.0, -1.0), (1000.0, 10.0))), grid=cg.DefaultGrid_1m, ncolumns=2)
    eng = cg.Engine(tile)
    eng.set_state(cg.initialcondition(tile), 0.0)
    with pytest.raises(L.CgbError) as e:
        eng.rhs(0.0)
    assert e.value.code == L.ERR_UNSUPPORTED          # FreeWater layer
    eng.close()
