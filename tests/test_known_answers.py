"""Hand-derived known answers for the FreezeCurves.jl boundary (freeze curves are NOT vendored in the CryoGrid.jl tree and its
tests pin no numeric value for them: SURVEY.md 8c).  The literals below were derived from the PUBLISHED equations --
van Genuchten (1980), Dall'Amico et al. (2011, The Cryosphere 5, eq. 17-20), Painter & Karra (2014, Vadose Zone J. 13(4)) as
parameterised by FreezeCurves.jl (β, ω), and the freezing-point depression of FreezeCurves' DallAmicoSalt -- with 40-digit
arithmetic by scripts/derive_known_answers.py, independently of oracle/, hostmath.py and the CUDA kernels.  They pin the three
restatements (oracle C, host NumPy mirror, CUDA) to the equations; they do NOT pin them to FreezeCurves.jl's code: that needs
tests/golden/*.json from scripts/gen_golden.jl (tests/test_golden_fixtures.py).

Constants: g = 9.80665 m/s², Lsl = 3.34e5 J/kg, Tm = 273.15 K, R = 8.314459 J/K/mol, ρw = 1000 kg/m³, θres = 0.
Tuples are (θw, ∂θw/∂T [1/K], T* - 273.15 [°C])."""
import ctypes as C

import numpy as np
import pytest

import cryogrid_jl_b200 as cg
from cryogrid_jl_b200 import hostmath as hm

KNOWN = {
    "vg_theta(-1; a=1,n=2,ths=.5)": 0.35355339059327376,
    "vg_theta(-10; a=2,n=1.6,ths=.4)": 0.066084266011474735,
    "vg_psi(.25; a=1,n=2,ths=.5)": -1.7320508075688773,
}
# (kind, sat, θsat, α, n, β, ω, saltconc, T) -> (θw, ∂θw/∂T, T* in °C)
SFCC_KNOWN = [
    ((1, 1.0, 0.5, 1.0, 2.0, 1.0, 1.0, 0.0, -0.01), (0.3128234252813227, 19.037370637977422, 0.0)),
    ((1, 1.0, 0.5, 1.0, 2.0, 1.0, 1.0, 0.0, -0.1), (0.039971751996243991, 0.39716293974990404, 0.0)),
    ((1, 1.0, 0.5, 1.0, 2.0, 1.0, 1.0, 0.0, -1.0), (0.0040098806948338064, 0.0040096227930503895, 0.0)),
    ((1, 1.0, 0.5, 1.0, 2.0, 1.0, 1.0, 0.0, -10.0), (0.00040100083623133976, 4.0100057830492216e-5, 0.0)),
    ((1, 0.5, 0.5, 1.0, 2.0, 1.0, 1.0, 0.0, -0.005), (0.25, 0.0, -0.013891080912024497)),          # above T*: θw = θtot
    ((1, 0.5, 0.5, 1.0, 2.0, 1.0, 1.0, 0.0, -0.05), (0.079185141572145726, 1.544003563279553, -0.013891080912024497)),
    ((1, 0.5, 0.5, 1.0, 2.0, 1.0, 1.0, 0.0, -1.0), (0.0040096796172999942, 0.0040094245867016632, -0.013891080912024497)),
    ((1, 0.8, 0.4, 2.0, 1.6, 1.0, 1.0, 0.0, -2.0), (0.0096231672423880034, 0.0028868109886910264, -0.0035236460986599523)),
    ((2, 0.5, 0.5, 1.0, 2.0, 2.0, 0.5, 0.0, -1.0), (0.0020049380786648567, 0.0020049061951342601, -0.0069455404560122485)),
    ((2, 1.0, 0.5, 1.0, 2.0, 2.0, 0.5, 0.0, -1.0), (0.0020049887057508122, 0.0020049564656951471, 0.0)),
    ((3, 1.0, 0.4, 4.06, 2.03, 1.0, 1.0, 900.0, -1.0), (0.4, 0.0, -1.6716003307502537)),           # above the depressed melting point
    ((3, 1.0, 0.4, 4.06, 2.03, 1.0, 1.0, 900.0, -3.0), (0.00048618766161571419, 0.00037697419193712776, -1.6716003307502537)),
    ((3, 1.0, 0.4, 4.06, 2.03, 1.0, 1.0, 300.0, -3.0), (0.00026069984878091517, 0.00010992333188608284, -0.55720011025008458)),
]
RTOL = 2e-13     # the double-precision restatements go through two pow() calls: a few 1e-15 relative each
DRTOL = 2e-11    # ∂θw/∂T: T + 273.15 - T* cancels to ~0.04 K next to T* (conditioning ~1e-12 relative in double precision)


def _fc(orc, kind, alpha, n, beta, omega):
    return orc.fc_struct(dict(kind=kind, alpha=alpha, n=n, theta_res=0.0, Tm=0.0, beta=beta, omega=omega), solver=1)


def test_van_genuchten_known_answers(orc):
    l = orc.lib()
    assert l.cg_vg_theta(-1.0, 0.5, 0.0, 1.0, 2.0, None) == pytest.approx(KNOWN["vg_theta(-1; a=1,n=2,ths=.5)"], rel=RTOL)
    assert l.cg_vg_theta(-10.0, 0.4, 0.0, 2.0, 1.6, None) == pytest.approx(KNOWN["vg_theta(-10; a=2,n=1.6,ths=.4)"], rel=RTOL)
    assert l.cg_vg_psi(0.25, 0.5, 0.0, 1.0, 2.0) == pytest.approx(KNOWN["vg_psi(.25; a=1,n=2,ths=.5)"], rel=RTOL)
    assert l.cg_vg_theta(0.3, 0.5, 0.0, 1.0, 2.0, None) == 0.5            # ψ > 0: saturated
    th, _ = hm.vg_theta(np.float64(-1.0), 0.5, 0.0, 1.0, 2.0)
    assert float(th) == pytest.approx(KNOWN["vg_theta(-1; a=1,n=2,ths=.5)"], rel=RTOL)
    assert float(hm.vg_psi(np.float64(0.25), 0.5, 0.0, 1.0, 2.0)) == pytest.approx(KNOWN["vg_psi(.25; a=1,n=2,ths=.5)"], rel=RTOL)


@pytest.mark.parametrize("case", SFCC_KNOWN, ids=[f"kind{c[0][0]}_sat{c[0][1]}_b{c[0][5]}_c{c[0][7]}_T{c[0][8]}" for c in SFCC_KNOWN])
def test_sfcc_closed_forms_oracle_and_host(orc, case):
    (kind, sat, ths, a, n, beta, omega, c, T), (thw, dth, _) = case
    fc = _fc(orc, kind, a, n, beta, omega)
    dT, dc = C.c_double(0.0), C.c_double(0.0)
    v = orc.lib().cg_sfcc(C.byref(fc), T, sat, ths, c, C.byref(dT), C.byref(dc))
    assert v == pytest.approx(thw, rel=RTOL)
    assert dT.value == pytest.approx(dth, rel=DRTOL, abs=1e-300)
    hv, hd = hm.sfcc_theta(kind, np.float64(T), sat, ths, a, n, 0.0, 0.0, beta, omega, c)
    assert float(hv) == pytest.approx(thw, rel=RTOL)
    assert float(hd) == pytest.approx(dth, rel=DRTOL, abs=1e-300)


def test_dallamicosalt_dc_matches_finite_difference_of_the_known_form(orc):
    # ∂θw/∂c (what the reference gets from 2-D ForwardDiff duals, salt_diffusion.jl:17-22) against a central difference of the
    # closed form itself
    fc = _fc(orc, 3, 4.06, 2.03, 1.0, 1.0)
    for T, c in ((-3.0, 900.0), (-3.0, 300.0), (-2.0, 650.0)):
        dT, dc = C.c_double(0.0), C.c_double(0.0)
        orc.lib().cg_sfcc(C.byref(fc), T, 1.0, 0.4, c, C.byref(dT), C.byref(dc))
        h = 1e-3
        fd = (orc.lib().cg_sfcc(C.byref(fc), T, 1.0, 0.4, c + h, None, None) - orc.lib().cg_sfcc(C.byref(fc), T, 1.0, 0.4, c - h, None, None)) / (2 * h)
        assert dc.value == pytest.approx(fd, rel=1e-6)
        assert dc.value > 0.0          # more salt -> lower freezing point -> more liquid water at the same T


def test_freewater_known_answers(orc):
    # FreezeCurves.freewater (call site heat_conduction.jl:23): θw = θtot clamp(H / (L θtot), 0, 1), θtot floored at 1e-8
    L = 3.34e8
    for thtot, H, want in ((0.5, -1e7, 0.0), (0.5, 0.0, 0.0), (0.5, 0.25 * L, 0.25), (0.5, 0.5 * L, 0.5), (0.5, 3e8, 0.5),
                           (0.0, 1.0, 1e-8 * (1.0 / (L * 1e-8))), (0.0, 10.0, 1e-8)):
        out = C.c_double(0.0)
        orc.lib().cg_freewater(H, thtot, L, C.byref(out))
        assert out.value == pytest.approx(want, rel=1e-15, abs=1e-30)
    # enthalpyinv (heat_conduction.jl:39-57): T = H/C frozen, (H - Lθ)/C thawed, 0 in between
    assert orc.lib().cg_enthalpyinv_freewater(-2.0e6, 0.5, 2.0e6, L) == -1.0
    assert orc.lib().cg_enthalpyinv_freewater(0.5 * L + 3.0e6, 0.5, 3.0e6, L) == pytest.approx(1.0, rel=1e-15)
    assert orc.lib().cg_enthalpyinv_freewater(0.1 * L, 0.5, 3.0e6, L) == 0.0


@pytest.mark.gpu
def test_sfcc_known_answers_on_gpu():
    # the CUDA closed forms against the same literals: build H = T C(θw) + L θw from the KNOWN θw (soil_heat.jl:7-17 heat
    # capacity incl. its (1-θsat) quirk), let the engine invert it (on-device Newton, diagnostics mode = fully converged) and
    # read T, θw, ∂θw/∂T back through cgb_get_diag.
    tp = cg.ThermalProperties()
    for kind in (1, 2):
        cases = [c for c in SFCC_KNOWN if c[0][0] == kind]
        for (k, sat, ths, a, n, beta, omega, c, T), (thw, dth, _) in cases:
            swrc = cg.VanGenuchten(alpha=a, n=n)
            fc = cg.DallAmico(swrc=swrc) if kind == 1 else cg.PainterKarra(swrc=swrc, beta=beta, omega=omega)
            soil = cg.SimpleSoil(por=ths, sat=sat, org=0.0, freezecurve=fc)
            tile = cg.SoilHeatTile("H", cg.ConstantBC(cg.Neumann, 0.0), cg.ConstantBC(cg.Neumann, 0.0), cg.SoilProfile((0.0, soil)), None,
                                   cg.initializer("T", cg.TemperatureProfile((0.0, T), (1000.0, T))), grid=cg.DefaultGrid_1m,
                                   fcsolver=cg.SFCCNewtonSolver(), ncolumns=2)
            Cq = hm.sfcc_heatcap(tp, ths, 0.0, thw, sat * ths, ths)
            H = np.full((tile.grid.ncells, 2), T * Cq + tp.L * thw)
            eng = cg.Engine(tile)
            eng.set_state(H, 0.0)
            Tg, Wg, Dg = eng.diag("T", 0.0), eng.diag("thw", 0.0), eng.diag("dthwdT", 0.0)
            eng.close()
            if dth > 0.0:
                assert np.max(np.abs(Tg - T)) <= 1e-11 * max(1.0, abs(T) * 1e2), (kind, T, Tg[0, 0])
            assert np.max(np.abs(Wg / thw - 1.0)) <= 1e-11, (kind, T, Wg[0, 0], thw)
            assert np.max(np.abs(Dg - dth)) <= 1e-10 * max(dth, 1e-12) + 1e-300, (kind, T, Dg[0, 0], dth)
