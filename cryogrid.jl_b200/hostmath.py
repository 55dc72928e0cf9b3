"""Host-side (NumPy, vectorised over columns) pieces of the reference that run ONCE per problem, before
the device hot path: initial conditions and the SFCC presolver table.  Not a CPU fallback of the
stepping path -- stepping only exists in libcgb200.so.

References (relative to the CryoGrid.jl tree): Soils/soil_heat.jl:26-63 (initialize_soil_heat!),
initializers.jl:42-67 (InterpInitializer), Heat/heat_init.jl:58-85 (ThermalSteadyStateInit),
Heat/heat_methods.jl:40-105, Soils/para/simple.jl:19-45; FreezeCurves.jl 0.9 closed forms
(parity unpinned).
"""
import numpy as np

G = 9.80665       # Physics.jl:7
LSL = 3.34e5      # Physics.jl:5
RHO_W = 1000.0    # Physics.jl:4
R_GAS = 8.314459  # FreezeCurves.DallAmicoSalt default


class ThermalProperties:
    """Heat/thermal_properties.jl:10-19 + Soils/para/simple.jl:32-45 defaults."""

    def __init__(self, kh_w=0.57, kh_i=2.2, kh_a=0.025, kh_m=3.8, kh_o=0.25,
                 ch_w=4.2e6, ch_i=1.9e6, ch_a=0.00125e6, ch_m=2.0e6, ch_o=2.5e6, L=RHO_W * LSL):
        self.kh_w, self.kh_i, self.kh_a, self.kh_m, self.kh_o = kh_w, kh_i, kh_a, kh_m, kh_o
        self.ch_w, self.ch_i, self.ch_a, self.ch_m, self.ch_o = ch_w, ch_i, ch_a, ch_m, ch_o
        self.L = L

    def as_dict(self):
        return dict(kh_w=self.kh_w, kh_i=self.kh_i, kh_a=self.kh_a, kh_m=self.kh_m, kh_o=self.kh_o,
                    ch_w=self.ch_w, ch_i=self.ch_i, ch_a=self.ch_a, ch_m=self.ch_m, ch_o=self.ch_o, L=self.L)


def soil_fractions(por, sat, org):
    thwi = por * sat
    thm = (1.0 - por) * (1.0 - org)
    tho = (1.0 - por) * org
    return thwi, thm, tho


def heatcapacity(tp, thw, thwi, thm, tho):
    tha = 1.0 - thwi - thm - tho
    thi = thwi - thw
    return tp.ch_w * thw + tp.ch_i * thi + tp.ch_a * tha + tp.ch_m * thm + tp.ch_o * tho


def thermalconductivity(tp, thw, thwi, thm, tho):
    tha = 1.0 - thwi - thm - tho
    thi = thwi - thw
    s = (np.sqrt(tp.kh_w) * thw + np.sqrt(tp.kh_i) * thi + np.sqrt(tp.kh_a) * tha
         + np.sqrt(tp.kh_m) * thm + np.sqrt(tp.kh_o) * tho)
    return s * s


def sfcc_heatcap(tp, por, org, thw, thwi, thsat):
    """Soils/soil_heat.jl:7-17 (including the (1-θsat) quirk)."""
    thi = thwi - thw
    tha = thsat - thwi
    thm = (1.0 - por) * (1.0 - org) * (1.0 - thsat)
    tho = (1.0 - por) * org * (1.0 - thsat)
    return tp.ch_w * thw + tp.ch_i * thi + tp.ch_a * tha + tp.ch_m * thm + tp.ch_o * tho


def vg_theta(psi, thsat, thres, alpha, n):
    m = 1.0 - 1.0 / n
    x = np.abs(-alpha * np.minimum(psi, 0.0))
    base = 1.0 + x ** n
    th = thres + (thsat - thres) * base ** (-m)
    with np.errstate(divide="ignore", invalid="ignore"):
        dth = np.where(x > 0.0, (thsat - thres) * m * n * alpha * x ** (n - 1.0) * base ** (-m - 1.0), 0.0)
    pos = psi > 0.0
    return np.where(pos, thsat, th), np.where(pos, 0.0, dth)


def vg_psi(th, thsat, thres, alpha, n):
    m = 1.0 - 1.0 / n
    Se = np.minimum((th - thres) / (thsat - thres), 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        psi = -1.0 / alpha * (Se ** (-1.0 / m) - 1.0) ** (1.0 / n)
    return np.where(th < thsat, psi, 0.0)


def sfcc_theta(kind, T, sat, thsat, alpha, n, thres=0.0, Tm=0.0, beta=1.0, omega=1.0, saltconc=0.0):
    """θw(T) and ∂θw/∂T for DallAmico (1), PainterKarra (2), DallAmicoSalt (3)."""
    thtot = sat * thsat
    thsat_e = np.maximum(thtot, thsat)
    psi0 = vg_psi(thtot, thsat_e, thres, alpha, n)
    TmK = Tm + 273.15
    if kind != 2:
        beta, omega = 1.0, 1.0
    if kind == 3:
        Lf = LSL * RHO_W
        TmK = TmK + TmK / Lf * (-R_GAS * saltconc * TmK)
    gT = omega * G / LSL * psi0
    Tstar = TmK + gT * TmK
    TK = T + 273.15
    frozen = (Tstar - TK) >= 0.0
    cfac = beta * LSL / (G * Tstar)
    psi = np.where(frozen, psi0 + cfac * (TK - Tstar), psi0)
    dpsi = np.where(frozen, cfac, 0.0)
    thw, dth = vg_theta(psi, thsat_e, thres, alpha, n)
    return thw, dth * dpsi


def interp_profile(zk, vk, z):
    """InterpInitializer (initializers.jl:42-67): gridded-linear, flat extrapolation."""
    zk = np.asarray(zk, dtype=np.float64)
    vk = np.asarray(vk, dtype=np.float64)
    if zk.size == 1:
        return np.full_like(np.asarray(z, dtype=np.float64), vk[0])
    return np.interp(z, zk, vk)


def lut_eval(Hk, Tk, H, extrap_line=True):
    """Gridded(Linear) interpolant of the presolver cache; linear or flat extrapolation."""
    Hk = np.asarray(Hk); Tk = np.asarray(Tk)
    i = np.clip(np.searchsorted(Hk, H, side="right") - 1, 0, Hk.size - 2)
    w = (H - Hk[i]) / (Hk[i + 1] - Hk[i])
    if not extrap_line:
        w = np.clip(w, 0.0, 1.0)
    return (1.0 - w) * Tk[i] + w * Tk[i + 1]


def _sfcc_enthalpy(fc, tp, por, sat, org, T):
    thw, dth = sfcc_theta(fc["kind"], np.float64(T), sat, por, fc["alpha"], fc["n"], fc["theta_res"], fc["Tm"],
                          fc["beta"], fc["omega"])
    thw = float(thw); dth = float(dth)
    C = sfcc_heatcap(tp, por, org, thw, sat * por, por)
    return T * C + tp.L * thw, C + dth * (tp.L + T * (tp.ch_w - tp.ch_i))


def sfcc_newton(fc, tp, por, sat, org, H, T0):
    lo, hi = -273.0, 1.0e4
    T = T0 if (lo < T0 < hi) else 0.0
    for _ in range(200):
        Hv, dHdT = _sfcc_enthalpy(fc, tp, por, sat, org, T)
        r = Hv - H
        if r > 0.0:
            hi = T
        else:
            lo = T
        Tn = T - r / dHdT
        small = abs(Tn - T) <= 1e-15 * max(1.0, abs(Tn))
        safe = (lo < Tn < hi) or (small and lo <= Tn <= hi)     # see sfcc_newton_solve in csrc/cgb_euler.cuh
        if not safe:
            Tn = 0.5 * (lo + hi)
        done = abs(Tn - T) <= 1e-15 * max(1.0, abs(Tn))         # converged, or the bracket has collapsed
        T = Tn
        if done:
            break
    return T


def presolver_table(fc, tp, por, sat, org, Tmin=-60.0, errtol=1e-4, Tmax=0.0):
    """SFCCPreSolver(Cache1D) knot builder -- restatement of FreezeCurves.initialize! (call site
    Soils/soil_heat.jl:26-36).  Marches H from H(Tmin) to H(Tmax); first-order predictor vs nonlinear solve;
    ΔH starts at Lθtot/10 and is halved until the error is below errtol.  Returns (Hk, Tk)."""
    thtot = max(sat * por, 1e-8)
    H, dHdT = _sfcc_enthalpy(fc, tp, por, sat, org, Tmin)
    Hmax, _ = _sfcc_enthalpy(fc, tp, por, sat, org, Tmax)
    T = Tmin
    dTdH = 1.0 / dHdT
    Hk, Tk = [H], [T]
    while H < Hmax:
        eps = np.inf
        dH = tp.L * thtot / 10.0
        guard = 0
        while abs(eps) > errtol and guard < 200:
            guard += 1
            Test = T + dH * dTdH
            Tsol = sfcc_newton(fc, tp, por, sat, org, H + dH, Test)
            eps = abs(Tsol - Test)
            dH *= 0.5
        Hnew = H + dH
        Tnew = sfcc_newton(fc, tp, por, sat, org, Hnew, T + dH * dTdH)
        _, dHdT = _sfcc_enthalpy(fc, tp, por, sat, org, Tnew)
        dTdH = 1.0 / dHdT
        H, T = Hnew, Tnew
        Hk.append(H); Tk.append(T)
    return np.array(Hk), np.array(Tk)
