#!/usr/bin/env python
"""Derives the hand-checkable known answers of tests/test_known_answers.py from the PUBLISHED equations with 40-digit
arithmetic (mpmath), independently of oracle/, hostmath.py and the CUDA kernels:

  van Genuchten (1980):           theta(psi) = theta_r + (theta_s - theta_r) (1 + (alpha |psi|)^n)^-(1 - 1/n),  psi <= 0
  Dall'Amico et al. (2011) eq 17-20:  T* = Tm + g Tm psi0 / Lf ;   psi(T) = psi0 + Lf / (g T*) (T - T*) H(T* - T)
  Painter & Karra (2014) as parameterised in FreezeCurves.jl: T* = Tm + omega g Tm psi0 / Lf ; psi(T) = psi0 + beta Lf / (g T*) (T - T*) H(T* - T)
  DallAmicoSalt (FreezeCurves.jl):  Tm <- Tm + Tm / (Lsl rho_w) (-R c Tm)  (freezing-point depression), then Dall'Amico

Prints a Python dict literal to paste into the test."""
import mpmath as mp

mp.mp.dps = 40
g, Lsl, TmK, R, rho_w = mp.mpf("9.80665"), mp.mpf("3.34e5"), mp.mpf("273.15"), mp.mpf("8.314459"), mp.mpf(1000)


def vg_theta(psi, ths, thr, a, n):
    if psi > 0:
        return ths, mp.mpf(0)
    m = 1 - 1 / n
    x = a * abs(psi)
    th = thr + (ths - thr) * (1 + x ** n) ** (-m)
    dth = (ths - thr) * m * n * a * x ** (n - 1) * (1 + x ** n) ** (-m - 1) if x > 0 else mp.mpf(0)   # d theta / d psi (psi < 0)
    return th, dth


def vg_psi(th, ths, thr, a, n):
    if th >= ths:
        return mp.mpf(0)
    m = 1 - 1 / n
    Se = (th - thr) / (ths - thr)
    return -(1 / a) * (Se ** (-1 / m) - 1) ** (1 / n)


def sfcc(T, sat, ths, a, n, beta=1, omega=1, c=None, slope_ref="Tstar"):
    a, n, ths, sat, T = (mp.mpf(str(v)) for v in (a, n, ths, sat, T))
    beta, omega = mp.mpf(str(beta)), mp.mpf(str(omega))
    Tm = TmK
    if c is not None:
        Tm = Tm + Tm / (Lsl * rho_w) * (-R * mp.mpf(str(c)) * Tm)
    psi0 = vg_psi(sat * ths, ths, 0, a, n)
    Tstar = Tm + omega * g * Tm / Lsl * psi0
    TK = T + TmK
    if TK <= Tstar:
        slope = beta * Lsl / (g * (Tstar if slope_ref == "Tstar" else omega * Tm if slope_ref == "omegaTm" else Tm))
        psi = psi0 + slope * (TK - Tstar)
    else:
        slope, psi = mp.mpf(0), psi0
    th, dth = vg_theta(psi, ths, 0, a, n)
    return th, dth * slope, Tstar - TmK


out = {}
out["vg_theta(-1; a=1,n=2,ths=.5)"] = vg_theta(mp.mpf(-1), mp.mpf("0.5"), 0, mp.mpf(1), mp.mpf(2))[0]
out["vg_theta(-10; a=2,n=1.6,ths=.4)"] = vg_theta(mp.mpf(-10), mp.mpf("0.4"), 0, mp.mpf(2), mp.mpf("1.6"))[0]
out["vg_psi(.25; a=1,n=2,ths=.5)"] = vg_psi(mp.mpf("0.25"), mp.mpf("0.5"), 0, mp.mpf(1), mp.mpf(2))
for T in ("-0.01", "-0.1", "-1", "-10"):
    th, d, _ = sfcc(T, 1, "0.5", 1, 2)
    out[f"dallamico sat=1 a=1 n=2 ths=.5 T={T}"] = (th, d)
for T in ("-0.005", "-0.05", "-1"):
    th, d, ts = sfcc(T, "0.5", "0.5", 1, 2)
    out[f"dallamico sat=.5 a=1 n=2 ths=.5 T={T}"] = (th, d, ts)
th, d, ts = sfcc("-2", "0.8", "0.4", 2, "1.6")
out["dallamico sat=.8 a=2 n=1.6 ths=.4 T=-2"] = (th, d, ts)
th, d, ts = sfcc("-1", "0.5", "0.5", 1, 2, beta=2, omega="0.5")
out["painterkarra b=2 w=.5 sat=.5 a=1 n=2 ths=.5 T=-1"] = (th, d, ts)
th2, d2, _ = sfcc("-1", "0.5", "0.5", 1, 2, beta=2, omega="0.5", slope_ref="Tm")
out["  ... same with the slope referred to Tm instead of T* (relative change of theta_w)"] = (th2 - th) / th
th, d, ts = sfcc("-1", 1, "0.5", 1, 2, beta=2, omega="0.5")
out["painterkarra b=2 w=.5 sat=1 a=1 n=2 ths=.5 T=-1"] = (th, d, ts)
for T, c in (("-1", 900), ("-3", 900), ("-3", 300)):
    th, d, ts = sfcc(T, 1, "0.4", "4.06", "2.03", c=c)
    out[f"dallamicosalt a=4.06 n=2.03 ths=.4 c={c} T={T}"] = (th, d, ts)
for k, v in out.items():
    if isinstance(v, tuple):
        print(f'    "{k}": ({", ".join(mp.nstr(x, 17) for x in v)}),')
    else:
        print(f'    "{k}": {mp.nstr(v, 17)},')
