// cgb_presolver.cuh -- SFCCPreSolver table builder on the device (remainder).
//
// The reference builds one presolver table per soil layer at initialisation: FreezeCurves.initialize!(solver, sfcc, hc; sat,
// θsat), call site Soils/soil_heat.jl:26-36 (default solver SFCCPreSolver(SFCCPreSolverCache1D()), Soils/ground.jl:27).  With
// per-member soil parameters every ensemble member owns its tables, and building them in a serial host loop (10^3 knots x a
// few safeguarded Newton solves each) is what made large LUT ensembles unusable in round 1.  Here: ONE THREAD PER PARAMETER CLASS
// (distinct (curve, α, n, θres, Tm, β, ω, por, sat, org) tuple, deduplicated on the host) marches H from H(Tmin) to H(Tmax)
// exactly like the CPU checker's table builder (tests/test_gpu_presolver.py): first-order predictor T + ΔH dT/dH, ΔH starting at Lθtot/10 and halved until the
// predictor matches the fully converged nonlinear solve within errtol.  The arithmetic is the same (pow, true divisions,
// the same safeguarded Newton to 1e-15), so knots agree with the host builders to rounding (CUDA pow vs libm pow: a few ulp).
#pragma once
#include "cgb_handle.h"
#include "cgb_common.cuh"

namespace cgb {

struct PresolverClass {
    int kind, _pad;                       // CGB_FC_DALLAMICO | CGB_FC_PAINTERKARRA
    double alpha, n, theta_res, Tm, beta, omega;
    double por, sat, org;                 // sat = θwi / por (soil_heat.jl:126)
};

struct PresolverConsts {
    double ch_w, ch_i, ch_a, ch_m, ch_o, L, g, Lsl;
    double Tmin, errtol, Tmax;
    int cap, _pad;
};

// θw(T) and dθw/dT: the closed forms in the operation order of the CPU checker
__device__ inline double ps_sfcc(const PresolverClass& c, const PresolverConsts& K, double T, double* dthdT)
{
    const double thsat = c.por, thtot = c.sat * thsat;
    const double thsat_e = thtot > thsat ? thtot : thsat;
    const double m = 1.0 - 1.0 / c.n;
    double psi0 = 0.0;
    if (thtot < thsat_e) {
        double Se = (thtot - c.theta_res) / (thsat_e - c.theta_res);
        psi0 = -1.0 / c.alpha * pow(pow(Se, -1.0 / m) - 1.0, 1.0 / c.n);
    }
    const double TmK = c.Tm + 273.15;
    double beta = 1.0, omega = 1.0;
    if (c.kind == CGB_FC_PAINTERKARRA) { beta = c.beta; omega = c.omega; }
    const double gT = omega * K.g / K.Lsl * psi0;
    const double Tstar = TmK + gT * TmK;
    const double TK = T + 273.15;
    double psi, dpsi;
    if (Tstar - TK >= 0.0) {
        double cfac = beta * K.Lsl / (K.g * Tstar);
        psi = psi0 + cfac * (TK - Tstar);
        dpsi = cfac;
    } else { psi = psi0; dpsi = 0.0; }
    double th, dth = 0.0;
    if (psi <= 0.0) {
        double x = fabs(-c.alpha * psi);
        double xn = pow(x, c.n);
        double base = 1.0 + xn;
        th = c.theta_res + (thsat_e - c.theta_res) * pow(base, -m);
        double dxn = (x > 0.0) ? c.n * pow(x, c.n - 1.0) * (-c.alpha) : 0.0;
        dth = (thsat_e - c.theta_res) * (-m) * pow(base, -m - 1.0) * dxn;
    } else th = thsat_e;
    *dthdT = dth * dpsi;
    return th;
}

// H(T) = T C(θw) + L θw with sfccheatcap (soil_heat.jl:7-17, incl. the (1-θsat) quirk) and dH/dT (heat_methods.jl:105)
__device__ inline double ps_enthalpy(const PresolverClass& c, const PresolverConsts& K, double T, double* dHdT)
{
    double dth;
    const double thw = ps_sfcc(c, K, T, &dth);
    const double thwi = c.sat * c.por, thsat = c.por;
    const double thi = thwi - thw, tha = thsat - thwi;
    const double thm = (1.0 - c.por) * (1.0 - c.org) * (1.0 - thsat), tho = (1.0 - c.por) * c.org * (1.0 - thsat);
    const double C = K.ch_w * thw + K.ch_i * thi + K.ch_a * tha + K.ch_m * thm + K.ch_o * tho;
    if (dHdT) *dHdT = C + dth * (K.L + T * (K.ch_w - K.ch_i));
    return T * C + K.L * thw;
}

// safeguarded Newton for H(T) = H on [-273, 1e4], converged to 1e-15 relative
__device__ inline double ps_newton(const PresolverClass& c, const PresolverConsts& K, double H, double T0)
{
    double lo = -273.0, hi = 1.0e4;
    double T = T0;
    if (!(T > lo && T < hi)) T = 0.0;
    for (int it = 0; it < 200; ++it) {
        double dHdT;
        double r = ps_enthalpy(c, K, T, &dHdT) - H;
        if (r > 0.0) hi = T; else lo = T;
        double Tn = T - r / dHdT;
        bool small = fabs(Tn - T) <= 1e-15 * fmax(1.0, fabs(Tn));
        bool safe = (Tn > lo && Tn < hi) || (small && Tn >= lo && Tn <= hi);
        if (!safe) Tn = 0.5 * (lo + hi);
        if (fabs(Tn - T) <= 1e-15 * fmax(1.0, fabs(Tn))) { T = Tn; break; }
        T = Tn;
    }
    return T;
}

// One thread per class.  Hk/Tk: [n_classes][cap]; nk[class] = number of knots, or -1 when cap was too small.
__global__ void __launch_bounds__(64) k_presolver_build(const PresolverClass* __restrict__ classes, int n_classes, const __grid_constant__ PresolverConsts K,
                                                        double* __restrict__ Hk, double* __restrict__ Tk, int* __restrict__ nk_out)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n_classes) return;
    const PresolverClass c = classes[id];
    double* hk = Hk + (size_t)id * K.cap;
    double* tk = Tk + (size_t)id * K.cap;
    const double thtot = c.sat * c.por;
    double dHdT;
    double H = ps_enthalpy(c, K, K.Tmin, &dHdT);
    const double Hmax = ps_enthalpy(c, K, K.Tmax, nullptr);
    double T = K.Tmin;
    double dTdH = 1.0 / dHdT;
    int nk = 0;
    hk[nk] = H; tk[nk] = T; ++nk;
    while (H < Hmax) {
        double eps = INFINITY;
        double dH = K.L * (thtot > 1e-8 ? thtot : 1e-8) / 10.0;
        int guard = 0;
        while (fabs(eps) > K.errtol && guard++ < 200) {
            double Test = T + dH * dTdH;
            double Tsol = ps_newton(c, K, H + dH, Test);
            eps = fabs(Tsol - Test);
            dH *= 0.5;
        }
        double Hnew = H + dH;
        double Tnew = ps_newton(c, K, Hnew, T + dH * dTdH);
        ps_enthalpy(c, K, Tnew, &dHdT);
        dTdH = 1.0 / dHdT;
        H = Hnew; T = Tnew;
        if (nk >= K.cap) { nk = -1; break; }
        hk[nk] = H; tk[nk] = T; ++nk;
    }
    nk_out[id] = nk;
}

} // namespace cgb
