// cgb_salt.cuh -- coupled heat + salt diffusion (SalineGround, HeatBalance(:T) + SaltMassBalance), CGEuler.
//
// Reference: Physics/Salt/salt_diffusion.jl:3-30 (freezethaw! with DallAmicoSalt, 2-D duals -> analytic partials
// here), :51-78 (conductivity / diffusivity harmonic means), salt_bc.jl:18-31 + heat_bc.jl:9-35 (boundaries),
// :124-165 (nonlineardiffusion! for dH and dc_F, per-cell 2x2 solve), :107-122 + heat_conduction.jl:162-182
// (salt CFL + MaxDelta(5e-3), heat CFL).  Same execution model as the heat kernel: one warp per column, T and c
// of CPL cells per lane in registers across all steps of a save interval, neighbours by shuffle, dt by warp reduction.
#pragma once
#include "cgb_handle.h"
#include "cgb_euler.cuh"
#include "cgb_euler_newton_tab.cuh"     // nt_fit_bin: degree-7 Chebyshev interpolation of one log-spaced bin

namespace cgb {

// Per-column table of the van Genuchten retention function g(x) = (1 + x^n)^-m over x = alpha |psi| (stepping kernel only): the same
// construction as the per-column curve tables of k_heat_newton_tab -- log-spaced bins from the bits of x (16 per octave,
// 2^ST_EMIN <= x < 2^(ST_EMAX+1)), a degree-7 polynomial in the mantissa remainder per bin, fitted by the warp when it picks up a
// column.  θw = θres + (θsat - θres) g(x) and dθw/dψ = -(θsat - θres) alpha g'(x): porosity stays outside the table (it varies with
// depth).  Replaces two exp/log pairs, two reciprocals and the derivative algebra (~90 FP64 instructions) by 4 x LDS.128 + 13 FMAs.
constexpr int ST_EMIN = -8, ST_EMAX = 15;
constexpr int ST_ROWS = (ST_EMAX + 1 - ST_EMIN) * 16;          // 384
constexpr int ST_STRIDE = 8;                                    // 64-byte rows (8 warps x 24 KB per SM)
constexpr int ST_IDX_BASE = (ST_EMIN + 1023) << 4;

struct VgRetention {
    double n, m;
    __device__ __forceinline__ double operator()(double x) const
    {
        double xn, pb;
        vg_powers_body<true>(x, n, m, xn, pb);
        return pb;
    }
};

// closed form outside the table (|psi| below 4 mm / alpha or above 65 km / alpha, psi > 0): out of line, it hardly ever runs
__device__ __noinline__ double vg_theta_fallback(double psi, double thsat, double thres, double alpha, double n, double m, double* dth)
{
    double d;
    const double th = vg_theta<true>(psi, thsat, thres, alpha, n, m, d);
    *dth = d;
    return th;
}

struct SaltDiag {
    double *thw, *dthwdT, *dthwdc, *C, *dHdT, *H, *kc;   // [N*M]
    double *k, *ds, *jH, *jc;                            // [(N+1)*M]
    double *dH, *dcF, *dT, *dc;                          // [N*M]
};

template <int CPL, bool DIAG>
__global__ void __launch_bounds__(128) k_salt_euler(const __grid_constant__ Dev P, const __grid_constant__ SaltDev S, double t_target, SaltDiag D)
{
    extern __shared__ double smem[];
    double* sInvDk = smem;
    double* sDk = sInvDk + CPL * 32;
    double* sW1 = sDk + CPL * 32;
    double* sW2 = sW1 + CPL * 32;
    double* sG = sW2 + CPL * 32;
    const int lane = threadIdx.x & 31;
    const int N = P.N;
    const long long M = P.M;
    double* __restrict__ tab = sG + CPL * 32 + (size_t)(threadIdx.x >> 5) * (ST_ROWS * ST_STRIDE);     // this warp's retention table (!DIAG)
    if (!DIAG) vgtab_init();                              // exp / log tables of the van Genuchten powers (stepping only)
    for (int s = threadIdx.x; s < CPL * 32; s += blockDim.x) {
        int j = s >> 5, ln = s & 31;
        int i = ln * CPL + j;
        bool cell = i < N, edge = (i > 0 && i < N);
        sInvDk[s] = cell ? P.inv_dk[i] : 0.0;
        sDk[s] = cell ? P.dk[i] : 1.0;
        sW1[s] = edge ? P.eW1[i] : 1.0;
        sW2[s] = edge ? P.eW2[i] : 1.0;
        sG[s] = edge ? P.eG[i] / (P.eW1[i] + P.eW2[i]) : 0.0;      // 1/Δxc of the upper edge (0: not an interior edge)
    }
    __syncthreads();
    const int jb = (N - 1) - lane * CPL;
    const double Lf = P.Lsl * P.rho_w;
    // warp-uniform quotients hoisted out of the step loop; the per-cell divisions below go through rcp_fast (correctly
    // rounded reciprocal, then one multiply: <= 1.5 ulp from the IEEE quotient)
    const double gl = P.g / P.Lsl, Lg = P.Lsl / P.g, inv_tau = 1.0 / S.tau;
    const double dC = P.ch_w - P.ch_i, dK = P.sk_w - P.sk_i;

    for (;;) {
        unsigned long long cu = 0;
        if (lane == 0) cu = atomicAdd(P.counter, 1ull);
        cu = __shfl_sync(FULL, cu, 0);
        if (cu >= (unsigned long long)M) break;
        const long long col = (long long)cu;
        // per-column soil / curve parameters and the per-cell constants derived from them (once per column)
        const double Ssat = S.sat[col], Sorg = S.org[col], Salpha = S.alpha[col], Sn = S.n[col], Sthres = S.theta_res[col];
        const double m = 1.0 - 1.0 / Sn;
        const double TmK0 = S.Tm[col] + 273.15;
        const double dTm_dc = TmK0 / Lf * (-P.R * TmK0);
        const double q0 = TmK0 / Lf;
        double por[CPL], psi0[CPL], thwi[CPL], Cb[CPL], Kb[CPL];
        {
            double thm = (1.0 - S.para_por) * (1.0 - Sorg), tho = (1.0 - S.para_por) * Sorg;   // mineral()/organic() use para.por
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                int i = lane * CPL + j;
                int ic = i < N ? i : N - 1;
                por[j] = S.por_per_column ? S.por_cell[(size_t)ic * M + col] : S.por_cell[ic];
                thwi[j] = por[j] * Ssat;
                double tha = 1.0 - thwi[j] - thm - tho;
                psi0[j] = vg_psi(Ssat * por[j], fmax(Ssat * por[j], por[j]), Sthres, Salpha, Sn);
                Cb[j] = P.ch_i * thwi[j] + P.ch_a * tha + P.ch_m * thm + P.ch_o * tho;
                Kb[j] = P.sk_i * thwi[j] + P.sk_a * tha + P.sk_m * thm + P.sk_o * tho;
            }
        }
        const bool use_tab = !DIAG && Sn != 2.0;                       // n == 2 is a single rsqrt already
        if (use_tab) {
            VgRetention g; g.n = Sn; g.m = m;
            __syncwarp();
            for (int b = lane; b < ST_ROWS; b += 32) nt_fit_bin(g, ST_EMIN + (b >> 4), b & 15, tab + b * ST_STRIDE);
            __syncwarp();
        }
        double T[CPL], c[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = lane * CPL + j;
            T[j] = (i < N) ? P.u[(size_t)i * M + col] : 0.0;
            c[j] = (i < N) ? P.u[(size_t)(N + i) * M + col] : 0.0;
        }
        double t = DIAG ? t_target : P.t[col];
        double dt = P.dt[col];
        const double benthic = S.benthic[col];
        const double toff = P.top_offset[col];
        const double vbot = P.bot_value ? P.bot_value[col] : P.bot.value;
        int status = 0;
        ForcingCursor ctop;
        if (P.top.kind == 2) forcing_seek(P.top, ctop, fmin(fmax(t, __ldg(P.top.t)), __ldg(P.top.t + P.top.n - 1)));
        long long nsteps = 0;

        while (DIAG || t < t_target) {
            // upper boundary temperature at the START of the step (ConstantTemperature, or a TemperatureBC with Input / PeriodicBC)
            const double vtop = forcing_eval(P.top, ctop, t, status) + toff;
            if (!DIAG && (status & 4)) break;
            double thw[CPL], dthT[CPL], dthc[CPL], Cc[CPL], dHdT[CPL], kc[CPL], dsm[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                // DallAmicoSalt: freezing point depression, then Dall'Amico (FreezeCurves.jl; parity unpinned)
                double TmK = TmK0 + q0 * (-P.R * c[j] * TmK0);
                double gT = gl * psi0[j];
                double Tstar = TmK + gT * TmK;
                double dTstar_dc = dTm_dc * (1.0 + gT);
                double TK = T[j] + 273.15;
                double psi, dpsi_dT, dpsi_dc;
                if (Tstar - TK >= 0.0) {
                    double rT = rcp_fast3(Tstar);
                    double cfac = Lg * rT;
                    psi = psi0[j] + cfac * (TK - Tstar);
                    dpsi_dT = cfac;
                    dpsi_dc = Lg * (-TK * (rT * rT)) * dTstar_dc;
                } else { psi = psi0[j]; dpsi_dT = 0.0; dpsi_dc = 0.0; }
                double dth;
                double thsat = sel_max(Ssat * por[j], por[j]);
                if (use_tab) {
                    const double x = Salpha * fabs(psi);
                    const int xh = __double2hiint(x), xl = __double2loint(x);
                    const int bin = (xh >> 16) - ST_IDX_BASE;
                    if (psi <= 0.0 && (unsigned)bin < (unsigned)ST_ROWS) {
                        const double mant = __hiloint2double((xh & 0x000fffff) | 0x3ff00000, xl);
                        const double cen = __hiloint2double(0x3ff08000 | (xh & 0x000f0000), 0);
                        const double r = mant - cen;
                        const double* cf = tab + bin * ST_STRIDE;
                        const double2 c01 = *reinterpret_cast<const double2*>(cf), c23 = *reinterpret_cast<const double2*>(cf + 2);
                        const double2 c45 = *reinterpret_cast<const double2*>(cf + 4), c67 = *reinterpret_cast<const double2*>(cf + 6);
                        double p = fma(c67.y, r, c67.x), q = c67.y;           // Horner with the derivative by synthetic division
                        q = fma(q, r, p); p = fma(p, r, c45.y);
                        q = fma(q, r, p); p = fma(p, r, c45.x);
                        q = fma(q, r, p); p = fma(p, r, c23.y);
                        q = fma(q, r, p); p = fma(p, r, c23.x);
                        q = fma(q, r, p); p = fma(p, r, c01.y);
                        q = fma(q, r, p); p = fma(p, r, c01.x);
                        const double inv_scale = __hiloint2double((2046 - ((xh >> 20) & 0x7ff)) << 20, 0);      // 2^-E
                        const double dsat = thsat - Sthres;
                        thw[j] = fma(dsat, p, Sthres);
                        dth = -(dsat * Salpha) * (q * inv_scale);             // dθ/dψ = -Δθ alpha g'(x)
                    } else if (!(x > 0.0) || psi > 0.0) {                     // psi0 = 0 above T* (saturated): θw = θsat
                        thw[j] = thsat; dth = 0.0;
                    } else {
                        thw[j] = vg_theta_fallback(psi, thsat, Sthres, Salpha, Sn, m, &dth);
                    }
                } else if (DIAG) {
                    thw[j] = vg_theta<false>(psi, thsat, Sthres, Salpha, Sn, m, dth);
                } else if (psi <= 0.0) {
                    // stepping with n == 2 (m = 1/2): θ = θres + Δθ (1+x²)^(-1/2), dθ/dψ = Δθ α x (1+x²)^(-3/2) -- one rsqrt, no table
                    const double x = fabs(-Salpha * psi), dsat = thsat - Sthres;
                    const double rs = rsqrt_fast(fma(x, x, 1.0));
                    dth = dsat * Salpha * x * (rs * rs * rs);
                    thw[j] = fma(dsat, rs, Sthres);
                } else {
                    thw[j] = thsat; dth = 0.0;
                }
                dthT[j] = dth * dpsi_dT;
                dthc[j] = dth * dpsi_dc;
                Cc[j] = fma(dC, thw[j], Cb[j]);
                dHdT[j] = Cc[j] + P.L * dthT[j];                    // salt_diffusion.jl:26
                dsm[j] = (S.ds0 * thw[j]) * inv_tau;                // salt_diffusion.jl:28
                double s = fma(dK, thw[j], Kb[j]);
                kc[j] = s * s;
            }
            double Tup = __shfl_up_sync(FULL, T[CPL - 1], 1);
            double cup = __shfl_up_sync(FULL, c[CPL - 1], 1);
            double kup = __shfl_up_sync(FULL, kc[CPL - 1], 1);
            double dup = __shfl_up_sync(FULL, dsm[CPL - 1], 1);
            double jH[CPL + 1], jc[CPL + 1], ke[CPL], de[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                double T1 = (j == 0) ? Tup : T[j - 1], c1 = (j == 0) ? cup : c[j - 1];
                double k1 = (j == 0) ? kup : kc[j - 1], d1 = (j == 0) ? dup : dsm[j - 1];
                int s = j * 32 + lane;
                double w1 = sW1[s], w2 = sW2[s], idx = sG[s];        // idx = 1/Δxc
                // harmonicmean (math.jl:41): (w1+w2)/(w1/k1 + w2/k2) = (w1+w2) k1 k2 / (w1 k2 + w2 k1), one reciprocal each
                ke[j] = ((w1 + w2) * (k1 * kc[j])) * rcp_fast3(fma(w1, kc[j], w2 * k1));
                de[j] = ((w1 + w2) * (d1 * dsm[j])) * rcp_fast3(fma(w1, dsm[j], w2 * d1));
                jH[j] = -ke[j] * (T[j] - T1) * idx;                  // math.jl:141
                jc[j] = -de[j] * (c[j] - c1) * idx;
                // idx == 0 on non-interior edges: the fluxes vanish there without a select
            }
            if (lane == 0) {
                ke[0] = kc[0]; de[0] = dsm[0];
                jH[0] = -kc[0] * (T[0] - vtop) * P.ctop;                                       // heat_bc.jl:9-17
                jc[0] = (S.surface_state == 0) ? -dsm[0] * (c[0] - benthic) * P.ctop : 0.0;      // salt_bc.jl:18-31
            }
            jH[CPL] = __shfl_down_sync(FULL, jH[0], 1);
            jc[CPL] = __shfl_down_sync(FULL, jc[0], 1);
            double dTv[CPL], dcv[CPL], dHv[CPL], dcF[CPL];
            double rate_s = 0.0, rate_c = 0.0, rate_h = 0.0;
            bool nonfinite = false;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                double jHl = jH[j + 1], jcl = jc[j + 1];
                if (j == jb) { jHl = vbot; jcl = 0.0; jH[j + 1] = jHl; jc[j + 1] = jcl; }
                int s = j * 32 + lane;
                dHv[j] = (jH[j] - jHl) * sInvDk[s];
                dcF[j] = (jc[j] - jcl) * sInvDk[s];
                // salt_diffusion.jl:150-163
                double A = dHdT[j], B = P.L * dthc[j], Dd = dHv[j];
                double E = thw[j] + c[j] * dthc[j], F = c[j] * dthT[j], Gg = dcF[j];
                double det = A * E - B * F;
                double rdet = rcp_fast3(det);
                dTv[j] = (-B * Gg + Dd * E) * rdet;
                dcv[j] = (-F * Dd + A * Gg) * rdet;
                if (lane * CPL + j < N) {
                    // the limiters are minima of quotients with a warp-uniform numerator: carry the MAXIMUM of the per-cell rates and
                    // take one reciprocal per limiter after the warp reduction (was three reciprocals per cell and step)
                    double idx2 = sInvDk[s] * sInvDk[s];
                    double rs = de[j] * idx2;                                                   // salt_diffusion.jl:112-119: dt = courant dx^2 / ds
                    double adc = fabs(dcv[j]);                                                  // dt = dc_max / |dc| (Inf for dc == 0)
                    double rh = (kc[j] * idx2) * rcp_fast3(dHdT[j]);                            // heat_conduction.jl:170-178: dt = courant dHdT dx^2 / k
                    if (rs > rate_s) rate_s = rs;                                               // NaN is skipped, as by fmin
                    if (adc > rate_c) rate_c = adc;
                    if (rh > rate_h) rate_h = rh;
                    nonfinite |= !(rs == rs) || !(adc == adc) || !(rh == rh) || rs < 0.0 || rh < 0.0;
                }
            }
            if (DIAG) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    int i = lane * CPL + j;
                    if (i < N) {
                        size_t o = (size_t)i * M + col;
                        if (D.thw) D.thw[o] = thw[j];
                        if (D.dthwdT) D.dthwdT[o] = dthT[j];
                        if (D.dthwdc) D.dthwdc[o] = dthc[j];
                        if (D.C) D.C[o] = Cc[j];
                        if (D.dHdT) D.dHdT[o] = dHdT[j];
                        if (D.H) D.H[o] = T[j] * Cc[j] + P.L * thw[j];
                        if (D.kc) D.kc[o] = kc[j];
                        if (D.k) D.k[o] = ke[j];
                        if (D.ds) D.ds[o] = de[j];
                        if (D.jH) D.jH[o] = jH[j];
                        if (D.jc) D.jc[o] = jc[j];
                        if (D.dH) D.dH[o] = dHv[j];
                        if (D.dcF) D.dcF[o] = dcF[j];
                        if (D.dT) D.dT[o] = dTv[j];
                        if (D.dc) D.dc[o] = dcv[j];
                        if (i == N - 1) {
                            size_t ob = (size_t)N * M + col;
                            if (D.k) D.k[ob] = kc[j];
                            if (D.ds) D.ds[ob] = dsm[j];
                            if (D.jH) D.jH[ob] = jH[j + 1];
                            if (D.jc) D.jc[ob] = jc[j + 1];
                        }
                    }
                }
                break;
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) { T[j] = fma(dt, dTv[j], T[j]); c[j] = fma(dt, dcv[j], c[j]); }
            t = t + dt;
            ++nsteps;
            // warp minima through the integer REDUX unit (positive doubles order like their bit patterns); a non-positive or
            // NaN limit is the reference's "dt <= 0" assertion (tile.jl:174)
            const bool bad = __any_sync(FULL, nonfinite);
            rate_s = warp_max_nonneg(rate_s); rate_c = warp_max_nonneg(rate_c); rate_h = warp_max_nonneg(rate_h);
            auto quot = [](double num, double rate) { return rate > 0.0 ? num * rcp_fast(rate) : INFINITY; };   // rate 0: no limit
            double lim = sel_min(sel_min(quot(S.courant, rate_s), quot(S.dc_max, rate_c)), quot(S.heat_courant, rate_h));
            if (bad || !(lim > 0.0)) { status |= 8; lim = 0.0; }
            double dtn = sel_max(sel_min(lim, P.dtmax), P.dtmin);
            if (t < t_target) dtn = sel_min(dtn, t_target - t);
            dt = sel_min(dtn, P.dtmax);
        }
        if (!DIAG) {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                int i = lane * CPL + j;
                if (i < N) { P.u[(size_t)i * M + col] = T[j]; P.u[(size_t)(N + i) * M + col] = c[j]; }
            }
            if (lane == 0) {
                P.t[col] = t; P.dt[col] = dt;
                atomicAdd(reinterpret_cast<unsigned long long*>(P.steps + col), (unsigned long long)nsteps);
                if (status) atomicOr(P.status + col, status);
            }
        }
    }
}

} // namespace cgb

// ---- host side -------------------------------------------------------------------------------------
static int salt_finalize(cgb_handle* h, const std::vector<double>&)
{
    using namespace cgb;
    if (h->cfg.physics != CGB_PHYSICS_HEAT_SALT) return CGB_OK;
    if (h->cfg.stepper != CGB_STEPPER_EULER) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "heat+salt supports CGEuler only (as the reference)");
    SaltDev& s = h->salt;
    ParamArray& por = h->params["por"];
    const size_t M = (size_t)h->M;
    s.por_per_column = (por.per == CGB_PER_CELL_COLUMN);
    std::vector<double> pc(s.por_per_column ? (size_t)h->N * M : (size_t)h->N);
    if (s.por_per_column) pc = por.v;
    else for (int i = 0; i < h->N; ++i) pc[i] = (por.per == CGB_PER_CELL) ? por.v[i] : por.v[0];
    double* p_por; int rc = dev_upload(h, &p_por, pc); if (rc) return rc;
    s.por_cell = p_por;
    auto per_column = [&](const char* name, const double** dst) {
        ParamArray& a = h->params[name];
        std::vector<double> v(M);
        for (size_t c = 0; c < M; ++c) v[c] = (a.per == CGB_PER_COLUMN) ? a.v[c] : a.v[0];
        double* p; int r = dev_upload(h, &p, v); *dst = p; return r;
    };
    rc |= per_column("benthic_salt", &s.benthic);
    rc |= per_column("sat", &s.sat); rc |= per_column("org", &s.org); rc |= per_column("vg_alpha", &s.alpha); rc |= per_column("vg_n", &s.n);
    rc |= per_column("theta_res", &s.theta_res); rc |= per_column("Tm", &s.Tm);
    if (rc) return rc;
    auto G = [&](const char* n) { return h->params[n].v[0]; };
    s.para_por = G("salt_para_por");
    s.tau = G("tau"); s.ds0 = G("ds0"); s.dc_max = G("salt_dc_max");
    s.courant = h->cfg.courant; s.heat_courant = 0.5;                 // salt_types.jl:13 ; heat_types.jl:17
    s.surface_state = (int)G("salt_surface_state");
    if (h->forc[1].kind != 0 && !h->params.count("bot_value"))
        CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "heat+salt supports a constant lower boundary value (GeothermalHeatFlux)");
    if (h->cfg.top_kind != CGB_BC_DIRICHLET || h->cfg.bot_kind != CGB_BC_NEUMANN)
        CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "heat+salt supports a Dirichlet top and Neumann bottom (examples/heat_sfcc_salt_constantbc.jl)");
    return CGB_OK;
}

template <int CPL, bool DIAG>
static int salt_launch_t(cgb_handle* h, double t_target, const cgb::SaltDiag& D)
{
    using namespace cgb;
    auto kern = k_salt_euler<CPL, DIAG>;
    size_t smem = (size_t)(5 * CPL * 32 + (DIAG ? 0 : 4 * ST_ROWS * ST_STRIDE)) * sizeof(double);
    int blocks_per_sm = 0;
    CUDA_TRY(h, cgb_prepare_kernel(h, (const void*)kern, smem, &blocks_per_sm));
    long long grid = std::min<long long>((h->M + 3) / 4, (long long)h->num_sms * blocks_per_sm);
    CUDA_TRY(h, cudaMemsetAsync(h->dev.counter, 0, sizeof(unsigned long long), h->stream));
    kern<<<(unsigned)grid, 128, smem, h->stream>>>(h->dev, h->salt, t_target, D);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return CGB_OK;
}

template <bool DIAG>
static int salt_launch_c(cgb_handle* h, double t_target, const cgb::SaltDiag& D)
{
    int cpl = (h->N + 31) / 32;
    if (cpl <= 2) return salt_launch_t<2, DIAG>(h, t_target, D);
    if (cpl <= 4) return salt_launch_t<4, DIAG>(h, t_target, D);
    if (cpl <= 6) return salt_launch_t<6, DIAG>(h, t_target, D);
    if (cpl <= 7) return salt_launch_t<7, DIAG>(h, t_target, D);
    if (cpl <= 9) return salt_launch_t<9, DIAG>(h, t_target, D);
    CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "heat+salt kernels are built for n_cells <= 288");
}

static int salt_launch(cgb_handle* h, double t_target, bool diag, const cgb::Diag&, void* sd)
{
    cgb::SaltDiag D{};
    if (diag) return salt_launch_c<true>(h, t_target, *static_cast<cgb::SaltDiag*>(sd));
    return salt_launch_c<false>(h, t_target, D);
}

// one diagnostic variable evaluated at the current state and time t, written to a DEVICE buffer (dev_out: N*M or (N+1)*M doubles)
static int salt_diag_to(cgb_handle* h, const char* var, double t, double* dev_out, size_t* count)
{
    using namespace cgb;
    SaltDiag D{};
    double* s = dev_out;
    size_t NM = (size_t)h->N * h->M, cnt = NM;
    std::string v(var);
    if (v == "thw") D.thw = s; else if (v == "dthwdT") D.dthwdT = s; else if (v == "dthwdc") D.dthwdc = s; else if (v == "C") D.C = s;
    else if (v == "dHdT") D.dHdT = s; else if (v == "H") D.H = s; else if (v == "kc") D.kc = s; else if (v == "dH") D.dH = s;
    else if (v == "dc_F") D.dcF = s; else if (v == "dT") D.dT = s; else if (v == "dc") D.dc = s;
    else if (v == "k") { D.k = s; cnt = NM + h->M; } else if (v == "ds") { D.ds = s; cnt = NM + h->M; }
    else if (v == "jH") { D.jH = s; cnt = NM + h->M; } else if (v == "jc") { D.jc = s; cnt = NM + h->M; }
    else CGB_FAIL(h, CGB_ERR_INVALID, "unknown heat+salt diagnostic '%s'", var);
    if (count) *count = cnt;
    Diag dummy{};
    return salt_launch(h, t, true, dummy, &D);
}

static int salt_get_diag(cgb_handle* h, const char* var, double t, double* out)
{
    size_t cnt = 0;
    int rc = salt_diag_to(h, var, t, h->d_scratch, &cnt);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_scratch, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return CGB_OK;
}
