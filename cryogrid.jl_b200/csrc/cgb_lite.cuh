// cgb_lite.cuh -- LiteImplicitEuler (CryoGridLite) step: Picard iteration with a batched Thomas solve.
//
// Variant "thomas-per-thread": one THREAD owns one column and marches down/up the cells, so with the
// column-fastest layout every load/store of a warp is one fully coalesced 256-byte row segment.  The forward
// sweep fuses freezethaw (heat_conduction.jl:18-57 / soil_heat.jl:116-140), conductivities
// (thermal_properties.jl:53-66), the implicit coefficients (heat_implicit.jl:10-24, 72-103), the linear-system
// assembly (cglite_step.jl:81-99) and the Thomas elimination (Numerics/math.jl:169-197) in the reference's own
// operation order; the backward sweep fuses back-substitution with the Picard update and the residual maximum
// (cglite_step.jl:58-72).  c', d', T and dH/dT travel through a global scratch (coalesced).
#pragma once
#include "cgb_handle.h"
#include "cgb_euler.cuh"

namespace cgb {

struct LiteScratch {
    double *cp, *dp, *T, *dHdT, *H0;   // [N*M] each
};

struct LiteCell { double T, dHdT, kc, thw, C, dth; };

__device__ __forceinline__ void lite_cell(const Dev& P, const double* lc, double H, double Tguess, LiteCell& c)
{
    // reference operation order (true divisions): this kernel is HBM-bound, FP64 latency is hidden
    int kind = reinterpret_cast<const int*>(lc + LC_KIND)[0];
    const double dC = P.ch_w - P.ch_i, dK = P.sk_w - P.sk_i;
    if ((kind & 0xff) == 0) {
        double thtot = lc[LC_THTOT];
        double Lth = P.L * thtot;
        double thw = ((H > 0.0 && H <= Lth) ? (H / Lth) : 0.0) + ((H > Lth) ? 1.0 : 0.0);
        thw *= thtot;
        double C = fma(dC, thw, lc[LC_CB]);
        double Lw = lc[LC_LTH];
        double T = (H < 0.0) ? H / C : ((H >= Lw) ? (H - Lw) / C : 0.0);
        c.T = T; c.thw = thw; c.C = C; c.dth = 0.0;
        c.dHdT = (T == 0.0) ? 1e8 : C;
    } else {
        int solver = (kind >> 8) & 0xff;
        double T = Tguess, thw, C, dHdT, dth;
        if (solver == 0) {
            T = lut_T(P, lc, H);
            thw = sfcc_theta(lc, T, dth);
            C = fma(dC, thw, lc[LC_CBS]);
            dHdT = C + dth * (P.L + T * dC);
        } else {
            sfcc_newton_solve<true>(lc, P.L, dC, H, T, thw, C, dHdT, dth);
        }
        c.T = T; c.thw = thw; c.C = C; c.dth = dth; c.dHdT = dHdT;
    }
    double s = fma(dK, c.thw, lc[LC_KB]);
    c.kc = s * s;
}

__device__ __forceinline__ double harmonic(double x1, double x2, double w1, double w2)
{
    return (w1 + w2) / (w1 * (1.0 / x1) + w2 * (1.0 / x2));   // math.jl:141
}

__device__ __forceinline__ double lite_forcing(const Forcing& F, double t, int& status)
{
    if (F.kind == 0) return F.value;
    if (F.kind == 1) return F.amplitude * sin(2.0 * M_PI * (t / F.period) + F.phase) + F.level;
    if (t < __ldg(F.t) || t > __ldg(F.t + F.n - 1)) { status |= 4; return nan(""); }
    return lerp_table(F.t, F.y, F.n, t);
}

template <bool DIAG>
__global__ void __launch_bounds__(128) k_lite_thomas(const __grid_constant__ Dev P, double t_target, LiteScratch S, Diag D)
{
    const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (col >= P.M) return;
    const int N = P.N;
    const long long M = P.M;
    double* u = P.u;
    double t = DIAG ? t_target : P.t[col];
    double dt = P.dt[col];
    const double nf = P.nf[col], nt = P.nt[col], toff = P.top_offset[col];
    int status = 0;
    long long nsteps = 0, niter = 0;
    double lc[LC_STRIDE];
    int cur_layer = -1;

    while (DIAG || t < t_target) {
        const double tn = DIAG ? t : t + dt;                                   // cglite_step.jl:9 (forcing at the NEW time)
        double vtop = lite_forcing(P.top, tn, status) + toff;
        if (P.top_kind == 0 && P.use_nfactor) vtop = ((vtop <= 0.0) ? nf : nt) * vtop;
        double vbot = P.bot_value ? P.bot_value[col] : lite_forcing(P.bot, tn, status);
        if (!DIAG)
            for (int i = 0; i < N; ++i) S.H0[(size_t)i * M + col] = u[(size_t)i * M + col];   // cache.uprev
        int iter = 1;
        double emax = INFINITY;
        double k_first_bot = 0.0;
        while (DIAG || (emax > P.lite_tol && iter <= P.lite_maxiters) || iter < P.lite_miniters) {   // cglite_step.jl:43
            // ---------------- forward sweep
            LiteCell prev{}, cur{}, nxt{};
            double Hcur = u[col], Hnxt = 0.0;
            int l0 = P.cell_layer[0];
            if (l0 != cur_layer) { layer_consts(P, col, l0, lc); cur_layer = l0; }
            lite_cell(P, lc, Hcur, Hcur / P.ch_w, cur);
            double cp_prev = 0.0, dp_prev = 0.0;
            double k_up = cur.kc;                                               // k[1] = kc[1]
            for (int i = 0; i < N; ++i) {
                size_t o = (size_t)i * M + col;
                double k_dn;
                if (i + 1 < N) {
                    Hnxt = u[o + M];
                    int l = P.cell_layer[i + 1];
                    if (l != cur_layer) { layer_consts(P, col, l, lc); cur_layer = l; }
                    lite_cell(P, lc, Hnxt, cur.T, nxt);
                    k_dn = harmonic(cur.kc, nxt.kc, P.dk[i], P.dk[i + 1]);      // thermal_properties.jl:56-60
                } else {
                    k_dn = cur.kc;                                              // k[end] = kc[end]
                }
                if (i == P.bot_layer_first) k_first_bot = k_up;
                double dki = P.dk[i];
                double an = (i > 0) ? k_up / P.dxc[i - 1] / dki : 0.0;          // heat_implicit.jl:10-24
                double as = (i + 1 < N) ? k_dn / P.dxc[i] / dki : 0.0;
                double ap = an + as;
                double bp = 0.0;
                if (i == 0) {                                                   // heat_implicit.jl:59-64,72-79
                    double jtop = (P.top_kind == 0) ? 2.0 * vtop * k_up / dki : vtop;
                    bp += jtop / dki;
                    ap += (P.top_kind == 0) ? k_up / (dki * dki / 2.0) : 0.0;
                }
                if (i == N - 1) {                                               // heat_implicit.jl:65-70,80-86 (quirks verbatim)
                    double jbot = (P.bot_kind == 0) ? 2.0 * vbot * k_dn / dki : vbot;
                    bp += jbot / dki;
                    ap += (P.bot_kind == 0) ? k_first_bot / (dki * dki / 2.0) : 0.0;
                }
                if (DIAG) {
                    if (D.T) D.T[o] = cur.T;
                    if (D.thw) D.thw[o] = cur.thw;
                    if (D.C) D.C[o] = cur.C;
                    if (D.dHdT) D.dHdT[o] = cur.dHdT;
                    if (D.dthwdT) D.dthwdT[o] = cur.dth;
                    if (D.kc) D.kc[o] = cur.kc;
                    if (D.k) { D.k[o] = k_up; if (i == N - 1) D.k[o + M] = k_dn; }
                    if (D.an) D.an[o] = an;
                    if (D.as) D.as[o] = as;
                    if (D.ap) D.ap[o] = ap;
                    if (D.bp) D.bp[o] = bp;
                    if (D.dH) D.dH[o] = 0.0;
                    if (D.jH) { D.jH[o] = 0.0; if (i == N - 1) D.jH[o + M] = 0.0; }
                } else {
                    // cglite_linsolve!: cglite_step.jl:81-99
                    double Sp = -cur.dHdT / dt;
                    double Sc = (S.H0[o] - Hcur) / dt - Sp * cur.T;
                    double A = -an, B = ap - Sp, Cc = -as, Dd = Sc + bp;
                    double cpi, dpi;
                    if (i == 0) { cpi = Cc / B; dpi = Dd / B; }                 // math.jl:176-190
                    else { double temp = B - A * cp_prev; cpi = Cc / temp; dpi = (Dd - A * dp_prev) / temp; }
                    S.cp[o] = cpi; S.dp[o] = dpi; S.T[o] = cur.T; S.dHdT[o] = cur.dHdT;
                    if (D.thw) D.thw[o] = cur.thw;          // saved θw = cache of the last tile evaluation
                    cp_prev = cpi; dp_prev = dpi;
                }
                prev = cur; cur = nxt; Hcur = Hnxt; k_up = k_dn;
            }
            if (DIAG) break;
            // ---------------- backward sweep: back-substitution + Picard update (cglite_step.jl:58-72)
            double x = 0.0;
            emax = -INFINITY;
            for (int i = N - 1; i >= 0; --i) {
                size_t o = (size_t)i * M + col;
                x = (i == N - 1) ? S.dp[o] : S.dp[o] - S.cp[o] * x;
                double e = x - S.T[o];
                u[o] = fma(S.dHdT[o], e, u[o]);
                emax = fmax(emax, fabs(e));
                if (!isfinite(e)) status |= 2;
            }
            ++niter; ++iter;
            if (status & 2) break;
        }
        if (DIAG) break;
        if (iter > P.lite_maxiters && emax > P.lite_tol) status |= 1;           // cglite_step.jl:74-77
        t = tn; ++nsteps;
        if (t < t_target) dt = fmin(dt, t_target - t);                          // integrator.jl:126-131
        dt = fmin(P.dtmax, dt);
        if (status & (2 | 4)) break;
    }
    if (!DIAG) {
        if (D.T && nsteps > 0)
            for (int i = 0; i < N; ++i) D.T[(size_t)i * M + col] = S.T[(size_t)i * M + col];   // saved T: last tile evaluation
        P.t[col] = t; P.dt[col] = dt; P.steps[col] += nsteps; P.iters[col] += niter;
        if (status) P.status[col] |= status;
    }
}

} // namespace cgb

static int lite_launch(cgb_handle* h, double t_target, bool diag, const cgb::Diag& D)
{
    using namespace cgb;
    size_t NM = (size_t)h->N * h->M;
    if (!h->d_lite) {
        cudaError_t e = cudaMalloc((void**)&h->d_lite, 5 * NM * sizeof(double));
        if (e != cudaSuccess) CGB_FAIL(h, CGB_ERR_ALLOC, "LiteImplicit scratch (%zu bytes): %s", 5 * NM * sizeof(double), cudaGetErrorString(e));
    }
    LiteScratch S{h->d_lite, h->d_lite + NM, h->d_lite + 2 * NM, h->d_lite + 3 * NM, h->d_lite + 4 * NM};
    unsigned grid = (unsigned)((h->M + 127) / 128);
    if (diag) k_lite_thomas<true><<<grid, 128, 0, h->stream>>>(h->dev, t_target, S, D);
    else k_lite_thomas<false><<<grid, 128, 0, h->stream>>>(h->dev, t_target, S, D);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return CGB_OK;
}
