// cgb_lite_warp.cuh -- LiteImplicitEuler, variant "warp per column": register-resident Picard iteration with a
// partitioned tridiagonal solve across the 32 lanes of a warp.
//
// Why: with one thread per column (cgb_lite.cuh) a launch lasts as long as its slowest column and a warp iterates until
// its slowest lane converges.  Here one WARP owns one column (lane l holds rows [l*CPL, (l+1)*CPL)), H, H0, T, dH/dT and
// the system rows stay in registers for all Picard iterations and all time steps of a save interval, columns are
// fetched from a global cursor, and the iteration count is warp-uniform by construction.
//
// Tridiagonal solve (replaces the sequential Thomas sweep of Numerics/math.jl:169-197; same linear system, rows
// assembled exactly as cglite_step.jl:81-99): Wang-type partition --
//   1. local forward elimination   (row j couples to x_p = last unknown of the previous lane, x_j, x_{j+1})
//   2. local backward elimination  (row j couples to x_p, x_j and the lane's last unknown x_{m-1})
//   3. reduced 32-unknown system in the lanes' last unknowns, solved by parallel cyclic reduction with shuffles
//   4. local back-substitution.
// The matrices are strictly diagonally dominant (B = ap + dH/dT/dt > an + as), so no pivoting is needed; results differ
// from the Thomas order by rounding only (checked against the oracle in tests/test_gpu_lite.py).
#pragma once
#include "cgb_handle.h"
#include "cgb_euler.cuh"

namespace cgb {

// SAVEW: keep θw of the last evaluation for the saved output (9-16 more doubles per lane; a separate instantiation so that the
// plain stepping kernel does not carry them)
template <int CPL, int MODE, int MINB, bool SAVEW>
__global__ void __launch_bounds__(128, MINB) k_lite_warp(const __grid_constant__ Dev P, const __grid_constant__ FastTargets TG,
                                                         double* __restrict__ saveT, double* __restrict__ saveW)
{
    extern __shared__ double smem[];
    double* sDk = smem;                    // [CPL*32] cell thickness (1 for padding)
    double* sIdk = sDk + CPL * 32;         // 1/dk (0 for padding)
    double* sW1 = sIdk + CPL * 32;         // harmonic-mean weights of the upper edge
    double* sW2 = sW1 + CPL * 32;
    double* sG = sW2 + CPL * 32;           // (w1+w2)/dxc of the upper edge (0 for non-interior edges)
    double* sLCall = sG + CPL * 32;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = P.N;
    const long long M = P.M;
    double* lcw = sLCall + warp * (P.n_layers * LC_STRIDE);
    for (int s = threadIdx.x; s < CPL * 32; s += blockDim.x) {
        int j = s >> 5, ln = s & 31;
        int i = ln * CPL + j;
        bool cell = i < N, edge = (i > 0 && i < N);
        sDk[s] = cell ? P.dk[i] : 1.0;
        sIdk[s] = cell ? P.inv_dk[i] : 0.0;
        sW1[s] = edge ? P.eW1[i] : 1.0;
        sW2[s] = edge ? P.eW2[i] : 1.0;
        sG[s] = edge ? P.eG[i] : 0.0;
    }
    __syncthreads();
    const double dC = P.ch_w - P.ch_i, dK = P.sk_w - P.sk_i, inv_chw = 1.0 / P.ch_w;
    // seek hints in shared memory: the cursors of the warp's previous column at its start (ta, tb, ya, yb, idx) x (top, bottom)
    double* hint = sLCall + 4 * (P.n_layers * LC_STRIDE) + warp * 16;
    if (lane == 0) { hint[4] = -1.0; hint[12] = -1.0; }
    __syncwarp();
    int loff[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        int i = lane * CPL + j;
        loff[j] = P.cell_layer[i < N ? i : N - 1] * LC_STRIDE;
    }
    const int jb = (N - 1) - lane * CPL;                 // local slot of the bottom row
    const int ef = P.bot_layer_first;                    // heat_implicit.jl:83 quirk: k[first edge of the bottom layer]

    for (;;) {
        unsigned long long cu = 0;
        if (lane == 0) cu = atomicAdd(P.counter, 1ull);
        cu = __shfl_sync(FULL, cu, 0);
        if (cu >= (unsigned long long)M) break;
        const long long col = (long long)cu;
        __syncwarp();
        stage_layer_consts(P, col, lcw, lane);
        __syncwarp();

        double H[CPL], H0[CPL], T[CPL], W[SAVEW ? CPL : 1];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = lane * CPL + j;
            H[j] = (i < N) ? P.u[(size_t)i * M + col] : 0.0;
            T[j] = H[j] * inv_chw;                  // Newton seed (soil_heat.jl:127); FreeWater overwrites it
            if (SAVEW) W[j] = 0.0;
        }
        double t = P.t[col], dt = P.dt[col];
        const double nf = P.nf[col], nt = P.nt[col], toff = P.top_offset[col];
        int status = 0;
        long long nsteps = 0, niter = 0;
        ForcingCursor ctop, cbot;
        // the columns of an ensemble usually sit at the same time: reuse the previous column's starting cursor when it
        // brackets t (no global loads), else seek
        auto start_cursor = [&](const Forcing& F, ForcingCursor& c, double* hn) {
            if (hn[4] >= 0.0 && hn[0] <= t && t < hn[1]) {
                c.ta = hn[0]; c.tb = hn[1]; c.ya = hn[2]; c.yb = hn[3]; c.idx = (int)hn[4];
            } else {
                forcing_seek(F, c, fmin(fmax(t, __ldg(F.t)), __ldg(F.t + F.n - 1)));
                __syncwarp();
                if (lane == 0) { hn[0] = c.ta; hn[1] = c.tb; hn[2] = c.ya; hn[3] = c.yb; hn[4] = (double)c.idx; }
                __syncwarp();
            }
        };
        if (P.top.kind == 2) start_cursor(P.top, ctop, hint);
        if (P.bot.kind == 2 && !P.bot_value) start_cursor(P.bot, cbot, hint + 8);

        // device-side saveat: the column is carried through all save times of the launch in registers
        for (int kt = 0; kt < TG.n && !(status & 6); ++kt) {
        const double t_target = TG.t[kt];
        bool stepped = false;
        while (t < t_target) {
            const double tn = t + dt;                                               // cglite_step.jl:9
            double vtop = forcing_eval(P.top, ctop, tn, status) + toff;
            if (P.top_kind == 0 && P.use_nfactor) vtop = ((vtop <= 0.0) ? nf : nt) * vtop;
            const double vbot = P.bot_value ? P.bot_value[col] : forcing_eval(P.bot, cbot, tn, status);
            if (status & 4) break;
            const double idt = 1.0 / dt;
#pragma unroll
            for (int j = 0; j < CPL; ++j) H0[j] = H[j];                              // cache.uprev
            int iter = 1;
            double emax = INFINITY;
            while ((emax > P.lite_tol && iter <= P.lite_maxiters) || iter < P.lite_miniters) {   // cglite_step.jl:43
                // ---- diagnostics
                double kc[CPL], dHdT[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    double thw, C, dth;
                    freezethaw_cell<MODE>(P, lcw + loff[j], H[j], dC, dK, T[j], thw, C, dHdT[j], dth, kc[j]);
                    if (SAVEW) W[j] = thw;
                }
                // ---- kd = k_edge/dxc on the upper edge of every local cell (harmonic mean, math.jl:141)
                double kup = __shfl_up_sync(FULL, kc[CPL - 1], 1);
                double kd[CPL + 1];
                double kfirst = 0.0;   // k on edge `ef` (Dirichlet bottom quirk only)
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    double k1 = (j == 0) ? kup : kc[j - 1];
                    int s = j * 32 + lane;
                    double w1 = sW1[s], w2 = sW2[s], g = sG[s];
                    double r = rcp_fast3(fma(w2, k1, w1 * kc[j]));
                    double kk = k1 * kc[j];
                    kd[j] = (kk * g) * r;                                            // 0 on non-interior edges
                    if (P.bot_kind == 0 && lane * CPL + j == ef) kfirst = (ef == 0) ? kc[j] : (kk * (w1 + w2)) * r;
                }
                kd[CPL] = __shfl_down_sync(FULL, kd[0], 1);
                if (lane == 31) kd[CPL] = 0.0;
                if (P.bot_kind == 0) kfirst = __shfl_sync(FULL, kfirst, ef / CPL);
                // ---- rows: A = -an, B = ap + dHdT/dt, C = -as, D = (H0-H)/dt + dHdT/dt*T + bp   (cglite_step.jl:81-99)
                double fa[CPL], gc[CPL], rb[CPL], d[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    int s = j * 32 + lane;
                    double idk = sIdk[s];
                    // padding rows (idk = 0) and the edge below the bottom cell (kd = 0: sG is 0 on non-interior edges) need no
                    // select: an = as = 0 there, the row reduces to (dHdT/dt) x = (dHdT/dt) T
                    double an = kd[j] * idk, as = kd[j + 1] * idk;                   // heat_implicit.jl:10-24
                    double sp = dHdT[j] * idt;                                       // -Sp
                    double bj = (an + as) + sp;                                      // B = ap + dHdT/dt (reciprocal taken below)
                    double dj = fma(sp, T[j], (H0[j] - H[j]) * idt);                 // D without bp: bp is zero except in the two boundary rows
                    if (j == 0 && lane == 0) {                                       // heat_implicit.jl:59-64,72-79
                        double jtop = (P.top_kind == 0) ? 2.0 * vtop * kc[0] * idk : vtop;
                        dj = fma(jtop, idk, dj);
                        if (P.top_kind == 0) bj = fma(kc[0], 2.0 * idk * idk, bj);     // k / (dk^2 / 2) without the division
                    }
                    if (j == jb) {                                                   // heat_implicit.jl:65-70,80-86 (quirks verbatim)
                        double jbot = (P.bot_kind == 0) ? 2.0 * vbot * kc[j] * idk : vbot;
                        dj = fma(jbot, idk, dj);
                        if (P.bot_kind == 0) bj = fma(kfirst, 2.0 * idk * idk, bj);
                    }
                    fa[j] = -an;
                    gc[j] = -as;
                    rb[j] = bj;
                    d[j] = dj;
                }
                // ---- 1. forward elimination: row j -> f_j x_p + den_j x_j + c_j x_{j+1} = d_j, DIVISION-FREE in the dependent chain.
                // The textbook sweep (w = a_j / den_{j-1}; den_j = b_j - w c_{j-1}; ...) carries one reciprocal (MUFU + 4 dependent
                // FMAs) per row: ~65 cycles x CPL with ILP 1 -- half of the latency of one Picard iteration.  With the continuant
                // q_j = b_j q_{j-1} - a_j c_{j-1} q_{j-2} (q_{-1} = 1, q_0 = b_0) one has den_j = q_j / q_{j-1}, and the scaled
                // unknowns F_j = f_j q_{j-1}, E_j = d_j' q_{j-1} obey F_j = -a_j F_{j-1}, E_j = d_j q_{j-1} - a_j E_{j-1}: one FMA
                // per row in each chain, then CPL independent reciprocals.  Same algebra, rounding differs in the last bits; the
                // systems are diagonally dominant (q_j / q_{j-1} >= b_j - |a_j|), q_j <= (max den)^CPL stays far below 1e308.
                // In place: rb[j] <- q_j, d[j] <- E_j, fa[j] <- F_j.
#pragma unroll
                for (int j = 1; j < CPL; ++j) {
                    const double ac = fa[j] * gc[j - 1];
                    const double qj = (j >= 2) ? fma(rb[j], rb[j - 1], -ac * rb[j - 2]) : fma(rb[j], rb[j - 1], -ac);
                    d[j] = fma(-fa[j], d[j - 1], d[j] * rb[j - 1]);
                    fa[j] = -fa[j] * fa[j - 1];
                    rb[j] = qj;
                }
                double rq[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) rq[j] = rcp_fast3(rb[j]);
                const double den_last = (CPL >= 2) ? rb[CPL - 1] * rq[CPL >= 2 ? CPL - 2 : 0] : rb[0];    // q_last / q_{last-1}
#pragma unroll
                for (int j = CPL - 1; j >= 1; --j) {
                    fa[j] *= rq[j - 1];
                    d[j] *= rq[j - 1];
                    rb[j] = rb[j - 1] * rq[j];                                       // 1 / den_j = q_{j-1} / q_j
                }
                rb[0] = rq[0];
                // ---- 2. backward elimination: row j -> f_j x_p + b_j x_j + g_j x_{m-1} = d_j   (j <= m-2)
                const double c_last = gc[CPL - 1];
#pragma unroll
                for (int j = CPL - 3; j >= 0; --j) {
                    double w = gc[j] * rb[j + 1];
                    fa[j] = fma(-w, fa[j + 1], fa[j]);
                    d[j] = fma(-w, d[j + 1], d[j]);
                    gc[j] = -w * gc[j + 1];
                }
                // ---- 3. reduced system in y_l = x_{m-1} of lane l; x_0 of lane l+1 = (d0 - f0 y_l - g0 y_{l+1}) / b0
                double rf, rg, rd;
                if (CPL >= 2) { rf = fa[0] * rb[0]; rg = gc[0] * rb[0]; rd = d[0] * rb[0]; }
                else { rf = 0.0; rg = 0.0; rd = 0.0; }
                double nf0 = __shfl_down_sync(FULL, rf, 1), ng0 = __shfl_down_sync(FULL, rg, 1), nd0 = __shfl_down_sync(FULL, rd, 1);
                double A, B, C, D;
                if (CPL >= 2) {
                    A = fa[CPL - 1];
                    B = fma(-c_last, nf0, den_last);
                    C = -c_last * ng0;
                    D = fma(-c_last, nd0, d[CPL - 1]);
                } else {   // one row per lane: the reduced system is the system itself
                    A = fa[0]; B = den_last; C = c_last; D = d[0];
                }
                if (lane == 31) C = 0.0;
                if (lane == 0) A = 0.0;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {                                   // parallel cyclic reduction
                    double rB = rcp_fast3(B);
                    double Au = __shfl_up_sync(FULL, A, s), Cu = __shfl_up_sync(FULL, C, s), Du = __shfl_up_sync(FULL, D, s), rBu = __shfl_up_sync(FULL, rB, s);
                    double Ad = __shfl_down_sync(FULL, A, s), Cd = __shfl_down_sync(FULL, C, s), Dd = __shfl_down_sync(FULL, D, s), rBd = __shfl_down_sync(FULL, rB, s);
                    bool hu = lane >= s, hd = lane + s < 32;
                    // a lane without a neighbour at distance s gets its own (finite) values back from the shuffle: a zero multiplier is enough
                    double al = hu ? -A * rBu : 0.0, ga = hd ? -C * rBd : 0.0;
                    B = B + al * Cu + ga * Ad;
                    D = D + al * Du + ga * Dd;
                    A = al * Au;
                    C = ga * Cd;
                }
                const double y = D * rcp_fast3(B);
                double xp = __shfl_up_sync(FULL, y, 1);
                if (lane == 0) xp = 0.0;
                // ---- 4. back-substitution + Picard update (cglite_step.jl:58-72)
                double em = 0.0;
                bool bad = false;
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    double x;
                    if (CPL >= 2) x = (j == CPL - 1) ? y : (d[j] - fa[j] * xp - gc[j] * y) * rb[j];
                    else x = y;
                    {
                        double e = x - T[j];
                        H[j] = fma(dHdT[j], e, H[j]);
                        em = sel_max(fabs(e), em);
                        bad |= !isfinite(e);
                    }
                }
                emax = warp_max_nonneg(em);                                          // em >= 0 (NaN never enters: see `bad`)
                ++niter; ++iter;
                if (__any_sync(FULL, bad)) { status |= 2; break; }
            }
            if (iter > P.lite_maxiters && emax > P.lite_tol) status |= 1;           // cglite_step.jl:74-77
            t = tn; ++nsteps; stepped = true;
            if (t < t_target) dt = fmin(dt, t_target - t);                          // integrator.jl:126-131
            dt = fmin(P.dtmax, dt);
            if (status & 2) break;
        }
        if ((saveT || saveW) && stepped) {
            // saved diagnostics = cache of the LAST tile evaluation (start of the last Picard iteration), as the
            // reference saves them (problem.jl:67-68, Numerics/statevars.jl:81-106, cglite_step.jl:45-62)
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                int i = lane * CPL + j;
                if (i < N) {
                    size_t o = (size_t)kt * TG.save_stride + (size_t)i * M + col;
                    if (saveT) saveT[o] = T[j];
                    if (SAVEW) saveW[o] = W[j];
                }
            }
        }
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = lane * CPL + j;
            if (i < N) P.u[(size_t)i * M + col] = H[j];
        }
        if (lane == 0) {
            P.t[col] = t; P.dt[col] = dt;
            // counters through RED (no load -> no stall at the end of the column)
            atomicAdd(reinterpret_cast<unsigned long long*>(P.steps + col), (unsigned long long)nsteps);
            atomicAdd(reinterpret_cast<unsigned long long*>(P.iters + col), (unsigned long long)niter);
            if (status) atomicOr(P.status + col, status);
        }
    }
}

} // namespace cgb

template <int CPL, int MODE, int MINB, bool SAVEW>
static int lite_warp_launch_b(cgb_handle* h, const cgb::FastTargets& tg, double* saveT, double* saveW)
{
    using namespace cgb;
    auto kern = k_lite_warp<CPL, MODE, MINB, SAVEW>;
    size_t smem = (size_t)(5 * CPL * 32 + 4 * h->nl * LC_STRIDE + 4 * 16) * sizeof(double);
    int blocks_per_sm = 0;
    CUDA_TRY(h, cgb_prepare_kernel(h, (const void*)kern, smem, &blocks_per_sm));
    long long grid = std::min<long long>((h->M + 3) / 4, (long long)h->num_sms * blocks_per_sm);
    CUDA_TRY(h, cudaMemsetAsync(h->dev.counter, 0, sizeof(unsigned long long), h->stream));
    kern<<<(unsigned)grid, 128, smem, h->stream>>>(h->dev, tg, saveT, saveW);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return CGB_OK;
}

// register budget: 2 resident CTAs (x4 warps) per SM -- MINB = 3 and 4 spill and were 8 % / 13 % slower on B200
template <int CPL, int MODE>
static int lite_warp_launch_m(cgb_handle* h, const cgb::FastTargets& tg, double* saveT, double* saveW)
{
    return saveW ? lite_warp_launch_b<CPL, MODE, 2, true>(h, tg, saveT, saveW) : lite_warp_launch_b<CPL, MODE, 2, false>(h, tg, saveT, saveW);
}

template <int CPL>
static int lite_warp_launch_t(cgb_handle* h, const cgb::FastTargets& tg, double* saveT, double* saveW)
{
    bool fw = true;
    for (int l = 0; l < h->nl; ++l) fw &= h->fc_kind[l] == CGB_FC_FREEWATER;
    return fw ? lite_warp_launch_m<CPL, 3>(h, tg, saveT, saveW) : lite_warp_launch_m<CPL, 0>(h, tg, saveT, saveW);
}

static int lite_warp_launch(cgb_handle* h, const cgb::FastTargets& tg, double* saveT, double* saveW)
{
    int cpl = (h->N + 31) / 32;
    switch (cpl) {
    case 1: return lite_warp_launch_t<1>(h, tg, saveT, saveW);
    case 2: return lite_warp_launch_t<2>(h, tg, saveT, saveW);
    case 3: case 4: return lite_warp_launch_t<4>(h, tg, saveT, saveW);
    case 5: case 6: return lite_warp_launch_t<6>(h, tg, saveT, saveW);
    case 7: return lite_warp_launch_t<7>(h, tg, saveT, saveW);
    case 8: return lite_warp_launch_t<8>(h, tg, saveT, saveW);
    case 9: return lite_warp_launch_t<9>(h, tg, saveT, saveW);
    case 10: return lite_warp_launch_t<10>(h, tg, saveT, saveW);
    case 11: case 12: return lite_warp_launch_t<12>(h, tg, saveT, saveW);
    default: return lite_warp_launch_t<16>(h, tg, saveT, saveW);
    }
}
