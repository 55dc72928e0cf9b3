// cgb_euler_fast.cuh -- the FreeWater + MaxDelta fast path of the fused CGEuler kernel (BASELINE cfg2).
//
// Same algorithm and data layout as k_heat_step (cgb_euler_step.cuh): one warp per column, CPL cells per lane held in
// registers for the whole save interval.  What differs is the arithmetic budget -- the kernel is bound by the FP64
// pipe (64 lanes/clk/SM), so everything here minimises FP64-pipe instructions per cell-step:
//   * one clamp serves both θw and the enthalpy inverse:  hc = clamp(H, 0, Lθ);  x = hc/Lθ;  T = (H - hc)/C
//     (== heat_conduction.jl:18-57 whenever θwi >= 1e-8, which the launcher checks on the host)
//   * C and sqrt-sum conductivity are single FMAs in x; 1/C and the edge denominator use MUFU.RCP64H + 2 Newton steps
//   * harmonic mean + flux share one reciprocal:  j = -(T2-T1) k1 k2 G / (Δ2 k1 + Δ1 k2)        (math.jl:41,141)
//   * static grid constants (Δk, 1/Δk, G) live in registers; per-layer soil constants are broadcast LDS
//   * max|dH| over the warp is reduced on the integer REDUX unit (two 32-bit reduce-max), not on the FP64 pipe
//   * next dt = Δmax * rcp(max|dH|); forcing interpolation multiplies by a cached 1/(t_{k+1}-t_k)
// ~26 FP64 instructions per cell-step + ~3 amortised per-step instructions.
#pragma once
#include "cgb_euler.cuh"

namespace cgb {

constexpr int FC_STRIDE = 8;   // fast-path per-(column,layer) constants
enum FastConst { FC_LTH = 0, FC_INVLTH, FC_CB, FC_KB, FC_DCTH, FC_DKTH, FC_THWI };

template <bool CUBIC>
__device__ __forceinline__ double rcp_sel(double a) { return CUBIC ? rcp_fast3(a) : rcp_fast(a); }

template <int CPL, int MINB, bool CUBIC, bool SAVE>
__global__ void __launch_bounds__(128, MINB) k_heat_euler_fast(const __grid_constant__ Dev P, const __grid_constant__ FastTargets TG,
                                                                double* __restrict__ saveT, double* __restrict__ saveW)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = P.N;
    const long long M = P.M;
    double* lcw = smem + warp * (P.n_layers * FC_STRIDE);

    const double dC = P.ch_w - P.ch_i;
    const double dK = P.sk_w - P.sk_i;

    // static per-cell grid constants in registers: rho = dk / dk_up, Gd = (dk_up + dk) / (dxc dk_up), idk = 1 / dk, so that the
    // harmonic-mean flux (T1 - T) k1 k G / (dk k1 + dk_up k)  [math.jl:41,141]  is  (T1 - T)(k1 k) Gd / (rho k1 + k): one FMA for the
    // denominator instead of a multiply and an FMA (round 2; one FP64 instruction of 28 per cell-step)
    double rho[CPL], idk[CPL], G[CPL];
    bool is_edgeN[CPL];   // this slot's upper edge is the bottom boundary (global edge index N)
    int loff[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        int i = lane * CPL + j;
        bool cell = i < N, edge = (i > 0 && i < N);
        double dku = edge ? P.dk[i - 1] : 1.0;
        rho[j] = edge ? P.dk[i] / dku : 1.0;
        idk[j] = cell ? P.inv_dk[i] : 0.0;
        G[j] = edge ? P.eG[i] / dku : 0.0;
        loff[j] = P.cell_layer[cell ? i : N - 1] * FC_STRIDE;
        is_edgeN[j] = (i == N);
    }
    const bool top_dirichlet = P.top_kind == 0, use_nf = P.use_nfactor != 0;   // bottom: Neumann only (launcher)
    const int fkind = P.top.kind;
    const int fn = P.top.n;
    const double* __restrict__ ft = P.top.t;
    const double* __restrict__ fy = P.top.y;
    const double ft_first = (fkind == 2) ? __ldg(ft) : 0.0, ft_last = (fkind == 2) ? __ldg(ft + fn - 1) : 0.0;

    for (;;) {
        unsigned long long cu = 0;
        if (lane == 0) cu = atomicAdd(P.counter, 1ull);
        cu = __shfl_sync(FULL, cu, 0);
        if (cu >= (unsigned long long)M) break;
        const long long col = (long long)cu;

        __syncwarp();
        for (int l = lane; l < P.n_layers; l += 32) {
            size_t ix = (size_t)l * M + col;
            double por = P.por[ix], sat = P.sat[ix], org = P.org[ix];
            double thwi = por * sat;
            double thm = (1.0 - por) * (1.0 - org), tho = (1.0 - por) * org;
            double tha = 1.0 - thwi - thm - tho;
            double* c = lcw + l * FC_STRIDE;
            double Lth = P.L * thwi;                     // θwi >= 1e-8 guaranteed by the launcher
            c[FC_LTH] = Lth;
            c[FC_INVLTH] = 1.0 / Lth;
            c[FC_CB] = P.ch_i * thwi + P.ch_a * tha + P.ch_m * thm + P.ch_o * tho;
            c[FC_KB] = P.sk_i * thwi + P.sk_a * tha + P.sk_m * thm + P.sk_o * tho;
            c[FC_DCTH] = dC * thwi;
            c[FC_DKTH] = dK * thwi;
            c[FC_THWI] = thwi;
        }
        __syncwarp();

        double H[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = lane * CPL + j;
            H[j] = (i < N) ? P.u[(size_t)i * M + col] : 0.0;
        }
        double t = P.t[col];
        double dt = P.dt[col];
        const double nf = P.nf[col], nt = P.nt[col], toff = P.top_offset[col];
        const double vbot = P.bot_value ? P.bot_value[col] : P.bot.value;   // launcher: bottom is constant in time here
        // Neumann bottom flux enters as the additive term of the flux FMA of the slot below the last cell
        double bflux[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) bflux[j] = is_edgeN[j] ? vbot : 0.0;
        const double bflux_next = (lane * CPL + CPL == N) ? vbot : 0.0;   // bottom edge is slot 0 of the (absent or next) lane
        int status = 0;
        long long nsteps = 0;
        // forcing cursor (warp-uniform); finv = 1/(t[k+1]-t[k]) comes from a host-built table (no division in the loop)
        int fidx = 0;
        double fta = 0.0, ftb = 0.0, fya = 0.0, fyb = 0.0, finv = 0.0;
        const double* __restrict__ fiv = P.top_inv_dt;
        if (fkind == 2) {
            double ts = fmin(fmax(t, ft_first), ft_last);
            int lo = 0, hi = fn - 1;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (__ldg(ft + mid) <= ts) lo = mid; else hi = mid; }
            fidx = lo; fta = __ldg(ft + lo); ftb = __ldg(ft + lo + 1); fya = __ldg(fy + lo); fyb = __ldg(fy + lo + 1);
            finv = __ldg(fiv + lo);
        }
        // upper boundary value at time tq (warp-uniform).  Out-of-range times are clamped here; the BoundsError status is
        // raised at the top of the step that would actually use them.
        auto top_value = [&](double tq) -> double {
            double v;
            if (fkind == 2) {
                double tc = sel_min(sel_max(tq, ft_first), ft_last);
                while (tc >= ftb && fidx < fn - 2) {
                    ++fidx; fta = ftb; fya = fyb;
                    ftb = __ldg(ft + fidx + 1); fyb = __ldg(fy + fidx + 1); finv = __ldg(fiv + fidx);
                }
                double w = (tc - fta) * finv;
                v = fma(w, fyb, (1.0 - w) * fya);
            } else if (fkind == 1) {
                v = P.top.amplitude * sin(2.0 * M_PI * (tq / P.top.period) + P.top.phase) + P.top.level;   // simple_bc.jl:37
            } else {
                v = P.top.value;
            }
            v += toff;
            if (top_dirichlet && use_nf) v = ((v <= 0.0) ? nf : nt) * v;             // heat_bc.jl:78-86
            return v;
        };
        // H -> T, kc for all local cells (heat_conduction.jl:18-57, thermal_properties.jl:37)
        double T[CPL], kc[CPL];
        auto stage_a = [&]() {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const double* c = lcw + loff[j];
                double hc = sel_min(relu_bits(H[j]), c[FC_LTH]);
                double x = hc * c[FC_INVLTH];
                double C = fma(c[FC_DCTH], x, c[FC_CB]);
                double s = fma(c[FC_DKTH], x, c[FC_KB]);
                kc[j] = s * s;
                T[j] = (H[j] - hc) * rcp_sel<CUBIC>(C);
            }
        };

        // Software-pipelined step loop: the diagnostics (stage A) and the boundary value of step n+1 are computed while
        // the warp reduction and the reciprocal of step n's limiter are in flight -- they do not depend on dt_{n+1}.
        double vtop = top_value(t);
        stage_a();
        for (int kt = 0; kt < TG.n; ++kt) {
        const double t_target = TG.t[kt];
        // unrolled x3: the loop-carried copies of the software pipeline (next -> current T, kc, boundary value) disappear into
        // register renaming across the copies (47 of 399 instructions per step were moves); measured 3.52 -> 3.62e11 (x2: 3.59, x4: 3.61)
        // one explicit Euler step of the column at time t with step dt (fluxes, update, limiter, next dt) -- also computes T, kc and
        // the boundary value of the NEXT step (software pipeline)
        auto one_step = [&]() {
                // ---- edge fluxes (upper edge of every local cell): harmonic mean + flux with one reciprocal
                double Tup = __shfl_up_sync(FULL, T[CPL - 1], 1);
                double kup = __shfl_up_sync(FULL, kc[CPL - 1], 1);
                double je[CPL + 1];
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    double T1 = (j == 0) ? Tup : T[j - 1];
                    double k1 = (j == 0) ? kup : kc[j - 1];
                    double den = fma(rho[j], k1, kc[j]);
                    double q = ((T1 - T[j]) * (k1 * kc[j])) * G[j];
                    je[j] = fma(q, rcp_sel<CUBIC>(den), bflux[j]);   // q == 0 on the bottom edge (G = 0): je = -Qgeo there
                }
                if (lane == 0)   // heat_bc.jl:9-17 (Dirichlet) / methods.jl:156 (Neumann)
                    je[0] = top_dirichlet ? (kc[0] * (vtop - T[0])) * P.ctop : vtop;
                je[CPL] = __shfl_down_sync(FULL, je[0], 1);
                if (bflux_next != 0.0 || lane == 31) je[CPL] = bflux_next;               // heat_bc.jl:32-35 / methods.jl:156
                // ---- divergence, limiter reduction, Euler update
                double ad[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    double dH = (je[j] - je[j + 1]) * idk[j];
                    ad[j] = fabs(dH);
                    H[j] = fma(dt, dH, H[j]);                                            // basic_solvers.jl:66
                }
                // max |dH| over the local cells as a tree (depth log2 CPL instead of CPL dependent compares); NaN is dropped
#pragma unroll
                for (int w = 1; w < CPL; w *= 2)
#pragma unroll
                    for (int j = 0; j + w < CPL; j += 2 * w) ad[j] = sel_max(ad[j + w], ad[j]);
                double maxabs = sel_max(ad[0], 0.0);
                // start the warp reduction (integer REDUX unit) ...
                unsigned mh = (unsigned)__double2hiint(maxabs), ml = (unsigned)__double2loint(maxabs);
                unsigned rh = __reduce_max_sync(FULL, mh);
                unsigned rl = __reduce_max_sync(FULL, mh == rh ? ml : 0u);
                const double tn = t + dt;
                ++nsteps;
                // ... and overlap it with the next step's boundary value and diagnostics
                vtop = top_value(tn);
                stage_a();
                // ---- next dt (basic_solvers.jl:76-77, heat_conduction.jl:183-191, steplimiters.jl:15-18), saveat! clip
                maxabs = __hiloint2double((int)rh, (int)rl);
                double mc = sel_min(sel_max(maxabs, 1e-300), 1e300);                     // keeps the reciprocal finite
                double lim = P.maxdelta * rcp_fast(mc);                                  // 2 Newton steps: correctly rounded 1/x
                if (maxabs == INFINITY) lim = INFINITY;                                  // |Δmax/inf| = 0 -> "not > 0" -> Inf
                double dtn = sel_max(sel_min(lim, P.dtmax), P.dtmin);
                double rem = t_target - tn;
                dt = (rem > 0.0) ? sel_min(dtn, rem) : dtn;                              // integrator.jl:126-131 (:89 is implied)
                t = tn;
        };
        if (SAVE) {
            // interior steps of the save interval: the same loop as the plain stepping kernel (with the save code inside the unrolled
            // loop the kernel spilled 200 bytes; measured 36.4 -> 35.6 ms on the 73 daily saves of the bench window); then the last
            // step of the interval, which saves first
#pragma unroll 3
            while (t + dt < t_target) {
                if (fkind == 2 && (t < ft_first || t > ft_last)) { status |= 4; break; }     // providers.jl:30-33
                one_step();
            }
            if (!status && t < t_target) {
                if (fkind == 2 && (t < ft_first || t > ft_last)) status |= 4;
                else {
                    {
                        // Last step of the interval: what the reference SAVES for diagnostic variables is the cache of this (the last)
                        // RHS evaluation, i.e. T, θw of the state BEFORE the step (problem.jl:67-68 -> Numerics/statevars.jl:81-106
                        // retrieve(); CGEuler evaluates the tile before u += dt*du, basic_solvers.jl:61-66).  Reproduced verbatim.
#pragma unroll
                        for (int j = 0; j < CPL; ++j) {
                            int i = lane * CPL + j;
                            if (i < N) {
                                size_t o = (size_t)kt * TG.save_stride + (size_t)i * M + col;
                                if (saveT) saveT[o] = T[j];
                                if (saveW) {
                                    const double* c = lcw + loff[j];
                                    double hc = sel_min(relu_bits(H[j]), c[FC_LTH]);
                                    saveW[o] = (hc * c[FC_INVLTH]) * c[FC_THWI];
                                }
                            }
                        }
                    }
                    one_step();
                }
            }
        } else {
#pragma unroll 3
            while (t < t_target) {
                if (fkind == 2 && (t < ft_first || t > ft_last)) { status |= 4; break; }     // providers.jl:30-33
                one_step();
            }
        }
        if (status) break;
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = lane * CPL + j;
            if (i < N) P.u[(size_t)i * M + col] = H[j];
        }
        if (lane == 0) {
            P.t[col] = t;
            P.dt[col] = dt;
            P.steps[col] += nsteps;
            if (status) P.status[col] |= status;
        }
    }
}

} // namespace cgb
