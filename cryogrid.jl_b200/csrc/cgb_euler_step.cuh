// cgb_euler_step.cuh -- the stepping kernel of the general explicit path (any freeze curve, limiter, forcing).
//
// Same execution model as cgb_euler.cuh (one warp per column, H in registers for the whole save interval), but the step
// loop is software-pipelined like the FreeWater fast path: the RHS diagnostics of step k+1 (stage A: H -> T, θw, kc) are
// evaluated right after the Euler update of step k, while the warp-wide max|dH| reduction (integer REDUX unit) and the
// reciprocal for the next dt are still in flight.  Stage A works on the CPL cells of a lane phase by phase, so the
// table loads / transcendental chains of different cells overlap (the kernel is latency bound: 12 warps per SM).
//   k_heat_diag<CPL> (cgb_euler.cuh) is the one-evaluation diagnostics kernel.
// Reference: basic_solvers.jl:61-77 (CGEuler step), heat_conduction.jl:18-57,150-191, soil_heat.jl:116-140,
// heat_bc.jl:9-35,78-86, steplimiters.jl:15-18, integrator.jl:89,126-131, problem.jl:67-68 (what a save stores).
#pragma once
#include "cgb_euler.cuh"

namespace cgb {

template <int CPL, int MODE>
struct StageA {
    // H -> T, θw, kc for the CPL cells of a lane, plus dH/dT where it is used (MODE 0/2: Newton predictor; deriv: CFL limiter)
    static __device__ __forceinline__ void run(const Dev& P, const double* lcw, const int (&loff)[CPL], const double (&H)[CPL],
                                               double (&T)[CPL], double (&thw)[CPL], double (&kc)[CPL], double (&dHdT)[CPL],
                                               double dC, double dK, bool deriv)
    {
        if (MODE == 1 || MODE == 4) {
            double C[CPL], dth[CPL];
            if (deriv) stage_a_lut<CPL, true, MODE == 4>(P, lcw, loff, H, T, thw, C, dHdT, dth, kc, dC, dK);
            else stage_a_lut<CPL, false, MODE == 4>(P, lcw, loff, H, T, thw, C, dHdT, dth, kc, dC, dK);
        } else {
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                double C, dth;
                freezethaw_cell<MODE, false>(P, lcw + loff[j], H[j], dC, dK, T[j], thw[j], C, dHdT[j], dth, kc[j]);
            }
        }
    }
};

// 3 resident CTAs (12 warps) per SM = 168 registers up to 9 cells per lane: measured faster than 2 CTAs with ~200 registers
// and than 4 CTAs with 128 registers + spills; deeper grids (10..16 cells per lane) get 2 CTAs x 255 registers
#define CGB_STEP_BOUNDS __launch_bounds__(128, (CPL <= 9 ? 3 : 2))

template <int CPL, int MODE>
__global__ void CGB_STEP_BOUNDS k_heat_step(const __grid_constant__ Dev P, const __grid_constant__ FastTargets TG, Diag D)
{
    extern __shared__ double smem[];
    double* sInvDk = smem;
    double* sDk = sInvDk + CPL * 32;
    double* sW1 = sDk + CPL * 32;
    double* sW2 = sW1 + CPL * 32;
    double* sG = sW2 + CPL * 32;
    double* sLCall = sG + CPL * 32;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = P.N;
    const long long M = P.M;
    double* lcw = sLCall + warp * (P.n_layers * LC_STRIDE);
    // Cell -> (lane, slot) mapping.  Contiguous (cell = lane*CPL + slot: neighbours mostly in the same lane) for the modes whose
    // per-cell work is uniform.  INTERLEAVED (cell = slot*32 + lane) for the iterative modes (on-device Newton): the cells that
    // need a second evaluation are the dynamic ones near the surface / the freezing front, i.e. neighbours; contiguous, they sit
    // in a few lanes and every slot of the warp waits for them (measured: 2.2 evaluations per cell and warp for 1.24 per
    // thread); interleaved, they fill whole slots, and a slot's iteration loop runs with all lanes busy or not at all.
    // (CGB_LUT_ILV: A/B knob -- the LUT modes interleaved too, so that a warp-wide knot lookup addresses adjacent cells)
#ifndef CGB_LUT_ILV
#define CGB_LUT_ILV 0
#endif
    constexpr bool ILV = (MODE == 0 || MODE == 2) || (CGB_LUT_ILV && (MODE == 1 || MODE == 4));
    auto cell_of = [&](int ln, int j) { return ILV ? j * 32 + ln : ln * CPL + j; };

    if (MODE != 3 && MODE != 4) vgtab_init();             // exp / log tables of the van Genuchten powers (shared memory)
    for (int s = threadIdx.x; s < CPL * 32; s += blockDim.x) {
        int j = s >> 5, ln = s & 31;
        int i = cell_of(ln, j);
        bool cell = i < N, edge = (i > 0 && i < N);
        sInvDk[s] = cell ? P.inv_dk[i] : 0.0;
        sDk[s] = cell ? P.dk[i] : 1.0;
        sW1[s] = edge ? P.eW1[i] : 1.0;
        sW2[s] = edge ? P.eW2[i] : 1.0;
        sG[s] = edge ? P.eG[i] : 0.0;
    }
    __syncthreads();

    const double dC = P.ch_w - P.ch_i;
    const double dK = P.sk_w - P.sk_i;
    // local slot of the bottom cell (-1 / out of [0,CPL) in the lanes that do not hold it)
    const int jb = ILV ? ((lane == ((N - 1) & 31)) ? ((N - 1) >> 5) : -1) : (N - 1) - lane * CPL;
    const bool cfl = (P.limiter == 1) || MODE == 5;   // the temperature form needs dH/dT anyway
    const bool save = (D.T != nullptr) || (D.thw != nullptr);

    int loff[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        int i = cell_of(lane, j);
        loff[j] = P.cell_layer[i < N ? i : N - 1] * LC_STRIDE;
    }

    for (;;) {
        unsigned long long cu = 0;
        if (lane == 0) cu = atomicAdd(P.counter, 1ull);
        cu = __shfl_sync(FULL, cu, 0);
        if (cu >= (unsigned long long)M) break;
        const long long col = (long long)cu;

        __syncwarp();
        stage_layer_consts(P, col, lcw, lane);
        __syncwarp();

        double H[CPL], T[CPL], thw[CPL], kc[CPL], dHdT[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = cell_of(lane, j);
            H[j] = (i < N) ? P.u[(size_t)i * M + col] : 0.0;
            T[j] = H[j] / P.ch_w;   // Newton seed (soil_heat.jl:127)
        }
        double t = P.t[col];
        double dt = P.dt[col];
        const double nf = P.nf[col], nt = P.nt[col], toff = P.top_offset[col];
        int status = 0;
        long long nsteps = 0;
        ForcingCursor ctop, cbot;
        if (P.top.kind == 2) forcing_seek(P.top, ctop, fmin(fmax(t, __ldg(P.top.t)), __ldg(P.top.t + P.top.n - 1)));
        if (P.bot.kind == 2 && !P.bot_value) forcing_seek(P.bot, cbot, fmin(fmax(t, __ldg(P.bot.t)), __ldg(P.bot.t + P.bot.n - 1)));

        // boundary values at the START of a step (basic_solvers.jl:61-63)
        double vtop = 0.0, vbot = 0.0;
        auto boundary = [&](double tt) {
            vtop = forcing_eval(P.top, ctop, tt, status) + toff;
            if (P.top_kind == 0 && P.use_nfactor) vtop = ((vtop <= 0.0) ? nf : nt) * vtop;   // heat_bc.jl:78-86
            vbot = P.bot_value ? P.bot_value[col] : forcing_eval(P.bot, cbot, tt, status);
        };
        // device-side saveat: the column is carried through all save times of the launch in registers
        for (int kt = 0; kt < TG.n && !(status & 4); ++kt) {
        const double t_target = TG.t[kt];
        if (t < t_target) {
            boundary(t);
            StageA<CPL, MODE>::run(P, lcw, loff, H, T, thw, kc, dHdT, dC, dK, cfl);
        }

        while (t < t_target && !(status & 4)) {
            if (save && !(t + dt < t_target)) {
                // last step of the interval: the reference saves the diagnostics cached by this RHS evaluation (state BEFORE
                // the step): problem.jl:67-68, Numerics/statevars.jl:81-106, basic_solvers.jl:61-66
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    int i = cell_of(lane, j);
                    if (i < N) {
                        size_t o = (size_t)kt * TG.save_stride + (size_t)i * M + col;
                        if (D.T) D.T[o] = T[j];
                        if (D.thw) D.thw[o] = thw[j];
                    }
                }
            }
            // ---- CFL limiter: min over cells of courant * (∂H∂T / kc) * Δx² (heat_conduction.jl:150-181)
            double mincfl = INFINITY;
            if (P.limiter == 1) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    double dx = sDk[j * 32 + lane];
                    double c = P.courant * (dHdT[j] * rcp_fast(kc[j])) * (dx * dx);
                    if (cell_of(lane, j) < N && c < mincfl) mincfl = c;
                }
            }
            // ---- edge fluxes (upper edge of every local cell): harmonic mean folded into one reciprocal
            double Tup = 0.0, kup = 0.0;
            if (!ILV) { Tup = __shfl_up_sync(FULL, T[CPL - 1], 1); kup = __shfl_up_sync(FULL, kc[CPL - 1], 1); }
            double je[CPL + 1];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                double T1, k1;
                if (ILV) {
                    // upper neighbour = previous lane, same slot; lane 0 takes lane 31's previous slot (one rotating shuffle)
                    double sT = (lane == 31 && j > 0) ? T[j > 0 ? j - 1 : 0] : T[j];
                    double sk = (lane == 31 && j > 0) ? kc[j > 0 ? j - 1 : 0] : kc[j];
                    T1 = __shfl_sync(FULL, sT, (lane + 31) & 31);
                    k1 = __shfl_sync(FULL, sk, (lane + 31) & 31);
                } else {
                    T1 = (j == 0) ? Tup : T[j - 1];
                    k1 = (j == 0) ? kup : kc[j - 1];
                }
                int s = j * 32 + lane;
                double den = fma(sW2[s], k1, sW1[s] * kc[j]);
                double kk = k1 * kc[j];
                je[j] = -(((T[j] - T1) * kk) * sG[s]) * rcp_fast3(den);
            }
            // Top: Dirichlet -k[1](T[1]-T_ub)/(Δk[1]/2) (heat_bc.jl:9-17) ; Neumann: value (methods.jl:156)
            if (lane == 0) je[0] = (P.top_kind == 0) ? -(kc[0] * (T[0] - vtop)) * P.ctop : vtop;
            if (!ILV) je[CPL] = __shfl_down_sync(FULL, je[0], 1);
            // ---- divergence, Euler update, local max|dH|
            double ad[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                double jlow;
                if (ILV) {
                    // lower edge = upper edge of the next lane's cell in this slot; lane 31 takes lane 0's next slot
                    double sj = (lane == 0) ? ((j + 1 < CPL) ? je[j + 1 < CPL ? j + 1 : j] : 0.0) : je[j];
                    jlow = __shfl_sync(FULL, sj, (lane + 1) & 31);
                } else {
                    jlow = je[j + 1];
                }
                // Bottom: Dirichlet -k[end](T_lb-T[end])/(Δk[end]/2) (heat_bc.jl:18-27) ; Neumann: value
                if (j == jb) jlow = (P.bot_kind == 0) ? -(kc[j] * (vbot - T[j])) * P.cbot : vbot;
                double dH = (je[j] - jlow) * sInvDk[j * 32 + lane];
                if (MODE == 5) dH *= rcp_fast(dHdT[j]);                               // temperature form: du = dT = dH / dHdT
                ad[j] = fabs(dH);
                double inc = dt * dH;
                H[j] += inc;                                                          // u += dt*du (basic_solvers.jl:66)
                if (MODE == 0 || MODE == 2) T[j] = fma(inc, rcp_fast1(dHdT[j]), T[j]); // first-order predictor = next Newton seed
            }
#pragma unroll
            for (int w = 1; w < CPL; w *= 2)
#pragma unroll
                for (int j = 0; j + w < CPL; j += 2 * w) ad[j] = sel_max(ad[j + w], ad[j]);     // NaN is dropped, as by fmax
            double maxabs = sel_max(ad[0], 0.0);
            // start the warp reductions (integer REDUX unit) ...
            unsigned mh = (unsigned)__double2hiint(maxabs), ml = (unsigned)__double2loint(maxabs);
            unsigned rh = __reduce_max_sync(FULL, mh);
            unsigned rl = __reduce_max_sync(FULL, mh == rh ? ml : 0u);
            bool bad_cfl = false;
            if (P.limiter == 1) {
                bad_cfl = __any_sync(FULL, !(mincfl > 0.0));                        // <= 0 or NaN: the reference asserts (tile.jl:174)
                mincfl = warp_min_pos(mincfl > 0.0 ? mincfl : INFINITY);
            }
            const double tn = t + dt;
            ++nsteps;
            // ... and overlap them with the next step's boundary values and diagnostics
            if (tn < t_target) {
                boundary(tn);
                StageA<CPL, MODE>::run(P, lcw, loff, H, T, thw, kc, dHdT, dC, dK, cfl);
            }
            // ---- next dt: timestep() from the OLD du (basic_solvers.jl:76-77, steplimiters.jl:15-18), then the saveat! clip
            maxabs = __hiloint2double((int)rh, (int)rl);
            double mc = sel_min(sel_max(maxabs, 1e-300), 1e300);                     // keeps the reciprocal finite
            double lim = P.maxdelta * rcp_fast(mc);                                  // 2 Newton steps: correctly rounded 1/x
            if (maxabs == INFINITY) lim = INFINITY;                                  // |Δmax/inf| = 0 -> "not > 0" -> Inf (:189)
            if (P.limiter == 1) {
                lim = sel_min(lim, mincfl);                                          // Inf stays Inf (heat_conduction.jl:180)
                if (bad_cfl) { status |= 8; lim = 0.0; }
            }
            double dtn = sel_max(sel_min(lim, P.dtmax), P.dtmin);
            double rem = t_target - tn;
            dt = (rem > 0.0) ? sel_min(dtn, rem) : dtn;                              // integrator.jl:126-131 (:89 is implied)
            t = tn;
        }
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = cell_of(lane, j);
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
