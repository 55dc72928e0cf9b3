// cgb_euler_sfcc_fast.cuh -- the SFCC + presolver-LUT fast path of the fused CGEuler kernel (the north_star target
// configuration: many columns of ONE soil class -- one SFCCPreSolver table -- that differ in forcing / boundary values).
//
// Same algorithm, data layout and step loop as k_heat_euler_fast (one warp per column, CPL cells per lane in registers for the
// whole save interval, software-pipelined, REDUX limiter reduction); what differs is the enthalpy inverse
// (Soils/soil_heat.jl:116-140):
//     T  = presolver interpolant at H                 (FreezeCurves.SFCCPreSolver, gridded-linear in H)
//     θw = sfcc(T)                                    (DallAmico / PainterKarra closed form)
// and how it is made cheap.  The general kernel (k_heat_step<.,1>) spends ~330 instructions per cell-step on it: ~18 LDS of
// per-(column,layer) constants, scattered global loads of knots (25.7 L1 sectors per request, profiles/r1f) and two pow()s
// through exp/log (~120 FP64 instructions).  Here, because the class is unique:
//   * every soil / curve constant is a kernel parameter (constant bank operand: no load, no register);
//   * the presolver knots (H_i, slope_i, T_i) and their bucket index are staged ONCE per CTA in shared memory;
//   * θw(T) is evaluated from a CURVE TABLE in shared memory: u = T* - (T + 273.15) is split by its floating-point exponent and
//     the top 6 mantissa bits into log-spaced bins (64 per octave); in a bin θw is a degree-5 polynomial in the mantissa
//     remainder r = m - c_k, |r| <= 1/128, fitted on the host at Chebyshev nodes of the closed form in long double and
//     VERIFIED there against it (cgb_tu_sfcc_fast.cu): measured 2e-15 relative for van Genuchten n = 1.6, 7e-15 for n = 2,
//     5e-14 for n = 3 -- the accuracy of a double-precision pow() chain -- with 6 FMAs and no transcendental (48-byte rows:
//     three 16-byte loads per cell; degree 7 on 16 bins per octave has the same accuracy with four); a class whose fit exceeds
//     1e-13 (very steep curves) is not eligible and runs through the general kernel.  Also checked on the device by
//     cgb_selftest_sfcc_curve (tests/test_gpu_sfcc_fast.py).  Van Genuchten n == 2 needs no table at all (one rsqrt).
//     u is formed with the reference's rounding sequence (TK = T + 273.15, then T* - TK), because near T* that cancellation --
//     not the curve evaluation -- decides the low bits of θw.
// ~31 FP64 instructions per cell-step instead of ~125, no global memory access in the step loop.
#pragma once
#include "cgb_euler.cuh"

namespace cgb {

template <int CPL, int THREADS, bool SAVE, bool N2>
__global__ void __launch_bounds__(THREADS, 1) k_heat_sfcc_fast(const __grid_constant__ Dev P, const __grid_constant__ SfccFast S,
                                                               const __grid_constant__ FastTargets TG,
                                                               double* __restrict__ saveT, double* __restrict__ saveW)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    {
        // stage the class tables once per CTA (coalesced 16-byte copies)
        const int4* src = reinterpret_cast<const int4*>(S.blob);
        int4* dst = reinterpret_cast<int4*>(smem_raw);
        for (int i = threadIdx.x; i < S.blob_bytes / 16; i += THREADS) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const double* __restrict__ sKH = reinterpret_cast<const double*>(smem_raw + S.off_seg);       // knot enthalpies H_i
    const double* __restrict__ sKS = sKH + S.nkp;                                                 // segment slopes dT/dH
    const double* __restrict__ sKT = sKS + S.nkp;                                                 // knot temperatures T_i
    const unsigned short* __restrict__ sBucket = reinterpret_cast<const unsigned short*>(smem_raw + S.off_bucket);
    const double* __restrict__ sCurve = reinterpret_cast<const double*>(smem_raw + S.off_curve);

    const int lane = threadIdx.x & 31;
    const int N = P.N;
    const long long M = P.M;

    // INTERLEAVED cell -> (lane, slot) mapping: cell = slot*32 + lane.  The 32 lanes of a warp then look up 32 ADJACENT cells at a
    // time, whose enthalpies and temperatures are close: they read the same or neighbouring table rows, which the shared-memory
    // banks serve as broadcasts / without conflicts.  (The contiguous mapping of the FreeWater kernel made every lane read its own
    // row: measured 109 shared-memory wavefronts per warp and cell-step for an ideal 23, LSU data pipe 92 % busy --
    // profiles/r2c_ncu_sfcc_fast_contiguous.md.)  Price: every edge now crosses a lane (3 rotating 64-bit shuffles per cell).
    // Static per-cell grid constants in registers: rho = dk/dk_up, Gd = (dk_up + dk)/(dxc dk_up), idk = 1/dk, so that the
    // harmonic-mean flux  -(T - T1) k1 k G / (dk k1 + dk_up k)  [math.jl:41,141]  becomes  (T1 - T)(k1 k) Gd / (rho k1 + k).
    double rho[CPL], Gd[CPL], idk[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        int i = j * 32 + lane;
        bool cell = i < N, edge = (i > 0 && i < N);
        double dku = edge ? P.dk[i - 1] : 1.0;
        rho[j] = edge ? P.dk[i] / dku : 1.0;
        Gd[j] = edge ? P.eG[i] / dku : 0.0;
        idk[j] = cell ? P.inv_dk[i] : 0.0;
    }
    const int jb = (lane == ((N - 1) & 31)) ? ((N - 1) >> 5) : -1;             // slot of the bottom cell in the lane that holds it
    const bool top_dirichlet = P.top_kind == 0, use_nf = P.use_nfactor != 0;   // bottom: Neumann, constant in time (launcher)
    const int fkind = P.top.kind;
    const int fn = P.top.n;
    const double* __restrict__ ft = P.top.t;
    const double* __restrict__ fy = P.top.y;
    const double ft_first = (fkind == 2) ? __ldg(ft) : 0.0, ft_last = (fkind == 2) ? __ldg(ft + fn - 1) : 0.0;
    const double two_over_period = (fkind == 1) ? 2.0 / P.top.period : 0.0, phase_over_pi = P.top.phase * 0.31830988618379067154;

    for (;;) {
        unsigned long long cu = 0;
        if (lane == 0) cu = atomicAdd(P.counter, 1ull);
        cu = __shfl_sync(FULL, cu, 0);
        if (cu >= (unsigned long long)M) break;
        const long long col = (long long)cu;

        double H[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = j * 32 + lane;
            H[j] = (i < N) ? P.u[(size_t)i * M + col] : 0.0;
        }
        double t = P.t[col];
        double dt = P.dt[col];
        const double nf = P.nf[col], nt = P.nt[col], toff = P.top_offset[col];
        const double vbot = P.bot_value ? P.bot_value[col] : P.bot.value;
        int status = 0;
        long long nsteps = 0;
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
                v = fma(P.top.amplitude, sinpi(fma(tq, two_over_period, phase_over_pi)), P.top.level);   // simple_bc.jl:37
            } else {
                v = P.top.value;
            }
            v += toff;
            if (top_dirichlet && use_nf) v = ((v <= 0.0) ? nf : nt) * v;             // heat_bc.jl:78-86
            return v;
        };

        // H -> T (presolver interpolant), θw (curve table), kc for all local cells, phase by phase so that the shared-memory
        // chains of different cells overlap
        double T[CPL], kc[CPL], thw[SAVE ? CPL : 1];
        auto stage_a = [&]() {
            int lo[CPL];
            unsigned wd[CPL];
            unsigned trips = 0;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                // bucket of H (flat extrapolation is part of the table: zero-slope end segments, see cgb_tu_sfcc_fast.cu); the
                // unsigned conversion saturates negative arguments to 0
                unsigned b = min(__double2uint_rz(fma(H[j], S.invbw, S.boff)), (unsigned)(S.nb - 1));
                wd[j] = sBucket[b];
                lo[j] = (int)(wd[j] & 0xfffu);
                trips = max(trips, wd[j]);
            }
            if (__any_sync(FULL, trips > 0xfffu)) {
                // knots pile up only at the kink of the freeze curve: a bucket there holds up to 2^trips segments
                const int nit = (int)(__reduce_max_sync(FULL, trips) >> 12);
                int hi[CPL];
#pragma unroll
                for (int j = 0; j < CPL; ++j) hi[j] = lo[j] + (1 << (wd[j] >> 12)) - 1;
                for (int k = 0; k < nit; ++k) {
#pragma unroll
                    for (int j = 0; j < CPL; ++j) {
                        int mid = (lo[j] + hi[j] + 1) >> 1;
                        bool p = sKH[mid] <= H[j];                          // +Inf padded beyond the last knot
                        lo[j] = p ? mid : lo[j];
                        hi[j] = p ? hi[j] : mid - 1;
                    }
                }
            }
            if (N2) {
                // van Genuchten n == 2 (the curve of examples/heat_sfcc_constantbc.jl): θw = θres + Δθ (1 + x²)^(-1/2) with
                // x = α|ψ0 - cfac u| for u = T* - TK > 0 and α|ψ0| above T*  --  one rsqrt, no table
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    T[j] = fma(H[j] - sKH[lo[j]], sKS[lo[j]], sKT[lo[j]]);
                    const double u = S.tstarK - (T[j] + 273.15);
                    const double x = fma(S.kappa, relu_bits(u), S.x0);
                    const double p = fma(S.dth, rsqrt_fast(fma(x, x, 1.0)), S.thres);
                    if (SAVE) thw[j] = p;
                    const double sk = fma(S.dK, p, S.kb);
                    kc[j] = sk * sk;
                }
            } else {
            int bin[CPL];
            double r[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                T[j] = fma(H[j] - sKH[lo[j]], sKS[lo[j]], sKT[lo[j]]);
                // u = T* - TK with the reference's rounding sequence (TK = T + 273.15 first)
                const double u = S.tstarK - (T[j] + 273.15);
                const int uh = __double2hiint(u), ul = __double2loint(u);
                // row 0 of the curve table is the constant θw of the thawed branch: everything below the table (u tiny, zero or
                // negative -- the sign bit makes the shifted word negative) lands there, without a select
                bin[j] = min(max((uh >> (20 - SF_BIN_BITS)) - S.idx_base, 0), S.nbins - 1);
                const double m = __hiloint2double((uh & 0x000fffff) | 0x3ff00000, ul);
                const double c = __hiloint2double(0x3ff00000 | (uh & SF_BIN_MASK) | SF_BIN_HALF, 0);
                r[j] = m - c;                                               // exact, |r| <= 2^-(SF_BIN_BITS+1) (finite for any u)
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const double* cf = sCurve + bin[j] * SF_CURVE_STRIDE;
                const double2 c01 = *reinterpret_cast<const double2*>(cf), c23 = *reinterpret_cast<const double2*>(cf + 2);
                const double2 c45 = *reinterpret_cast<const double2*>(cf + 4);
                // two interleaved Horner chains (even / odd powers): half the dependency depth of the plain chain
                const double r2 = r[j] * r[j];
                double pe = fma(c45.x, r2, c23.x), po = fma(c45.y, r2, c23.y);
                pe = fma(pe, r2, c01.x); po = fma(po, r2, c01.y);
                const double p = fma(po, r[j], pe);
                if (SAVE) thw[j] = p;
                const double sk = fma(S.dK, p, S.kb);
                kc[j] = sk * sk;
            }
            }
        };

        double vtop = top_value(t);
        stage_a();
        for (int kt = 0; kt < TG.n; ++kt) {
        const double t_target = TG.t[kt];
#pragma unroll 2
        while (t < t_target) {
            if (fkind == 2 && (t < ft_first || t > ft_last)) { status |= 4; break; }     // providers.jl:30-33
            if (SAVE && !(t + dt < t_target)) {
                // last step of the interval: the reference saves the cache of this RHS evaluation (state BEFORE the step)
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    int i = j * 32 + lane;
                    if (i < N) {
                        size_t o = (size_t)kt * TG.save_stride + (size_t)i * M + col;
                        if (saveT) saveT[o] = T[j];
                        if (saveW) saveW[o] = thw[SAVE ? j : 0];
                    }
                }
            }
            // ---- flux through the UPPER edge of every local cell; the upper neighbour is the previous lane's cell of the same slot
            // (lane 0: lane 31's cell of the previous slot = the value rotated in one slot earlier)
            double je[CPL];
            double Tprev = 0.0, kprev = 1.0;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const double Tr = __shfl_sync(FULL, T[j], (lane + 31) & 31), kr = __shfl_sync(FULL, kc[j], (lane + 31) & 31);
                const double T1 = (lane == 0) ? Tprev : Tr, k1 = (lane == 0) ? kprev : kr;
                Tprev = Tr; kprev = kr;
                const double den = fma(rho[j], k1, kc[j]);
                const double q = ((T1 - T[j]) * (k1 * kc[j])) * Gd[j];       // Gd == 0 on the top edge and beyond the bottom cell
                je[j] = q * rcp_fast3(den);
            }
            if (lane == 0) je[0] = top_dirichlet ? (kc[0] * (vtop - T[0])) * P.ctop : vtop;   // heat_bc.jl:9-17 / methods.jl:156
            // ---- divergence, limiter reduction, Euler update; the lower edge is the next lane's upper edge of the same slot
            // (lane 31: lane 0's edge of the NEXT slot), the bottom cell takes the Neumann value (heat_bc.jl:32-35)
            double ad[CPL];
            double jn[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) jn[j] = __shfl_sync(FULL, je[j], (lane + 1) & 31);
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                double jlow = (lane == 31) ? ((j + 1 < CPL) ? jn[j + 1 < CPL ? j + 1 : j] : 0.0) : jn[j];
                if (j == jb) jlow = vbot;
                const double dH = (je[j] - jlow) * idk[j];
                ad[j] = fabs(dH);
                H[j] = fma(dt, dH, H[j]);                                            // basic_solvers.jl:66
            }
#pragma unroll
            for (int w = 1; w < CPL; w *= 2)
#pragma unroll
                for (int j = 0; j + w < CPL; j += 2 * w) ad[j] = sel_max(ad[j + w], ad[j]);
            double maxabs = sel_max(ad[0], 0.0);
            unsigned mh = (unsigned)__double2hiint(maxabs), ml = (unsigned)__double2loint(maxabs);
            unsigned rh = __reduce_max_sync(FULL, mh);
            unsigned rl = __reduce_max_sync(FULL, mh == rh ? ml : 0u);
            const double tn = t + dt;
            ++nsteps;
            vtop = top_value(tn);
            stage_a();
            maxabs = __hiloint2double((int)rh, (int)rl);
            double mc = sel_min(sel_max(maxabs, 1e-300), 1e300);
            double lim = P.maxdelta * rcp_fast(mc);
            if (maxabs == INFINITY) lim = INFINITY;
            double dtn = sel_max(sel_min(lim, P.dtmax), P.dtmin);
            double rem = t_target - tn;
            dt = (rem > 0.0) ? sel_min(dtn, rem) : dtn;                              // integrator.jl:126-131
            t = tn;
        }
        if (status) break;
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = j * 32 + lane;
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

// self-test: θw from the curve table for n inputs T [°C] (one CTA stages the blob like the stepping kernel does)
__global__ void k_selftest_sfcc_curve(const __grid_constant__ SfccFast S, const double* __restrict__ Tin, long long n, double* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int4* src = reinterpret_cast<const int4*>(S.blob);
    int4* dst = reinterpret_cast<int4*>(smem_raw);
    for (int i = threadIdx.x; i < S.blob_bytes / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    __syncthreads();
    const double* sCurve = reinterpret_cast<const double*>(smem_raw + S.off_curve);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double u = S.tstarK - (Tin[i] + 273.15);
        const int uh = __double2hiint(u), ul = __double2loint(u);
        if (S.n2) {
            const double x = fma(S.kappa, relu_bits(u), S.x0);
            out[i] = fma(S.dth, rsqrt_fast(fma(x, x, 1.0)), S.thres);
            continue;
        }
        const int bin = min(max((uh >> (20 - SF_BIN_BITS)) - S.idx_base, 0), S.nbins - 1);
        const double m = __hiloint2double((uh & 0x000fffff) | 0x3ff00000, ul);
        const double c = __hiloint2double(0x3ff00000 | (uh & SF_BIN_MASK) | SF_BIN_HALF, 0);
        const double r = m - c;
        const double* cf = sCurve + bin * SF_CURVE_STRIDE;
        double p = cf[SF_CURVE_DEG];
#pragma unroll
        for (int k = SF_CURVE_DEG - 1; k >= 0; --k) p = fma(p, r, cf[k]);
        out[i] = p;
    }
}

} // namespace cgb
