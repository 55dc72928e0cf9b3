// cgb_euler.cuh -- device functions of the explicit enthalpy-heat path (layer constants, freeze curves, SFCC lookup /
// Newton, forcing) and the one-evaluation diagnostics kernel k_heat_diag; the stepping kernels are in
// cgb_euler_step.cuh (any freeze curve) and cgb_euler_fast.cuh (FreeWater).  Execution model of all of them:
//
// One WARP owns one column; lane l holds cells [l*CPL, (l+1)*CPL) of H in REGISTERS for the whole
// save interval, so a column is read from HBM once per cgb_advance_to and written once, however many
// steps it takes.  Per step a lane does, for its CPL cells (independent chains -> ILP):
//   H -> (θw, C, T, kc)            FreeWater closed form / SFCC LUT / SFCC Newton
//                                   (heat_conduction.jl:18-57, soil_heat.jl:116-140, thermal_properties.jl:53-66)
//   edge conductivity + flux        harmonic mean folded into one reciprocal (math.jl:41,141)
//   boundary fluxes                 heat_bc.jl:9-35, methods.jl:156
//   divergence + Euler update       math.jl:68, basic_solvers.jl:66
//   max|dH| -> next dt              REDUX / shuffles; heat_conduction.jl:183-191, basic_solvers.jl:76-77,
//                                   integrator.jl:126-131 (clip to the save time)
// Neighbour cells across lanes travel by __shfl_up/down.  Static grid constants and the per-(column,layer)
// soil constants live in shared memory ([slot][lane] layout: conflict-free).  Warps fetch columns from a
// global cursor (no CTA-wide barrier in the loop), so a slow column never stalls its CTA.
#pragma once
#include "cgb_common.cuh"

namespace cgb {

// ---- per-(column,layer) constants ---------------------------------------------------------------
__device__ __forceinline__ void layer_consts(const Dev& P, long long col, int l, double* c)
{
    size_t ix = (size_t)l * P.M + col;
    double por = P.por[ix], sat = P.sat[ix], org = P.org[ix];
    double thwi = por * sat;
    double thm = (1.0 - por) * (1.0 - org);
    double tho = (1.0 - por) * org;
    double tha = 1.0 - thwi - thm - tho;
    double thtot = fmax(1e-8, thwi);
    c[LC_THWI] = thwi;
    c[LC_LTH] = P.L * thwi;
    c[LC_THTOT] = thtot;
    c[LC_INVLTHTOT] = 1.0 / (P.L * thtot);
    c[LC_CB] = P.ch_i * thwi + P.ch_a * tha + P.ch_m * thm + P.ch_o * tho;
    c[LC_KB] = P.sk_i * thwi + P.sk_a * tha + P.sk_m * thm + P.sk_o * tho;
    c[LC_POR] = por;
    int kind = P.fc_kind[l];
    int* ci = reinterpret_cast<int*>(c + LC_KIND);
    ci[0] = kind; ci[1] = 0;
    int* to = reinterpret_cast<int*>(c + LC_TABOFF);
    int* tl = reinterpret_cast<int*>(c + LC_TABLEN);
    to[0] = 0; to[1] = 0; tl[0] = 0; tl[1] = 0;
    if ((kind & 0xff) != 0) {
        double alpha = P.vg_alpha[ix], n = P.vg_n[ix], thres = P.theta_res[ix];
        double beta = 1.0, omega = 1.0;
        if ((kind & 0xff) == 2) { beta = P.beta[ix]; omega = P.omega[ix]; }
        // temperature form: the reference calls sfcc(T; θsat) with the functor's DEFAULT saturation 1.0 (soil_heat.jl:106)
        double sat_e = P.tform ? 1.0 : thwi / por;
        double thtot_s = sat_e * por;
        double thsat_e = fmax(thtot_s, por);
        double psi0 = vg_psi(thtot_s, thsat_e, thres, alpha, n);
        double TmK = P.Tm[ix] + 273.15;
        double gT = omega * P.g / P.Lsl * psi0;
        double Tstar = TmK + gT * TmK;
        c[LC_THRES] = thres; c[LC_ALPHA] = alpha; c[LC_N] = n; c[LC_M] = 1.0 - 1.0 / n;
        c[LC_PSI0] = psi0; c[LC_TSTAR] = Tstar; c[LC_CFAC] = beta * P.Lsl / (P.g * Tstar);
        c[LC_SAT] = sat_e;
        // sfccheatcap base (θw = 0) with the reference's (1-θsat) quirk (soil_heat.jl:13-14)
        c[LC_CBS] = P.ch_i * thwi + P.ch_a * (por - thwi) + P.ch_m * (thm * (1.0 - por)) + P.ch_o * (tho * (1.0 - por));
        if (P.tform) {
            // no inverse-enthalpy solver in the temperature form
        } else if (((kind >> 8) & 0xff) == 0) {
            int id = P.table_map ? P.table_map[ix] : l;
            int off = P.lut_off[id], nk = P.lut_len[id];
            to[0] = off; to[1] = nk;
            tl[0] = P.lut_extrap[id]; tl[1] = 0;
            reinterpret_cast<int*>(c + LC_BOFF)[0] = P.lut_boff[id];
            double hmin = P.lutH[off], hmax = P.lutH[off + nk - 1], nb = (double)P.lut_nb[id];
            c[LC_HMIN] = hmin; c[LC_HMAX] = hmax; c[LC_INVBW] = nb / (hmax - hmin); c[LC_NBM1] = nb - 1.0;
        } else {
            // the thawed branch (T >= T*) is linear in T: θw = θ(ψ0), C = C(θw); its lower end is the kink enthalpy H*
            double dth;
            double thu = vg_theta(psi0, thsat_e, thres, alpha, n, c[LC_M], dth);
            double Cu = fma(P.ch_w - P.ch_i, thu, c[LC_CBS]);
            double TstarC = Tstar - 273.15;
            c[LC_NW_THU] = thu; c[LC_NW_CU] = Cu; c[LC_NW_TSTARC] = TstarC;
            c[LC_NW_HSTAR] = TstarC * Cu + P.L * thu;
        }
    }
}

__device__ __forceinline__ void stage_layer_consts(const Dev& P, long long col, double* lc, int lane)
{
    for (int l = lane; l < P.n_layers; l += 32) layer_consts(P, col, l, lc + l * LC_STRIDE);
}

// The freeze curve has a kink at T* (ψ stops depending on T above it), where dH/dT jumps by orders of magnitude; a Newton
// iteration started on the wrong side needs many bracketing steps, and in a warp every lane waits for the one cell at the
// freezing front.  So the side is decided first, exactly: H >= H(T*) is the thawed branch, which is LINEAR in T (closed
// form, no iteration); below it the iteration runs on the smooth frozen branch only, bracketed by [-273, T*].
// EXACT: iterate to 1e-15 relative and re-evaluate θw at the root (diagnostics; the same root as the oracle's safeguarded
// Newton over the whole range).  Otherwise (stepping): stop once |ΔT| <= 1e-9 -- Newton is quadratic, so the accepted T is
// accurate to ~1e-15 -- and carry θw, C to the new T with a first-order correction instead of another evaluation.
template <bool EXACT>
__device__ __forceinline__ void sfcc_newton_solve(const double* lc, double L, double dC, double H, double& T, double& thw,
                                                  double& C, double& dHdT, double& dth)
{
    if (H >= lc[LC_NW_HSTAR]) {
        thw = lc[LC_NW_THU];
        C = lc[LC_NW_CU];
        T = EXACT ? (H - L * thw) / C : (H - L * thw) * rcp_fast(C);
        dth = 0.0;
        dHdT = C;
        return;
    }
    double lo = -273.0, hi = lc[LC_NW_TSTARC];
    if (!(T > lo && T < hi)) T = hi;                   // a cell that has just crossed the kink starts at it
    for (int it = 0; it < 200; ++it) {
        thw = sfcc_theta<!EXACT>(lc, T, dth);
        C = fma(dC, thw, lc[LC_CBS]);
        double r = (T * C + L * thw) - H;
        dHdT = C + dth * (L + T * dC);                 // heat_methods.jl:105
        if (r > 0.0) hi = T; else lo = T;
        double Tn = EXACT ? T - r / dHdT : fma(-r, rcp_fast1(dHdT), T);   // a Newton correction: 1e-13 relative is plenty
        // a Newton step is taken if it falls strictly inside the bracket, or ON its end when it is a converged step (the
        // strict test alone would bisect there; the inclusive test alone lets Newton 2-cycle between the bracket ends).
        // NaN -> bisect.
        const double tol = (EXACT ? 1e-15 : 1e-9) * fmax(1.0, fabs(Tn));
        bool small = fabs(Tn - T) <= tol;
        bool safe = (Tn > lo && Tn < hi) || (small && Tn >= lo && Tn <= hi);
        if (!safe) Tn = 0.5 * (lo + hi);
        double dT = Tn - T;
        bool done = fabs(dT) <= tol;                   // a bisection step this small means the bracket has collapsed
        T = Tn;
        if (done) {
            if (EXACT) {
                thw = sfcc_theta<!EXACT>(lc, T, dth);
                C = fma(dC, thw, lc[LC_CBS]);
                dHdT = C + dth * (L + T * dC);
            } else {
                thw = fma(dth, dT, thw);
                C = fma(dC, thw, lc[LC_CBS]);
            }
            break;
        }
    }
}

__device__ __forceinline__ double lut_T(const Dev& P, const double* lc, double H)
{
    const int* to = reinterpret_cast<const int*>(lc + LC_TABOFF);
    const int* tl = reinterpret_cast<const int*>(lc + LC_TABLEN);
    int off = to[0], boff = reinterpret_cast<const int*>(lc + LC_BOFF)[0];
    return lut_eval_bucket(P.lutH + off, P.lutT + off, P.lutS + off, P.lutB + boff, lc[LC_HMIN], lc[LC_HMAX], lc[LC_INVBW], lc[LC_NBM1], tl[0], H);
}

// H -> (T, θw, C, dHdT, dθw/dT, kc) for one cell.
// MODE: 0 = freeze curve / solver read from the layer constants (generic), 1 = every layer is SFCC + presolver LUT (4: and n == 2),
//       2 = every layer is SFCC + Newton, 3 = every layer is FreeWater.  EXACT: fully converged Newton (diagnostics).
template <int MODE, bool EXACT>
__device__ __forceinline__ void freezethaw_body(const Dev& P, const double* lc, double H, double dC, double dK,
                                                double& T, double& thw, double& C, double& dHdT, double& dth, double& kc)
{
    int kind = (MODE == 0) ? reinterpret_cast<const int*>(lc + LC_KIND)[0] : (MODE == 3 ? 0 : ((MODE == 1 || MODE == 4) ? 1 : (1 | (1 << 8))));
    if (MODE == 5 || (MODE == 0 && P.tform)) {
        // temperature form (soil_heat.jl:100-113): H IS the temperature; θw, ∂θw∂T straight from the curve, the REGULAR heat capacity
        T = H;
        thw = sfcc_theta<!EXACT>(lc, T, dth);
        C = fma(dC, thw, lc[LC_CB]);
        dHdT = C + dth * (P.L + T * dC);   // heat_methods.jl:105
    } else if ((kind & 0xff) == 0) {
        // FreeWater: heat_conduction.jl:18-57 (θw = clamp(H/Lθ,0,1) θtot; T piecewise; ∂H∂T = 1e8 at T == 0)
        // One clamp instead of two clamps and a three-way select (the form of the FreeWater fast kernel): hc = clamp(H, 0, Lθ) is the
        // latent part of H, θw = hc/(Lθ) θtot, and T = (H - hc)/C is H/C below, EXACTLY 0 inside and (H - Lθ)/C above the phase change.
        double hc = sel_min(relu_bits(H), lc[LC_LTH]);
        thw = (hc * lc[LC_INVLTHTOT]) * lc[LC_THTOT];
        C = fma(dC, thw, lc[LC_CB]);
        T = (H - hc) * rcp_fast3(C);
        dHdT = (T == 0.0) ? 1e8 : C;
        dth = 0.0;
    } else if (((kind >> 8) & 0xff) == 0) {
        // SFCC + presolver: soil_heat.jl:116-140 -- T from the table, then θw, ∂θw∂T at T
        T = lut_T(P, lc, H);
        thw = sfcc_theta<!EXACT>(lc, T, dth);
        C = fma(dC, thw, lc[LC_CBS]);
        dHdT = C + dth * (P.L + T * dC);   // heat_methods.jl:105
    } else {
        sfcc_newton_solve<EXACT>(lc, P.L, dC, H, T, thw, C, dHdT, dth);
    }
    double s = fma(dK, thw, lc[LC_KB]);
    kc = s * s;
}

// The generic (MODE 0) body carries all three branches; it is kept out of line so that the CPL-times unrolled step
// loop stays inside the instruction cache (the inlined version was 170 KB of SASS and stalled on instruction fetch).
template <bool EXACT>
__device__ __noinline__ void freezethaw_generic(const Dev& P, const double* lc, double H, double dC, double dK,
                                                double& T, double& thw, double& C, double& dHdT, double& dth, double& kc)
{
    freezethaw_body<0, EXACT>(P, lc, H, dC, dK, T, thw, C, dHdT, dth, kc);
}

template <int MODE, bool EXACT = true>
__device__ __forceinline__ void freezethaw_cell(const Dev& P, const double* lc, double H, double dC, double dK,
                                                double& T, double& thw, double& C, double& dHdT, double& dth, double& kc)
{
    if (MODE == 0) freezethaw_generic<EXACT>(P, lc, H, dC, dK, T, thw, C, dHdT, dth, kc);
    else freezethaw_body<MODE, EXACT>(P, lc, H, dC, dK, T, thw, C, dHdT, dth, kc);
}

// Stage A for stratigraphies whose layers ALL use an SFCC with the presolver LUT (MODE 1), written phase by phase over
// the CPL cells of a lane so that the independent chains overlap: (1) bucket loads, (2) ONE fixed-trip bisection loop
// that advances all cells together (the per-cell formulation ran CPL dependent load chains one after the other),
// (3) T from the segment, (4) matric potential, (5) the van Genuchten powers of all cells in one block (exp/log chains
// of different cells interleave; warp-uniform rsqrt shortcut when every n == 2), (6) C, dH/dT, kc.
// Same arithmetic as freezethaw_body<1, .>.
// DERIV = false skips dθw/dT, C and dH/dT (only the CFL limiter and the diagnostics kernel need them).  N2 = every van
// Genuchten n of the problem is 2 (decided on the host): the rsqrt shortcut, and no exp/log code in the kernel at all.
template <int CPL, bool DERIV, bool N2>
__device__ __forceinline__ void stage_a_lut(const Dev& P, const double* lcw, const int (&loff)[CPL], const double (&H)[CPL],
                                            double (&T)[CPL], double (&thw)[CPL], double (&C)[CPL], double (&dHdT)[CPL],
                                            double (&dth)[CPL], double (&kc)[CPL], double dC, double dK)
{
    int lo[CPL], hi[CPL];
    double Hq[CPL];
    unsigned trips = 0;
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const double* lc = lcw + loff[j];
        double hmin = lc[LC_HMIN], hmax = lc[LC_HMAX];
        Hq[j] = (reinterpret_cast<const int*>(lc + LC_TABLEN)[0] == 0) ? sel_min(sel_max(H[j], hmin), hmax) : H[j];
        double fb = sel_min(sel_max((Hq[j] - hmin) * lc[LC_INVBW], 0.0), lc[LC_NBM1]);
        unsigned w = __ldg(P.lutB + reinterpret_cast<const int*>(lc + LC_BOFF)[0] + (int)fb);
        lo[j] = reinterpret_cast<const int*>(lc + LC_TABOFF)[0] + (int)(w & 0xffffffu);
        hi[j] = lo[j] + (1 << (w >> 24)) - 1;
        trips = max(trips, w >> 24);
    }
    // trip count of the bisection = the widest candidate range any cell of the warp drew (0 or 1 almost everywhere; the
    // knots of a presolver table pile up only at the kink of the freeze curve)
    const int nit = (int)__reduce_max_sync(FULL, trips);
    for (int k = 0; k < nit; ++k) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int mid = (lo[j] + hi[j] + 1) >> 1;
            bool p = __ldg(P.lutH + mid) <= Hq[j];
            lo[j] = p ? mid : lo[j];
            hi[j] = p ? hi[j] : mid - 1;
        }
    }
    double x[CPL], dpsi[CPL], pb[CPL], q[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const double* lc = lcw + loff[j];
        const int i = lo[j];
        T[j] = fma(Hq[j] - __ldg(P.lutH + i), __ldg(P.lutS + i), __ldg(P.lutT + i));
        double TK = T[j] + 273.15, Tstar = lc[LC_TSTAR];
        bool frozen = (Tstar - TK) >= 0.0;
        double psi = frozen ? fma(lc[LC_CFAC], TK - Tstar, lc[LC_PSI0]) : lc[LC_PSI0];
        if (DERIV) dpsi[j] = frozen ? lc[LC_CFAC] : 0.0;
        x[j] = (psi <= 0.0) ? fabs(-lc[LC_ALPHA] * psi) : -1.0;      // -1 marks ψ > 0 (θw = θsat)
    }
    if (N2) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            double rs = rsqrt_fast(fma(x[j], x[j], 1.0));
            pb[j] = rs;
            if (DERIV) q[j] = (lcw + loff[j])[LC_ALPHA] * x[j] * (rs * rs * rs);
        }
    } else {
        // the powers of all cells in one straight-line block: with exp_fast / log_pos (~150 instructions per cell) the CPL
        // copies fit the instruction cache and their dependency chains interleave (+8 % over an out-of-line call per cell)
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const double* lc = lcw + loff[j];
            double n = lc[LC_N], m = lc[LC_M];
            double xs = sel_max(x[j], 1e-300);
            double xn, p;
            vg_powers_body<true>(xs, n, m, xn, p);
            bool pos = x[j] > 0.0;
            pb[j] = pos ? p : 1.0;
            if (DERIV) q[j] = pos ? m * n * lc[LC_ALPHA] * (xn * rcp_fast(xs)) * (p * rcp_fast(1.0 + xn)) : 0.0;
        }
    }
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        const double* lc = lcw + loff[j];
        double thsat = sel_max(lc[LC_THWI], lc[LC_POR]);
        double d = thsat - lc[LC_THRES];
        bool sat = x[j] < 0.0;
        thw[j] = sat ? thsat : fma(d, pb[j], lc[LC_THRES]);
        if (DERIV) {
            dth[j] = sat ? 0.0 : (d * q[j]) * dpsi[j];
            C[j] = fma(dC, thw[j], lc[LC_CBS]);
            dHdT[j] = C[j] + dth[j] * (P.L + T[j] * dC);     // heat_methods.jl:105
        }
        double sk = fma(dK, thw[j], lc[LC_KB]);
        kc[j] = sk * sk;
    }
}

// Warp-uniform forcing evaluation with a cursor into the table (t only moves forward).
struct ForcingCursor {
    int idx;
    double ta, tb, ya, yb;
};

__device__ __forceinline__ void forcing_seek(const Forcing& F, ForcingCursor& c, double t)
{
    // guess from the mean knot spacing (exact for the evenly spaced forcing tables of the reference's loaders), verified;
    // a bisection -- ~log2(n) dependent global loads per column and launch -- only when the guess misses
    int lo = min(max((int)((t - __ldg(F.t)) * F.inv_step), 0), F.n - 2);
    if (!(__ldg(F.t + lo) <= t && t <= __ldg(F.t + lo + 1))) {
        lo = 0;
        int hi = F.n - 1;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (__ldg(F.t + mid) <= t) lo = mid; else hi = mid;
        }
    }
    c.idx = lo;
    c.ta = __ldg(F.t + lo); c.tb = __ldg(F.t + lo + 1);
    c.ya = __ldg(F.y + lo); c.yb = __ldg(F.y + lo + 1);
}

__device__ __forceinline__ double forcing_eval(const Forcing& F, ForcingCursor& c, double t, int& status)
{
    if (F.kind == 0) return F.value;
    // simple_bc.jl:37  A sin(2π t/P + φ) + c, as sinpi(2 (t/P) + φ/π): exact argument reduction, a third of sin()'s instructions
    if (F.kind == 1) return fma(F.amplitude, sinpi(fma(2.0, t / F.period, F.phase * 0.31830988618379067154)), F.level);
    if (t < __ldg(F.t) || t > __ldg(F.t + F.n - 1)) { status |= 4; return nan(""); }              // providers.jl:30-33
    while (t >= c.tb && c.idx < F.n - 2) {
        ++c.idx;
        c.ta = c.tb; c.ya = c.yb;
        c.tb = __ldg(F.t + c.idx + 1); c.yb = __ldg(F.y + c.idx + 1);
    }
    if (t < c.ta) forcing_seek(F, c, t);
    double w = (t - c.ta) / (c.tb - c.ta);
    return (1.0 - w) * c.ya + w * c.yb;
}

// -------------------------------------------------------------------------------------------------
// One RHS evaluation at the current state and time t: every diagnostic the reference caches (T, θw, C, ∂H∂T, ∂θw∂T, kc, k,
// jH, dH) for cgb_rhs / cgb_get_diag / the "*_now" saved variables.  Any freeze curve / solver (generic cell function with
// the fully converged Newton: bitwise the oracle's procedure).  The state is not modified.
template <int CPL>
__global__ void __launch_bounds__(128) k_heat_diag(const __grid_constant__ Dev P, double t, Diag D)
{
    extern __shared__ double smem[];
    double* sInvDk = smem;
    double* sW1 = sInvDk + CPL * 32;
    double* sW2 = sW1 + CPL * 32;
    double* sG = sW2 + CPL * 32;
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
        sInvDk[s] = cell ? P.inv_dk[i] : 0.0;
        sW1[s] = edge ? P.eW1[i] : 1.0;
        sW2[s] = edge ? P.eW2[i] : 1.0;
        sG[s] = edge ? P.eG[i] : 0.0;
    }
    __syncthreads();

    const double dC = P.ch_w - P.ch_i;
    const double dK = P.sk_w - P.sk_i;
    const int jb = (N - 1) - lane * CPL;   // local slot of the bottom cell (if in [0,CPL))

    int loff[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) {
        int i = lane * CPL + j;
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

        // ---- boundary values at t (basic_solvers.jl:61-63)
        int status = 0;
        ForcingCursor ctop, cbot;
        if (P.top.kind == 2) forcing_seek(P.top, ctop, fmin(fmax(t, __ldg(P.top.t)), __ldg(P.top.t + P.top.n - 1)));
        if (P.bot.kind == 2 && !P.bot_value) forcing_seek(P.bot, cbot, fmin(fmax(t, __ldg(P.bot.t)), __ldg(P.bot.t + P.bot.n - 1)));
        double vtop = forcing_eval(P.top, ctop, t, status) + P.top_offset[col];
        if (P.top_kind == 0 && P.use_nfactor) vtop = ((vtop <= 0.0) ? P.nf[col] : P.nt[col]) * vtop;   // heat_bc.jl:78-86
        const double vbot = P.bot_value ? P.bot_value[col] : forcing_eval(P.bot, cbot, t, status);

        // ---- H -> T, θw, C, dH/dT, dθw/dT, kc
        double T[CPL], kc[CPL], thw[CPL], C[CPL], dHdT[CPL], dth[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = lane * CPL + j;
            double H = (i < N) ? P.u[(size_t)i * M + col] : 0.0;
            T[j] = H / P.ch_w;   // Newton seed (soil_heat.jl:127)
            freezethaw_generic<true>(P, lcw + loff[j], H, dC, dK, T[j], thw[j], C[j], dHdT[j], dth[j], kc[j]);
        }
        // ---- edge conductivities and fluxes (upper edge of every local cell)
        double Tup = __shfl_up_sync(FULL, T[CPL - 1], 1);
        double kup = __shfl_up_sync(FULL, kc[CPL - 1], 1);
        double je[CPL + 1], ke[CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            double T1 = (j == 0) ? Tup : T[j - 1];
            double k1 = (j == 0) ? kup : kc[j - 1];
            int s = j * 32 + lane;
            double w1 = sW1[s], w2 = sW2[s];
            double r = rcp_fast(fma(w2, k1, w1 * kc[j]));
            double kk = k1 * kc[j];
            je[j] = -(((T[j] - T1) * kk) * sG[s]) * r;
            ke[j] = (kk * (w1 + w2)) * r;
        }
        if (lane == 0) {
            // Top: Dirichlet -k[1](T[1]-T_ub)/(Δk[1]/2)  (heat_bc.jl:9-17) ; Neumann: value (methods.jl:156)
            je[0] = (P.top_kind == 0) ? -(kc[0] * (T[0] - vtop)) * P.ctop : vtop;
            ke[0] = kc[0];
        }
        je[CPL] = __shfl_down_sync(FULL, je[0], 1);
        // ---- divergence and output
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            // Bottom: Dirichlet -k[end](T_lb-T[end])/(Δk[end]/2) (heat_bc.jl:18-27) ; Neumann: value
            if (j == jb) je[j + 1] = (P.bot_kind == 0) ? -(kc[j] * (vbot - T[j])) * P.cbot : vbot;
            int i = lane * CPL + j;
            if (i < N) {
                size_t o = (size_t)i * M + col;
                if (D.T) D.T[o] = T[j];
                if (D.thw) D.thw[o] = thw[j];
                if (D.C) D.C[o] = C[j];
                if (D.dHdT) D.dHdT[o] = dHdT[j];
                if (D.dthwdT) D.dthwdT[o] = dth[j];
                if (D.kc) D.kc[o] = kc[j];
                const double dHv = (je[j] - je[j + 1]) * sInvDk[j * 32 + lane];
                if (D.dH) D.dH[o] = dHv;
                if (D.dT) D.dT[o] = dHv / dHdT[j];                                  // temperatureflux!: heat_conduction.jl:79-84
                if (D.Hd) D.Hd[o] = T[j] * C[j] + P.L * thw[j];                     // enthalpy(T, C, L, θw): heat_methods.jl:93
                if (D.k) D.k[o] = ke[j];
                if (D.jH) D.jH[o] = je[j];
                if (i == N - 1) {
                    size_t ob = (size_t)N * M + col;
                    if (D.k) D.k[ob] = kc[j];
                    if (D.jH) D.jH[ob] = je[j + 1];
                }
            }
        }
        (void)status;   // a diagnostic evaluation outside the forcing range returns NaN fluxes; the column status is left alone
    }
}

} // namespace cgb
