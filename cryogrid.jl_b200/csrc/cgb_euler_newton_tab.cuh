// cgb_euler_newton_tab.cuh -- SFCC with PER-COLUMN parameters through the on-device Newton solver (BASELINE cfg4), with the
// freeze curve of every column tabulated on the fly.
//
// k_heat_step<.,2> evaluates the closed form (two pow() through exp / log, ~150 instructions) at every Newton evaluation, ~1.45
// evaluations per cell, warp and step: ~190 FP64-pipe instructions per cell-step.  Per-column parameters rule out a shared curve
// table -- but one WARP owns one COLUMN for a whole launch (hundreds of steps), so the warp fits the column's own curve table at
// column start: θw(u), u = T* - (T + 273.15), on log-spaced bins (exponent + top 4 mantissa bits, 16 per octave, 2^-16 K <= u < 64 K)
// as degree-7 polynomials in the mantissa remainder, by interpolation at 8 Chebyshev nodes of the closed form (Newton divided
// differences with precomputed node distances, expanded to monomials; 352 bins per column, 11 per lane; fitted once and kept in HBM).
// A Newton evaluation is then a table row (4 x 16-byte shared-memory loads) + two Horner chains (θw and dθw/du): ~45 instructions.
// Same table form as the single-class fast path (cgb_euler_sfcc_fast.cuh), same interleaved cell -> lane mapping (32 adjacent
// cells per warp-wide load address neighbouring rows).  Outside the table (within 15 µK of T*, or below -64 °C) the closed form is
// evaluated directly.  The Newton iteration itself is the one of sfcc_newton_solve<false>: side of the kink decided exactly
// (thawed branch in closed form), safeguarded step on the frozen branch, stop at |ΔT| <= 1e-9, first-order carry of θw.
// Eligibility (host): one layer, every column SFCC + Newton solver, MaxDelta limiter, constant Neumann bottom.
// Reference: Soils/soil_heat.jl:116-140 (freezethaw!, SFCCNewtonSolver call site), heat_methods.jl:93,105.
#pragma once
#include "cgb_euler.cuh"

namespace cgb {

constexpr int NT_EMIN = -16, NT_EMAX = 5;                       // table covers 2^NT_EMIN <= u < 2^(NT_EMAX+1) [K]
constexpr int NT_ROWS = (NT_EMAX + 1 - NT_EMIN) * 16;           // 352
constexpr int NT_STRIDE = 10;                                   // doubles per row: 8 coefficients + 2 padding (80-byte rows: LDS.128 bank skew)
constexpr int NT_IDX_BASE = (NT_EMIN + 1023) << 4;

// Chebyshev nodes t_j = cos((2j+1) pi/16) of [-1, 1] (compile-time: the node distances of the divided differences fold to constants)
#define NT_NODES {0.98078528040323043, 0.83146961230254524, 0.55557023301960218, 0.19509032201612825, \
                  -0.19509032201612825, -0.55557023301960218, -0.83146961230254524, -0.98078528040323043}

// per-column curve: θw(u) = θres + d (1 + (α (|ψ0| + cfac u))^n)^-m for u > 0
struct ColumnCurve {
    double thres, d, alpha, n, m, apsi0, cfac;
    __device__ __forceinline__ double operator()(double u) const
    {
        const double x = alpha * fma(cfac, u, apsi0);
        if (n == 2.0) return fma(d, rsqrt_fast(fma(x, x, 1.0)), thres);
        double xn, pb;
        vg_powers_body<true>(sel_max(x, 1e-300), n, m, xn, pb);
        return fma(d, pb, thres);
    }
    // θw and dθw/du by the closed form (outside the table).  Out of line on purpose: it is called from the seven pass-2 bodies and
    // hardly ever runs; inlined, the kernel was 9 900 SASS instructions instead of 8 100 and 6-9 % slower (instruction fetch).
    __device__ __noinline__ double eval(double u, double& dthdu) const
    {
        const double x = alpha * fma(cfac, u, apsi0);
        if (n == 2.0) {
            const double rs = rsqrt_fast(fma(x, x, 1.0));
            dthdu = -(d * alpha * cfac) * x * (rs * rs * rs);
            return fma(d, rs, thres);
        }
        double xn, pb;
        const double xs = sel_max(x, 1e-300);
        vg_powers_body<true>(xs, n, m, xn, pb);
        dthdu = (x > 0.0) ? -(d * m * n * alpha * cfac) * (xn * rcp_fast(xs)) * (pb * rcp_fast(1.0 + xn)) : 0.0;
        return fma(d, pb, thres);
    }
};

// one bin: interpolating polynomial at the 8 Chebyshev nodes of r in [-1/32, 1/32], u = 2^E (c_k + r), as monomials in r
template <class F>
__device__ __forceinline__ void nt_fit_bin(const F& f, int E, int k, double* __restrict__ row)
{
    constexpr double NT_NODE[8] = NT_NODES;
    const double scale = __hiloint2double((E + 1023) << 20, 0);          // 2^E
    const double ck = 1.0 + (k + 0.5) * 0.0625;
    double dd[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dd[j] = f(scale * fma(NT_NODE[j], 0.03125, ck));
    // Newton divided differences over t (node distances are compile-time constants after unrolling)
#pragma unroll
    for (int lev = 1; lev < 8; ++lev)
#pragma unroll
        for (int j = 7; j >= lev; --j) dd[j] = (dd[j] - dd[j - 1]) * (1.0 / (NT_NODE[j] - NT_NODE[j - lev]));
    // Newton form -> monomials of t: p <- p (t - t_i) + dd_i, i = 6 .. 0, starting from p = dd_7
    double c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i] = 0.0;
    c[0] = dd[7];
#pragma unroll
    for (int i = 6; i >= 0; --i) {
#pragma unroll
        for (int q = 7 - i; q >= 1; --q) c[q] = fma(-NT_NODE[i], c[q], c[q - 1]);
        c[0] = fma(-NT_NODE[i], c[0], dd[i]);
    }
    // t = 32 r
    double s = 1.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { row[q] = c[q] * s; s *= 32.0; }
}

template <int CPL, int WARPS, bool SAVE>
__global__ void __launch_bounds__(WARPS * 32, 1) k_heat_newton_tab(const __grid_constant__ Dev P, const __grid_constant__ FastTargets TG,
                                                                  double* __restrict__ saveT, double* __restrict__ saveW,
                                                                  double* __restrict__ cache, int cache_valid)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    vgtab_init();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* __restrict__ tab = reinterpret_cast<double*>(smem_raw) + (size_t)warp * NT_ROWS * NT_STRIDE;
    const int N = P.N;
    const long long M = P.M;
    const double dC = P.ch_w - P.ch_i, dK = P.sk_w - P.sk_i, L = P.L;

    // interleaved mapping and folded grid constants, as in k_heat_sfcc_fast
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
    const int jb = (lane == ((N - 1) & 31)) ? ((N - 1) >> 5) : -1;
    const bool top_dirichlet = P.top_kind == 0, use_nf = P.use_nfactor != 0;
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

        // ---- the column's constants (every lane holds a copy) and its curve table
        double lc[LC_STRIDE];
        layer_consts(P, col, 0, lc);
        const double tstarK = lc[LC_TSTAR], tstarC = lc[LC_NW_TSTARC], hstar = lc[LC_NW_HSTAR], thu = lc[LC_NW_THU], Cu = lc[LC_NW_CU];
        const double cbs = lc[LC_CBS], kb = lc[LC_KB];
        ColumnCurve f;
        f.thres = lc[LC_THRES]; f.d = sel_max(lc[LC_THWI], lc[LC_POR]) - lc[LC_THRES]; f.alpha = lc[LC_ALPHA]; f.n = lc[LC_N]; f.m = lc[LC_M];
        f.apsi0 = fabs(lc[LC_PSI0]); f.cfac = lc[LC_CFAC];
        __syncwarp();
        // The table is fitted in the first launch after the description changed and kept in HBM (22.5 KB per column, when it
        // fits: cgb_newton_tab_prepare); later launches (one per save interval) read it back with 16-byte loads.
        double2* __restrict__ crow = cache ? reinterpret_cast<double2*>(cache + (size_t)col * (NT_ROWS * 8)) : nullptr;
        if (crow && cache_valid) {
            for (int q = lane; q < NT_ROWS * 4; q += 32)
                *reinterpret_cast<double2*>(tab + (q >> 2) * NT_STRIDE + (q & 3) * 2) = __ldcs(crow + q);
        } else {
            for (int b = lane; b < NT_ROWS; b += 32) nt_fit_bin(f, NT_EMIN + (b >> 4), b & 15, tab + b * NT_STRIDE);
            if (crow) {
                __syncwarp();
                for (int q = lane; q < NT_ROWS * 4; q += 32)
                    __stcs(crow + q, *reinterpret_cast<const double2*>(tab + (q >> 2) * NT_STRIDE + (q & 3) * 2));
            }
        }
        __syncwarp();

        double H[CPL], T[CPL], kc[CPL], dHdT[CPL], thw[SAVE ? CPL : 1];
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            int i = j * 32 + lane;
            H[j] = (i < N) ? P.u[(size_t)i * M + col] : 0.0;
            T[j] = H[j] / P.ch_w;                   // Newton seed (soil_heat.jl:127)
            dHdT[j] = 1.0;
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
                v = fma(P.top.amplitude, sinpi(fma(tq, two_over_period, phase_over_pi)), P.top.level);
            } else {
                v = P.top.value;
            }
            v += toff;
            if (top_dirichlet && use_nf) v = ((v <= 0.0) ? nf : nt) * v;
            return v;
        };

        // θw and dθw/dT from the table row of u (bin must be in range): 4 x 16-byte loads, Horner with the derivative carried
        // by synthetic division (q <- q r + p; p <- p r + c)
        auto table_eval = [&](int bin, int uh, int ul, double& dth) -> double {
            const double m = __hiloint2double((uh & 0x000fffff) | 0x3ff00000, ul);
            const double c = __hiloint2double(0x3ff08000 | (uh & 0x000f0000), 0);
            const double r = m - c;
            const double* cf = tab + bin * NT_STRIDE;
            const double2 c01 = *reinterpret_cast<const double2*>(cf), c23 = *reinterpret_cast<const double2*>(cf + 2);
            const double2 c45 = *reinterpret_cast<const double2*>(cf + 4), c67 = *reinterpret_cast<const double2*>(cf + 6);
            double p = fma(c67.y, r, c67.x), q = c67.y;
            q = fma(q, r, p); p = fma(p, r, c45.y);
            q = fma(q, r, p); p = fma(p, r, c45.x);
            q = fma(q, r, p); p = fma(p, r, c23.y);
            q = fma(q, r, p); p = fma(p, r, c23.x);
            q = fma(q, r, p); p = fma(p, r, c01.y);
            q = fma(q, r, p); p = fma(p, r, c01.x);
            // dθ/du = p'(r) / 2^E ; 2^-E from the exponent field of u ; dθ/dT = -dθ/du
            const double inv_scale = __hiloint2double((2046 - ((uh >> 20) & 0x7ff)) << 20, 0);
            dth = -(q * inv_scale);
            return p;
        };
        // the same with the closed form outside the table
        auto curve = [&](double Tq, double& dth) -> double {
            const double u = tstarK - (Tq + 273.15);
            const int uh = __double2hiint(u), ul = __double2loint(u);
            const int bin = (uh >> 16) - NT_IDX_BASE;
            double th;
            if ((unsigned)bin < (unsigned)NT_ROWS) {
                th = table_eval(bin, uh, ul, dth);
            } else if (u > 0.0) {
                double dthdu;
                th = f.eval(u, dthdu);
                dth = -dthdu;
            } else {
                th = thu; dth = 0.0;
            }
            return th;
        };

        // H -> T, θw, kc, dH/dT for all local cells.  Pass 1: ONE Newton step from the predictor for every slot, branch-free
        // (seven independent dependency chains per lane: this kernel runs 8 warps per SM, so the latency has to be covered by
        // instruction-level parallelism).  A lane is finished when the step was inside the table, inside (-273, T*) and no longer
        // than the tolerance -- the stopping rule of the safeguarded iteration below.  Pass 2: that iteration (bracket restarted)
        // for the slots where some lane is not finished: the top 32-64 cells in most steps, nothing below.
        const double rcpCu = rcp_fast(Cu);
        auto stage_a = [&]() {
            unsigned again = 0;
            double thl[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const double Hj = H[j];
                const bool frozen = Hj < hstar;
                const double T0 = T[j];
                const double u = tstarK - (T0 + 273.15);
                const int uh = __double2hiint(u), ul = __double2loint(u);
                const int bin = (uh >> 16) - NT_IDX_BASE;
                const bool inr = (unsigned)bin < (unsigned)NT_ROWS;      // false for T0 >= T*, T0 < T* - 64 K and NaN
                double dth;
                const double th = table_eval(inr ? bin : 0, uh, ul, dth);
                const double Cj = fma(dC, th, cbs);
                const double r = fma(T0, Cj, L * th) - Hj;
                const double dH = fma(dth, fma(T0, dC, L), Cj);            // heat_methods.jl:105
                const double dT = -r * rcp_fast1(dH);
                const double Tn = T0 + dT;
                const bool conv = inr & (Tn < tstarC) & (fabs(dT) <= 1e-9 * sel_max(1.0, fabs(Tn)));   // no short circuit: branch-free
                const double Tt = (Hj - L * thu) * rcpCu;                  // thawed branch: linear in T
                T[j] = frozen ? (inr ? Tn : T0) : Tt;
                thl[j] = frozen ? fma(dth, dT, th) : thu;
                dHdT[j] = frozen ? dH : Cu;
                again |= (unsigned)(frozen & !conv) << j;
            }
            again = __reduce_or_sync(FULL, again);
            if (again) {
#pragma unroll
                for (int j = 0; j < CPL; ++j) {
                    if ((again >> j) & 1u) {
                        const double Hj = H[j];
                        double Tj = T[j], th = thl[j], dH = dHdT[j], dth = 0.0;
                        double lo = -273.0, hi = tstarC;
                        bool active = Hj < hstar;                              // lanes that are finished just confirm
                        if (active && !(Tj > lo && Tj < hi)) Tj = hi;          // a cell that has just crossed the kink starts at it
                        for (int it = 0; it < 200 && __any_sync(FULL, active); ++it) {
                            if (active) {
                                th = curve(Tj, dth);
                                const double Cj = fma(dC, th, cbs);
                                const double r = fma(Tj, Cj, L * th) - Hj;
                                dH = fma(dth, fma(Tj, dC, L), Cj);
                                if (r > 0.0) hi = Tj; else lo = Tj;
                                double Tn = fma(-r, rcp_fast1(dH), Tj);
                                const double tol = 1e-9 * sel_max(1.0, fabs(Tn));
                                const bool small = fabs(Tn - Tj) <= tol;
                                const bool safe = (Tn > lo && Tn < hi) || (small && Tn >= lo && Tn <= hi);
                                if (!safe) Tn = 0.5 * (lo + hi);
                                const double dT = Tn - Tj;
                                Tj = Tn;
                                if (fabs(dT) <= tol) {                         // converged (or the bracket has collapsed)
                                    th = fma(dth, dT, th);
                                    active = false;
                                }
                            }
                        }
                        T[j] = Tj; dHdT[j] = dH; thl[j] = th;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                if (SAVE) thw[j] = thl[j];
                const double sk = fma(dK, thl[j], kb);
                kc[j] = sk * sk;
            }
        };

        double vtop = top_value(t);
        stage_a();
        for (int kt = 0; kt < TG.n; ++kt) {
        const double t_target = TG.t[kt];
        while (t < t_target) {
            if (fkind == 2 && (t < ft_first || t > ft_last)) { status |= 4; break; }
            if (SAVE && !(t + dt < t_target)) {
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
            double je[CPL];
            double Tprev = 0.0, kprev = 1.0;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const double Tr = __shfl_sync(FULL, T[j], (lane + 31) & 31), kr = __shfl_sync(FULL, kc[j], (lane + 31) & 31);
                const double T1 = (lane == 0) ? Tprev : Tr, k1 = (lane == 0) ? kprev : kr;
                Tprev = Tr; kprev = kr;
                const double den = fma(rho[j], k1, kc[j]);
                const double q = ((T1 - T[j]) * (k1 * kc[j])) * Gd[j];
                je[j] = q * rcp_fast3(den);
            }
            if (lane == 0) je[0] = top_dirichlet ? (kc[0] * (vtop - T[0])) * P.ctop : vtop;
            double ad[CPL], jn[CPL];
#pragma unroll
            for (int j = 0; j < CPL; ++j) jn[j] = __shfl_sync(FULL, je[j], (lane + 1) & 31);
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                double jlow = (lane == 31) ? ((j + 1 < CPL) ? jn[j + 1 < CPL ? j + 1 : j] : 0.0) : jn[j];
                if (j == jb) jlow = vbot;
                const double dH = (je[j] - jlow) * idk[j];
                ad[j] = fabs(dH);
                const double inc = dt * dH;
                H[j] += inc;                                                          // basic_solvers.jl:66
                T[j] = fma(inc, rcp_fast1(dHdT[j]), T[j]);                            // first-order predictor = next Newton seed
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
            if (tn < t_target || kt + 1 < TG.n) { vtop = top_value(tn); stage_a(); }
            maxabs = __hiloint2double((int)rh, (int)rl);
            double mc = sel_min(sel_max(maxabs, 1e-300), 1e300);
            double lim = P.maxdelta * rcp_fast(mc);
            if (maxabs == INFINITY) lim = INFINITY;
            double dtn = sel_max(sel_min(lim, P.dtmax), P.dtmin);
            double rem = t_target - tn;
            dt = (rem > 0.0) ? sel_min(dtn, rem) : dtn;
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

} // namespace cgb
