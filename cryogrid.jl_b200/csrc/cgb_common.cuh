// cgb_common.cuh -- shared device structs and math helpers for libcgb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace cgb {

constexpr unsigned FULL = 0xffffffffu;

// per-(column, layer) derived constants staged in shared memory (one block per warp)
constexpr int LC_STRIDE = 24;
constexpr int LUT_NB_MIN = 64, LUT_NB_MAX = 16384;   // uniform H-buckets per presolver table: power of two ~ 8x the knot count (host-built index)
enum LayerConst {
    LC_THWI = 0,   // θwi = por*sat                        (Soils/para/simple.jl:19,87)
    LC_LTH,        // L*θwi                                 (heat_conduction.jl:40)
    LC_THTOT,      // max(1e-8, θwi)                        (FreezeCurves.freewater)
    LC_INVLTHTOT,  // 1/(L*θtot)
    LC_CB,         // heat capacity at θw = 0               (heat_methods.jl:52-57)
    LC_KB,         // Σ sqrt(k)θ at θw = 0                  (thermal_properties.jl:37)
    LC_POR,        // θsat
    LC_THRES,
    LC_ALPHA,
    LC_N,
    LC_M,          // 1 - 1/n
    LC_PSI0,       // ψ(θtot)
    LC_TSTAR,      // T* [K]
    LC_CFAC,       // β Lsl/(g T*)
    LC_CBS,        // sfccheatcap at θw = 0 (with the (1-θsat) quirk, soil_heat.jl:13-14)
    LC_KIND,       // int: freeze-curve kind | solver<<8
    LC_TABOFF,     // int2: LUT knot offset, LUT length
    LC_TABLEN,     // int2: extrapolation mode, unused
    LC_SAT,        // θwi/por
    LC_BOFF,       // int2: bucket-array offset, unused
    LC_HMIN,       // first knot
    LC_HMAX,       // last knot
    LC_INVBW,      // nb / (Hmax - Hmin)
    LC_NBM1,       // nb - 1 (as double: clamp bound of the bucket index)
    // layers solved by the on-device Newton iteration reuse the table slots:
    LC_NW_HSTAR = LC_HMIN,    // enthalpy at the kink T = T* (everything above is thawed: closed form)
    LC_NW_THU = LC_HMAX,      // θw of the thawed branch, θ(ψ0)
    LC_NW_CU = LC_INVBW,      // heat capacity of the thawed branch
    LC_NW_TSTARC = LC_NBM1    // T* in °C
};

struct Forcing {
    int kind;                 // 0 constant, 1 periodic, 2 table
    int n;
    double value;
    double period, amplitude, phase, level;
    const double* t;
    const double* y;
    double inv_step;          // (n-1)/(t[n-1]-t[0]): first guess of a knot index (exact for evenly spaced tables)
};

// Save times of one launch, passed by value: a column is advanced through all of them without leaving its warp, and the
// diagnostics to be saved at target k go to saveT/saveW + k*save_stride (device-side saveat).
constexpr int FAST_MAX_TARGETS = 16;
struct FastTargets {
    double t[FAST_MAX_TARGETS];
    int n;
    long long save_stride;   // elements between consecutive saves in saveT / saveW
};

// Everything a kernel needs; passed by value (__grid_constant__).
struct Dev {
    int N, n_layers, nvar;
    long long M;
    double* u;                // [nvar*N*M] column-fastest
    double* t;                // [M]
    double* dt;               // [M]
    long long* steps;         // [M]
    long long* iters;         // [M] (LiteImplicit)
    int* status;              // [M]
    unsigned long long* counter; // work-stealing cursor
    // stratigraphy / parameters
    const int* cell_layer;    // [N]
    const int* fc_kind;       // [n_layers] kind | solver<<8
    const double *por, *sat, *org;                               // [n_layers*M]
    const double *vg_alpha, *vg_n, *theta_res, *Tm, *beta, *omega; // [n_layers*M]
    const int* table_map;     // [n_layers*M] or nullptr
    const double* lutH; const double* lutT; const double* lutS;   // knots H, T and segment slopes dT/dH (concatenated, +Inf padded tables)
    const unsigned* lutB;     // per table nb buckets: first candidate segment | trips << 24 (concatenated)
    const int* lut_off; const int* lut_len; const int* lut_extrap; const int* lut_boff; const int* lut_nb;
    // thermal properties
    double sk_w, sk_i, sk_a, sk_m, sk_o;   // sqrt(kh_*)
    double ch_w, ch_i, ch_a, ch_m, ch_o, L;
    double g, Lsl, R, rho_w;
    // grid (device arrays)
    const double* zc;         // [N]   cell midpoints
    const double* dk;         // [N]   cell thickness
    const double* inv_dk;     // [N]
    const double* dxc;        // [N-1] midpoint distances
    const double* eW1;        // [N+1] weight of the upper cell (Δk[e-1]), interior edges
    const double* eW2;        // [N+1] weight of the lower cell (Δk[e])
    const double* eG;         // [N+1] (w1+w2)/Δxc[e-1]
    double ctop, cbot;        // 1/(Δk[0]/2), 1/(Δk[N-1]/2)
    // boundary conditions
    int top_kind, bot_kind, use_nfactor, limiter;
    Forcing top, bot;
    const double* top_inv_dt; // [top.n-1] 1/(t[k+1]-t[k]) of the top forcing table (or nullptr)
    const double *nf, *nt, *top_offset, *bot_value;   // [M] (bot_value may be nullptr -> forcing)
    double maxdelta, courant, dtmin, dtmax;
    int lite_miniters, lite_maxiters; double lite_tol;
    int bot_layer_first;      // first cell of the bottom layer (heat_implicit.jl:83 quirk)
    int tform;                // HeatBalance(:T): the state is T, dT = dH / dHdT (heat_conduction.jl:147-152, soil_heat.jl:100-113)
};

// ---- SFCC fast path (cgb_euler_sfcc_fast.cuh): tables of the unique soil class
constexpr int SF_CURVE_DEG = 5;        // polynomial degree per bin
constexpr int SF_CURVE_STRIDE = 6;     // doubles per bin = coefficients (48 B rows: 8 consecutive rows are bank-conflict free for LDS.128)
constexpr int SF_BIN_BITS = 6;         // 2^6 bins per octave of u = T* - TK
constexpr int SF_BIN_MASK = ((1 << SF_BIN_BITS) - 1) << (20 - SF_BIN_BITS);
constexpr int SF_BIN_HALF = 1 << (19 - SF_BIN_BITS);

// Everything the kernel needs about the (unique) soil class; passed by value -> constant bank.
struct SfccFast {
    // presolver interpolant
    double boff, invbw;        // bucket = (unsigned)(H * invbw + boff) with boff = -Hmin * invbw, saturated to [0, nb-1]
    int nk, nkp, nb, nbins;    // knots, padded knot-array length (three arrays H, slope, T of nkp doubles each), buckets, curve rows
    // freeze curve
    double tstarK;             // T* [K]
    double thw_thawed;         // θw for T >= T*  (= θ(ψ0)); also row 0 of the curve table
    int idx_base;              // (hi word >> 16) of the lower edge of curve row 1, minus 1 (row 0 = thawed constant)
    int _pad;
    double fit_error;          // measured max relative error of the curve table against the closed form (host, long double)
    double kb, dK;             // sqrt-sum conductivity at θw = 0, sk_w - sk_i (thermal_properties.jl:37)
    int n2, _pad1;             // van Genuchten n == 2: closed form through one rsqrt (no curve table in the blob)
    double kappa, x0, dth, thres; // n == 2: x = kappa max(u, 0) + x0, θw = thres + dth (1 + x²)^-1/2
    // shared-memory blob (copied verbatim from `blob`): byte offsets, all multiples of 16
    const double* blob;
    int blob_bytes, off_seg, off_bucket, off_curve;
};

// heat+salt extras (Physics/Salt)
struct SaltDev {
    const double* por_cell;   // porosity profile (SedimentCompactionInitializer): [N] shared by all columns, or [N*M] (por_per_column)
    int por_per_column, _pad;
    const double* benthic;    // [M] SaltGradient.benthicSalt
    const double *sat, *org, *alpha, *n, *theta_res, *Tm;   // [M] soil / DallAmicoSalt parameters of every column
    double para_por;
    double tau, ds0, dc_max, courant, heat_courant;
    int surface_state;
};

// Optional diagnostic outputs of one RHS evaluation (device pointers, any may be null)
struct Diag {
    double *T, *thw, *C, *dHdT, *dthwdT, *kc;   // [N*M]
    double *k, *jH;                             // [(N+1)*M]
    double *dH;                                 // [N*M]
    double *an, *as, *ap, *bp;                  // [N*M]
    double *dT, *Hd;                            // [N*M] temperature form: dT = dH / dHdT; H = T C + L θw (diagnostic enthalpy)
};

// ------------------------------------------------------------------------------------------------
// math helpers
// ------------------------------------------------------------------------------------------------

// Reciprocal from MUFU.RCP64H + two Newton steps (<= ~1 ulp for normal inputs; no slow path).
__device__ __forceinline__ double rcp_fast(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// One Newton step on MUFU.RCP64H: relative error ~1e-13 -- for Newton corrections and predictors, where the quotient is a small
// increment of the result.
__device__ __forceinline__ double rcp_fast1(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return fma(y, fma(-a, y, 1.0), y);
}

// Cubic (Halley-type) refinement: y1 = y0 (1 + e + e^2), e = 1 - a y0  -- 3 FMAs, error e0^3.
__device__ __forceinline__ double rcp_fast3(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    double p = fma(e, e, e);
    return fma(y, p, y);
}

// 1/sqrt(a) from MUFU.RSQ64H + two Newton steps (a > 0, normal range).
__device__ __forceinline__ double rsqrt_fast(double a)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double h = 0.5 * y;
    double e = fma(-a * y, y, 1.0);
    y = fma(h, e, y);
    h = 0.5 * y;
    e = fma(-a * y, y, 1.0);
    return fma(h, e, y);
}

__device__ __forceinline__ double rcp_raw(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
}

// min/max by compare+select (3 instructions; fmin/fmax cost ~7 on sm_100 because of their NaN-quieting sequence).
// NaN in `a` is dropped (comparison false), like fmin/fmax with a non-NaN second operand.
__device__ __forceinline__ double sel_max(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double sel_min(double a, double b) { return (a < b) ? a : b; }
// max(x, 0) on the integer pipe: clear everything when the sign bit is set (-0.0 -> +0.0, NaN with sign -> 0)
__device__ __forceinline__ double relu_bits(double x)
{
    int hi = __double2hiint(x), lo = __double2loint(x);
    int m = ~(hi >> 31);
    return __hiloint2double(hi & m, lo & m);
}

// Taylor coefficients 1/13! ... 1/2!, 1, 1 of exp and 1/21, 1/19, ... 1/3 of 2 atanh(s)/s - 2 (Horner order)
static __constant__ double CGB_EXP_C[14] = {1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
                                            2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.3888888888888889e-03,
                                            8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5, 1.0, 1.0};
static __constant__ double CGB_LOG_C[10] = {1.0 / 21.0, 1.0 / 19.0, 1.0 / 17.0, 1.0 / 15.0, 1.0 / 13.0, 1.0 / 11.0, 1.0 / 9.0, 1.0 / 7.0, 1.0 / 5.0, 1.0 / 3.0};

// exp(a) without the special-case handling of the CUDA math library (80 SASS instructions -> ~25): a is clamped to
// [-708, 709] (results stay normal, so the power of two can be added straight into the exponent field), reduced by
// k = rint(a log2 e) with a two-term ln 2, and evaluated with the degree-13 Taylor polynomial on |r| <= 0.3466
// (truncation 4e-18).  Measured against libm in tests/test_gpu_sfcc.py: <= 1 ulp over [-700, 700].
__device__ __forceinline__ double exp_fast(double a)
{
    a = sel_min(sel_max(a, -708.0), 709.0);
    const double MAGIC = 6755399441055744.0;             // 1.5 * 2^52: round-to-nearest-integer by addition
    double t = fma(a, 1.4426950408889634074, MAGIC);
    int k = __double2loint(t);
    double kd = t - MAGIC;
    double r = fma(kd, -6.93147180559945286227e-01, a);
    r = fma(kd, -2.31904681384629955842e-17, r);
    // coefficients from constant memory (c[bank][offset] operands): as FP64 literals each one costs two UMOVs per use
    double p = CGB_EXP_C[0];
#pragma unroll
    for (int i = 1; i < 14; ++i) p = fma(p, r, CGB_EXP_C[i]);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// log(x) for normal x > 0 (callers clamp to >= 1e-300): x = 2^e m with m in [sqrt(1/2), sqrt 2), s = (m-1)/(m+1),
// log m = 2 atanh s = 2s (1 + z/3 + z^2/5 + ... + z^10/21), z = s^2 <= 0.0295 (truncation 6e-19).  ~35 instructions
// against 104 (log) / 256 (log1p) for the library versions; <= 2 ulp (tests/test_gpu_sfcc.py).
__device__ __forceinline__ double log_pos(double x)
{
    int hi = __double2hiint(x), lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    int mh = (hi & 0x000fffff) | 0x3ff00000;
    bool big = mh > 0x3ff6a09e;                          // m > sqrt 2
    mh = big ? mh - 0x00100000 : mh;
    e = big ? e + 1 : e;
    double m = __hiloint2double(mh, lo);
    double f = m - 1.0;                                  // exact
    double s = f * rcp_fast(m + 1.0);
    double z = s * s;
    double p = CGB_LOG_C[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) p = fma(p, z, CGB_LOG_C[i]);
    double s2 = s + s;
    double lm = fma(s2 * z, p, s2);
    double ed = (double)e;
    return fma(ed, 6.93147180559945286227e-01, fma(ed, 2.31904681384629955842e-17, lm));
}


__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// max over the warp of a NON-NEGATIVE double through the integer REDUX unit
__device__ __forceinline__ double warp_max_nonneg(double v)
{
    // v >= 0 (or NaN-free by construction: fmax drops NaN): order of doubles == order of their bit patterns
    unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    unsigned mhi = __reduce_max_sync(FULL, hi);
    unsigned mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
    return __hiloint2double((int)mhi, (int)mlo);
}

// min over the warp of a POSITIVE double (or +inf): order of the bit patterns == order of the values
__device__ __forceinline__ double warp_min_pos(double v)
{
    unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    unsigned mhi = __reduce_min_sync(FULL, hi);
    unsigned mlo = __reduce_min_sync(FULL, hi == mhi ? lo : 0xffffffffu);
    return __hiloint2double((int)mhi, (int)mlo);
}

// Gridded(Linear) evaluation: i = searchsortedlast clamped, w = (x-x_i)/(x_{i+1}-x_i), (1-w) y_i + w y_{i+1}
__device__ __forceinline__ double lerp_table(const double* __restrict__ xk, const double* __restrict__ yk, int nk, double x)
{
    int lo = 0, hi = nk - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(xk + mid) <= x) lo = mid; else hi = mid;
    }
    double x0 = __ldg(xk + lo), x1 = __ldg(xk + lo + 1);
    double w = (x - x0) / (x1 - x0);
    return (1.0 - w) * __ldg(yk + lo) + w * __ldg(yk + lo + 1);
}

// Presolver interpolant through the bucket index.  The host sizes the uniform H-grid so that almost every bucket holds
// at most one knot: one bucket load gives the first candidate segment lo and the number of bisection trips over
// [lo, lo + 2^trips - 1] (0 or 1 almost everywhere), then T = T_i + (H - H_i) * slope_i.  Linear extrapolation uses the
// end segments; flat extrapolation clamps H to the knot range first.
__device__ __forceinline__ double lut_eval_bucket(const double* __restrict__ Hk, const double* __restrict__ Tk, const double* __restrict__ Sk,
                                                  const unsigned* __restrict__ B, double hmin, double hmax, double invbw, double nbm1,
                                                  int extrap, double H)
{
    if (extrap == 0) H = sel_min(sel_max(H, hmin), hmax);
    double fb = sel_min(sel_max((H - hmin) * invbw, 0.0), nbm1);
    unsigned w = __ldg(B + (int)fb);
    int lo = (int)(w & 0xffffffu), hi = lo + (1 << (w >> 24)) - 1;
    for (int k = (int)(w >> 24); k > 0; --k) {
        int mid = (lo + hi + 1) >> 1;
        bool p = __ldg(Hk + mid) <= H;
        lo = p ? mid : lo;
        hi = p ? hi : mid - 1;
    }
    return fma(H - __ldg(Hk + lo), __ldg(Sk + lo), __ldg(Tk + lo));
}

// ---- table-driven exp / log for the van Genuchten powers of the STEPPING kernels ------------------------------------------
// 64-entry tables in (static) shared memory, filled by vgtab_init() at kernel start: {1/c_i, log c_i} for c_i = 1 + (i + 1/2)/64
// and 2^(j/64).  log x = e ln2 + log c_i + log1p(m/c_i - 1) with |m/c_i - 1| <= 2^-7 (degree-6 series), exp a = 2^(k/64) e^r with
// |r| <= ln2/128 (degree-5 series): ~17 and ~14 instructions against ~35 and ~25 of log_pos / exp_fast, <= 3 ulp
// (cgb_selftest_explog, tests/test_gpu_sfcc.py).  The diagnostics kernels keep the series versions.
__shared__ double g_vgtab[192];

__device__ __forceinline__ void vgtab_init()
{
    for (int i = threadIdx.x; i < 64; i += blockDim.x) {
        const double c = 1.0 + (i + 0.5) / 64.0;
        g_vgtab[2 * i] = 1.0 / c;
        g_vgtab[2 * i + 1] = log(c);
        g_vgtab[128 + i] = exp2(i / 64.0);
    }
    __syncthreads();
}

// constants of exp_tab / log_tab in constant memory (c[bank][offset] operands): as FP64 literals every use costs two moves
static __constant__ double CGB_TAB_C[18] = {
    1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0,                       // 0..4   e^r - 1 series
    -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5,                            // 5..9   log1p series
    92.332482616893656877,                                              // 10     64 / ln 2
    -6.93147180369123816490e-01 / 64.0, -1.90821492927058770002e-10 / 64.0,   // 11, 12  -ln2_hi / 64 (33 significant bits), -ln2_lo / 64
    6.93147180369123816490e-01, 1.90821492927058770002e-10,             // 13, 14 ln2_hi, ln2_lo
    6755399441055744.0, -708.0, 709.0};                                 // 15 (1.5 * 2^52), 16, 17

__device__ __forceinline__ double exp_tab(double a)
{
    a = sel_min(sel_max(a, CGB_TAB_C[16]), CGB_TAB_C[17]);
    const double t = fma(a, CGB_TAB_C[10], CGB_TAB_C[15]);              // round-to-nearest-integer by adding 1.5 * 2^52
    const int k = __double2loint(t);
    const double kd = t - CGB_TAB_C[15];
    double r = fma(kd, CGB_TAB_C[11], a);                               // kd * (ln2_hi / 64) is exact
    r = fma(kd, CGB_TAB_C[12], r);
    double p = fma(r, CGB_TAB_C[0], CGB_TAB_C[1]);
    p = fma(p, r, CGB_TAB_C[2]);
    p = fma(p, r, CGB_TAB_C[3]);
    p = fma(p, r, CGB_TAB_C[4]);
    p = p * r;                                            // e^r - 1, |r| <= ln2 / 128 (truncation r^6/720 = 3.5e-17)
    const double T = g_vgtab[128 + (k & 63)];
    const double res = fma(T, p, T);
    return __hiloint2double(__double2hiint(res) + ((k >> 6) << 20), __double2loint(res));
}

__device__ __forceinline__ double log_tab(double x)      // normal x > 0
{
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int idx = (hi >> 14) & 63;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double2 t = *reinterpret_cast<const double2*>(&g_vgtab[2 * idx]);     // 1/c, log c
    const double r = fma(m, t.x, -CGB_TAB_C[4]);          // |r| <= 2^-7
    double p = fma(r, CGB_TAB_C[5], CGB_TAB_C[6]);
    p = fma(p, r, CGB_TAB_C[7]);
    p = fma(p, r, CGB_TAB_C[8]);
    p = fma(p, r, CGB_TAB_C[9]);
    p = p * r;
    const double l1 = fma(p, r, r);                       // log1p(r), truncation r^7/7 = 2.5e-16
    const double ed = (double)e;
    return fma(ed, CGB_TAB_C[13], fma(ed, CGB_TAB_C[14], t.y + l1));
}

// x^n and (1+x^n)^-m for x > 0 through exp / log (~150 instructions with the series versions, ~75 with the tables).  Inlined
// everywhere: with the library exp/log/log1p (~520 instructions) this had to stay out of line (the unrolled callers overflowed the
// instruction cache); with the lean versions inlining measured +9 % on the Newton kernel and +5 % on heat+salt.
template <bool TAB = false>
__device__ __forceinline__ void vg_powers_body(double x, double n, double m, double& xn, double& pb)
{
    if (TAB) {
        xn = exp_tab(n * log_tab(x));
        pb = exp_tab(-m * log_tab(1.0 + xn));
    } else {
        xn = exp_fast(n * log_pos(x));
        pb = exp_fast(-m * log_pos(1.0 + xn));   // 1 + x^n rounds with absolute error 1.1e-16: relative 1e-16 m in pb
    }
}
template <bool TAB = false>
__device__ __forceinline__ double2 vg_powers(double x, double n, double m)
{
    double2 r;
    vg_powers_body<TAB>(x, n, m, r.x, r.y);
    return r;
}

// van Genuchten θ(ψ) and dθ/dψ for ψ <= 0; θsat for ψ > 0.  x^n and (1+x^n)^-m go through exp/log (x > 0, base >= 1:
// no special cases needed), about half the cost of two pow() calls; error ~ (n |log x| + 2) ulp.
template <bool TAB = false>
__device__ __forceinline__ double vg_theta(double psi, double thsat, double thres, double alpha, double n, double m, double& dth)
{
    if (psi <= 0.0) {
        double x = fabs(-alpha * psi);
        double d = thsat - thres;
        if (n == 2.0) {
            // m = 1/2: θ = θres + Δθ (1+x²)^(-1/2), dθ/dψ = Δθ α x (1+x²)^(-3/2)  -- no transcendental at all
            double rs = rsqrt_fast(fma(x, x, 1.0));
            dth = d * alpha * x * (rs * rs * rs);
            return fma(d, rs, thres);
        }
        if (x > 0.0) {
            double2 r = vg_powers<TAB>(sel_max(x, 1e-300), n, m);
            double xn = r.x, pb = r.y;
            double base = 1.0 + xn;
            dth = d * m * n * alpha * (xn * rcp_fast(x)) * (pb * rcp_fast(base));
            return fma(d, pb, thres);
        }
        dth = 0.0;
        return thres + d;
    }
    dth = 0.0;
    return thsat;
}

__device__ __forceinline__ double vg_psi(double th, double thsat, double thres, double alpha, double n)
{
    double m = 1.0 - 1.0 / n;
    if (th < thsat) {
        double Se = (th - thres) / (thsat - thres);
        return -1.0 / alpha * pow(pow(Se, -1.0 / m) - 1.0, 1.0 / n);
    }
    return 0.0;
}

// SFCC θw(T) and dθw/dT from the staged layer constants (DallAmico / PainterKarra).
template <bool TAB = false>
__device__ __forceinline__ double sfcc_theta(const double* lc, double T, double& dthdT)
{
    double TK = T + 273.15;
    double Tstar = lc[LC_TSTAR];
    double psi, dpsi;
    if (Tstar - TK >= 0.0) { psi = lc[LC_PSI0] + lc[LC_CFAC] * (TK - Tstar); dpsi = lc[LC_CFAC]; }
    else { psi = lc[LC_PSI0]; dpsi = 0.0; }
    double dth;
    double thtot = lc[LC_THWI];
    double thsat = fmax(thtot, lc[LC_POR]);
    double thw = vg_theta<TAB>(psi, thsat, lc[LC_THRES], lc[LC_ALPHA], lc[LC_N], lc[LC_M], dth);
    dthdT = dth * dpsi;
    return thw;
}

} // namespace cgb
