// cgb_tu_sfcc_fast.cu -- host side of the SFCC single-class fast path (cgb_euler_sfcc_fast.cuh): eligibility, the
// shared-memory blob (presolver segments + bucket index + curve table) and the launcher.
#include "cgb_launch.h"
#include "cgb_euler_sfcc_fast.cuh"
using namespace cgb;

namespace {

// θw of the closed form at u = T* - TK (long double): ψ = ψ0 - cfac u, θ = θres + (θsat - θres)(1 + (α|ψ|)^n)^-(1-1/n)
struct CurveClass {
    long double psi0, cfac, alpha, n, m, thsat, thres;
    long double operator()(long double u) const
    {
        long double psi = psi0 - cfac * u;
        if (psi > 0.0L) return thsat;
        long double x = fabsl(-alpha * psi);
        if (x == 0.0L) return thsat;
        return thres + (thsat - thres) * powl(1.0L + powl(x, n), -m);
    }
};

// degree-SF_CURVE_DEG interpolating polynomial in r at the Chebyshev nodes of [-h, h], h = 2^-(SF_BIN_BITS+1) (Newton form in
// t = r / h, expanded to monomials of r)
void fit_bin(const CurveClass& f, int E, int k, double* coef)
{
    constexpr int NC = SF_CURVE_DEG + 1, NB = 1 << SF_BIN_BITS;
    const long double PI_L = 3.14159265358979323846264338327950288L;
    long double t[NC], y[NC];
    const long double ck = 1.0L + (k + 0.5L) / NB, scale = ldexpl(1.0L, E), h = 0.5L / NB;
    for (int j = 0; j < NC; ++j) {
        t[j] = cosl((2 * j + 1) * PI_L / (2.0L * NC));
        y[j] = f(scale * (ck + t[j] * h));
    }
    long double dd[NC];
    for (int j = 0; j < NC; ++j) dd[j] = y[j];
    for (int lev = 1; lev < NC; ++lev)
        for (int j = NC - 1; j >= lev; --j) dd[j] = (dd[j] - dd[j - 1]) / (t[j] - t[j - lev]);
    long double mono[NC] = {0}, basis[NC + 1] = {0};
    basis[0] = 1.0L;
    int deg = 0;
    for (int j = 0; j < NC; ++j) {
        for (int i = 0; i <= deg; ++i) mono[i] += dd[j] * basis[i];
        for (int i = deg + 1; i >= 1; --i) basis[i] = basis[i - 1] - t[j] * basis[i];
        basis[0] = -t[j] * basis[0];
        ++deg;
    }
    long double s = 1.0L;
    for (int j = 0; j < NC; ++j) { coef[j] = (double)(mono[j] * s); s /= h; }
}

bool all_equal(const std::vector<double>& v)
{
    for (double x : v)
        if (x != v[0]) return false;
    return true;
}

} // namespace

// Decides whether the handle qualifies for k_heat_sfcc_fast and builds / uploads the blob.  Called at the end of finalize().
int cgb_sfcc_fast_prepare(cgb_handle* h)
{
    h->sfcc_fast_ok = false;
    const bool disabled = getenv("CGB_SFCC_FAST") && atoi(getenv("CGB_SFCC_FAST")) == 0;   // re-read per finalize: tests switch in-process
    if (disabled) return CGB_OK;
    if (h->cfg.physics != CGB_PHYSICS_HEAT || h->cfg.stepper != CGB_STEPPER_EULER || h->cfg.limiter != CGB_LIMITER_MAXDELTA) return CGB_OK;
    if (h->cfg.bot_kind != CGB_BC_NEUMANN || !(h->params.count("bot_value") || h->forc[1].kind == 0)) return CGB_OK;
    const int kind = h->fc_kind[0];
    if (kind != CGB_FC_DALLAMICO && kind != CGB_FC_PAINTERKARRA) return CGB_OK;
    for (int l = 0; l < h->nl; ++l)
        if (h->fc_kind[l] != kind || h->fc_solver[l] != CGB_SOLVER_LUT) return CGB_OK;
    // one soil class: every (layer, column) shares the parameters and the presolver table
    for (const char* name : {"por", "sat", "org", "vg_alpha", "vg_n", "theta_res", "Tm", "beta", "omega"})
        if (!all_equal(h->params[name].v)) return CGB_OK;
    int table_id = 0;
    if (!h->table_map.empty()) {
        table_id = h->table_map[0];
        for (int id : h->table_map)
            if (id != table_id) return CGB_OK;
    } else if (h->nl != 1) return CGB_OK;
    auto it = h->tables.find(table_id);
    if (it == h->tables.end()) return CGB_OK;
    if (it->second.H.size() < 2 || it->second.H.size() > 4000) return CGB_OK;
    // Flat extrapolation (T = T_first below the first knot, T_last above the last) is built INTO the table: a sentinel knot far
    // below the first one and a zero slope on both end records, so the kernel needs no clamp.  knot0 = index of the first real knot.
    const bool flat = (it->second.extrap == CGB_EXTRAP_FLAT);
    std::vector<double> Hk, Tk;
    if (flat) { Hk.push_back(-1e300); Tk.push_back(it->second.T.front()); }
    Hk.insert(Hk.end(), it->second.H.begin(), it->second.H.end()); Tk.insert(Tk.end(), it->second.T.begin(), it->second.T.end());
    const int nk = (int)Hk.size(), knot0 = flat ? 1 : 0;

    auto G = [&](const char* n) { return h->params[n].v[0]; };
    const double por = G("por"), sat = G("sat"), org = G("org"), alpha = G("vg_alpha"), n = G("vg_n"), thres = G("theta_res");
    double beta = 1.0, omega = 1.0;
    if (kind == CGB_FC_PAINTERKARRA) { beta = G("beta"); omega = G("omega"); }
    // the same derived constants as layer_consts (cgb_euler.cuh)
    const double thwi = por * sat, thm = (1.0 - por) * (1.0 - org), tho = (1.0 - por) * org, tha = 1.0 - thwi - thm - tho;
    const double sat_e = thwi / por, thtot_s = sat_e * por, thsat_e = std::max(thtot_s, por), m = 1.0 - 1.0 / n;
    double psi0 = 0.0;
    if (thtot_s < thsat_e) {
        double Se = (thtot_s - thres) / (thsat_e - thres);
        psi0 = -1.0 / alpha * std::pow(std::pow(Se, -1.0 / m) - 1.0, 1.0 / n);
    }
    const double gconst = 9.80665, Lsl = 3.34e5;
    const double TmK = G("Tm") + 273.15, gT = omega * gconst / Lsl * psi0, Tstar = TmK + gT * TmK, cfac = beta * Lsl / (gconst * Tstar);
    CurveClass f{(long double)psi0, (long double)cfac, (long double)alpha, (long double)n, (long double)m, (long double)thsat_e, (long double)thres};
    SfccFast S{};
    S.tstarK = Tstar;
    {
        double x = std::fabs(-alpha * psi0);
        S.thw_thawed = (psi0 <= 0.0) ? (x > 0.0 ? thres + (thsat_e - thres) * std::pow(1.0 + std::pow(x, n), -m) : thres + (thsat_e - thres)) : thsat_e;
    }
    const double sk_i = std::sqrt(G("kh_i")), sk_a = std::sqrt(G("kh_a")), sk_m = std::sqrt(G("kh_m")), sk_o = std::sqrt(G("kh_o"));
    S.kb = sk_i * thwi + sk_a * tha + sk_m * thm + sk_o * tho;
    S.dK = std::sqrt(G("kh_w")) - sk_i;
    // van Genuchten n == 2: closed form with one rsqrt, no table
    S.n2 = (n == 2.0);
    S.kappa = alpha * cfac; S.x0 = std::fabs(-alpha * psi0); S.dth = thsat_e - thres; S.thres = thres;
    // curve table range: below 2^Emin the curve equals its thawed value to < 1e-14 relative (row 0); above 2^10 K nothing exists
    constexpr int NB = 1 << SF_BIN_BITS;
    int Emin = -70;
    for (int E = -70; E <= 8; ++E) {
        long double d = fabsl(f(ldexpl(1.0L, E)) - (long double)S.thw_thawed);
        if (d > 1e-14L * (long double)S.thw_thawed) { Emin = E - 1; break; }
        Emin = E;
    }
    const int Emax = 9;                                   // bins cover [2^Emin, 2^(Emax+1))
    S.nbins = S.n2 ? 0 : (Emax + 1 - Emin) * NB + 1;      // + row 0: the thawed constant
    S.idx_base = ((Emin + 1023) << SF_BIN_BITS) - 1;
    // presolver buckets: as many as fit (power of two), same construction as the general path (cgb_api.cu finalize)
    const int nkp = (nk + 16 + 1) & ~1;                                               // + Inf padding (2^tmax <= 16 entries), even
    const size_t seg_bytes = (size_t)nkp * 3 * sizeof(double);
    const size_t curve_bytes = (size_t)S.nbins * SF_CURVE_STRIDE * sizeof(double);
    const size_t budget = 224 * 1024;
    if (seg_bytes + curve_bytes + 2 * 1024 > budget) return CGB_OK;
    int nb = 1024;
    while (nb < 16384 && seg_bytes + curve_bytes + (size_t)(2 * nb) * 2 <= budget && nb < 16 * nk) nb <<= 1;
    std::vector<unsigned short> bucket(nb);
    const double hmin = Hk[knot0], range = Hk[nk - 1] - Hk[knot0], bw = range / nb, eps = 1e-12 * range;
    auto last_le = [&](double x) { return (int)(std::upper_bound(Hk.begin(), Hk.end(), x) - Hk.begin()) - 1; };
    int tmax = 0;
    for (int b = 0; b < nb; ++b) {
        int lo = std::min(std::max(last_le(hmin + b * bw - eps), 0), nk - 1);
        int hi = std::max(std::min(last_le(hmin + (b + 1) * bw + eps), nk - 1), lo);
        int trips = 0;
        while ((1 << trips) < hi - lo + 1) ++trips;
        tmax = std::max(tmax, trips);
        bucket[b] = (unsigned short)(lo | (trips << 12));
    }
    if (tmax > 4) return CGB_OK;                          // > 16 knots in one bucket: leave it to the general kernel
    std::vector<double> blob;
    auto align16 = [&]() { while ((blob.size() * sizeof(double)) % 16) blob.push_back(0.0); };
    S.off_seg = 0;
    blob.assign((size_t)nkp * 3, 0.0);
    for (int i = 0; i < nkp; ++i) {
        if (i < nk) {
            int sg = std::min(i, nk - 2);
            double slope = (Tk[sg + 1] - Tk[sg]) / (Hk[sg + 1] - Hk[sg]);
            if (flat && (i == 0 || i == nk - 1)) slope = 0.0;
            blob[i] = Hk[i]; blob[nkp + i] = slope; blob[2 * nkp + i] = Tk[i];
        } else blob[i] = INFINITY;
    }
    align16();
    S.off_bucket = (int)(blob.size() * sizeof(double));
    blob.resize(blob.size() + (size_t)nb * 2 / sizeof(double));
    std::memcpy(reinterpret_cast<unsigned char*>(blob.data()) + S.off_bucket, bucket.data(), (size_t)nb * 2);
    align16();
    S.off_curve = (int)(blob.size() * sizeof(double));
    long double worst = 0.0L;
    if (!S.n2) {
        blob.push_back(S.thw_thawed);
        for (int j = 1; j < SF_CURVE_STRIDE; ++j) blob.push_back(0.0);
    }
    for (int b = 0; b < S.nbins - 1; ++b) {
        int E = Emin + (b >> SF_BIN_BITS), k = b & (NB - 1);
        double c[SF_CURVE_STRIDE] = {0};
        fit_bin(f, E, k, c);
        for (int j = 0; j < SF_CURVE_STRIDE; ++j) blob.push_back(c[j]);
        // verify the fit between the Chebyshev nodes and at the bin ends, evaluated in double exactly as the kernel does
        for (int q = 0; q <= 16; ++q) {
            double r = (q - 8) / (16.0 * NB), r2 = r * r;
            double pe = std::fma(c[4], r2, c[2]), po = std::fma(c[5], r2, c[3]);
            pe = std::fma(pe, r2, c[0]); po = std::fma(po, r2, c[1]);
            double p = std::fma(po, r, pe);
            long double ref = f(ldexpl(1.0L, E) * (1.0L + (k + 0.5L) / NB + (long double)r));
            worst = std::max(worst, fabsl((long double)p - ref) / ref);
        }
    }
    // steep curves (van Genuchten n well above 3) exceed what a degree-7 fit on 1/16-octave bins resolves: general kernel then
    if (!(worst <= 1e-13L)) return CGB_OK;
    align16();
    S.blob_bytes = (int)(blob.size() * sizeof(double));
    S.invbw = nb / range; S.boff = -hmin * S.invbw;
    S.nk = nk; S.nkp = nkp; S.nb = nb;
    S.fit_error = (double)worst;
    double* d_blob = nullptr;
    int rc = dev_upload(h, &d_blob, blob);
    if (rc) return rc;
    S.blob = d_blob;
    h->sf = S;
    h->sfcc_fast_ok = true;
    return CGB_OK;
}

#ifndef CGB_SF_THREADS
#define CGB_SF_THREADS 384
#endif
constexpr int kSfThreads = CGB_SF_THREADS;

template <int CPL, bool SAVE, bool N2>
static int launch_sf_s(cgb_handle* h, const FastTargets& tg, double* saveT, double* saveW)
{
    auto kern = k_heat_sfcc_fast<CPL, kSfThreads, SAVE, N2>;
    size_t smem = (size_t)h->sf.blob_bytes;
    int blocks_per_sm = 0;
    if (smem > 48 * 1024) CUDA_TRY(h, cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    (void)blocks_per_sm;
    const int wpb = kSfThreads / 32;
    long long grid = std::min<long long>((h->M + wpb - 1) / wpb, (long long)h->num_sms);
    CUDA_TRY(h, cudaMemsetAsync(h->dev.counter, 0, sizeof(unsigned long long), h->stream));
    kern<<<(unsigned)grid, kSfThreads, smem, h->stream>>>(h->dev, h->sf, tg, saveT, saveW);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return CGB_OK;
}

template <int CPL>
static int launch_sf_t(cgb_handle* h, const FastTargets& tg, double* saveT, double* saveW)
{
    if (h->sf.n2) {
        if (saveT || saveW) return launch_sf_s<CPL, true, true>(h, tg, saveT, saveW);
        return launch_sf_s<CPL, false, true>(h, tg, nullptr, nullptr);
    }
    if (saveT || saveW) return launch_sf_s<CPL, true, false>(h, tg, saveT, saveW);
    return launch_sf_s<CPL, false, false>(h, tg, nullptr, nullptr);
}

int cgb_launch_sfcc_fast_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride)
{
    if (n < 1 || n > FAST_MAX_TARGETS) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_launch_sfcc_fast_multi: 1..%d targets", FAST_MAX_TARGETS);
    FastTargets tg{};
    for (int k = 0; k < n; ++k) tg.t[k] = targets[k];
    tg.n = n; tg.save_stride = save_stride;
    int cpl = (h->N + 31) / 32;
    switch (cpl) {
    case 1: return launch_sf_t<1>(h, tg, saveT, saveW);
    case 2: return launch_sf_t<2>(h, tg, saveT, saveW);
    case 3: case 4: return launch_sf_t<4>(h, tg, saveT, saveW);
    case 5: case 6: return launch_sf_t<6>(h, tg, saveT, saveW);
    case 7: return launch_sf_t<7>(h, tg, saveT, saveW);
    case 8: return launch_sf_t<8>(h, tg, saveT, saveW);
    case 9: return launch_sf_t<9>(h, tg, saveT, saveW);
    case 10: return launch_sf_t<10>(h, tg, saveT, saveW);
    case 11: case 12: return launch_sf_t<12>(h, tg, saveT, saveW);
    default: return launch_sf_t<16>(h, tg, saveT, saveW);
    }
}

int cgb_sfcc_fast_selftest(cgb_handle* h, const double* T, long long n, double* out)
{
    if (!h->sfcc_fast_ok) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "the handle does not qualify for the SFCC single-class fast path");
    double* d = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&d, 2 * n * sizeof(double)));
    cudaMemcpyAsync(d, T, n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    size_t smem = (size_t)h->sf.blob_bytes;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute((const void*)k_selftest_sfcc_curve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        k_selftest_sfcc_curve<<<64, 256, smem, h->stream>>>(h->sf, d, n, d + n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + n, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess || e2 != cudaSuccess) CGB_FAIL(h, CGB_ERR_CUDA, "cgb_selftest_sfcc_curve: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return CGB_OK;
}
