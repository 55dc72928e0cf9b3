// cgb_tu_presolver.cu -- cgb_build_sfcc_tables: presolver tables of every distinct soil class, built on the device
// (cgb_presolver.cuh).  Compiled with -fmad=false: the knots should agree with the host builders (the CPU checker's table builder and
// hostmath.presolver_table) to rounding, and those do not contract a*b+c.
#include "cgb_handle.h"
#include "cgb_presolver.cuh"

#include <array>
using namespace cgb;

extern "C" int cgb_build_sfcc_tables(cgb_handle* h, double Tmin, double errtol, double Tmax, int32_t extrap, int32_t* n_tables_out)
{
    if (!h) return CGB_ERR_INVALID;
    if (!(Tmax > Tmin) || !(errtol > 0.0)) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_build_sfcc_tables: need Tmax > Tmin and errtol > 0");
    if (h->cfg.physics != CGB_PHYSICS_HEAT) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_build_sfcc_tables: enthalpy-form heat physics only");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const size_t M = (size_t)h->M, nl = (size_t)h->nl;
    auto value = [&](const char* name, size_t l, size_t c) {
        const ParamArray& p = h->params[name];
        switch (p.per) {
        case CGB_PER_LAYER: return p.v[l];
        case CGB_PER_COLUMN: return p.v[c];
        case CGB_PER_COLUMN_LAYER: return p.v[l * M + c];
        default: return p.v[0];
        }
    };
    // deduplicate the (layer, column) pairs that use an SFCC with the LUT solver into parameter classes
    using Key = std::array<double, 10>;
    std::map<Key, int> ids;
    std::vector<PresolverClass> classes;
    std::vector<int> tmap(nl * M, 0);
    for (size_t l = 0; l < nl; ++l) {
        const int kind = h->fc_kind[l];
        if (kind == CGB_FC_FREEWATER || h->fc_solver[l] != CGB_SOLVER_LUT) continue;
        if (kind != CGB_FC_DALLAMICO && kind != CGB_FC_PAINTERKARRA) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_build_sfcc_tables: layer %zu: DallAmico / PainterKarra only", l);
        // columns share a class unless one of the parameters varies per column
        bool per_col = false;
        for (const char* n : {"por", "sat", "org", "vg_alpha", "vg_n", "theta_res", "Tm", "beta", "omega"}) {
            int per = h->params[n].per;
            per_col = per_col || per == CGB_PER_COLUMN || per == CGB_PER_COLUMN_LAYER;
        }
        for (size_t c = 0; c < (per_col ? M : 1); ++c) {
            const double por = value("por", l, c), sat = value("sat", l, c);
            PresolverClass pc{};
            pc.kind = kind; pc.alpha = value("vg_alpha", l, c); pc.n = value("vg_n", l, c); pc.theta_res = value("theta_res", l, c);
            pc.Tm = value("Tm", l, c); pc.beta = (kind == CGB_FC_PAINTERKARRA) ? value("beta", l, c) : 1.0;
            pc.omega = (kind == CGB_FC_PAINTERKARRA) ? value("omega", l, c) : 1.0;
            pc.por = por; pc.sat = (por * sat) / por; pc.org = value("org", l, c);       // sat = θwi / por (soil_heat.jl:126)
            Key k{(double)kind, pc.alpha, pc.n, pc.theta_res, pc.Tm, pc.beta, pc.omega, pc.por, pc.sat, pc.org};
            auto it = ids.find(k);
            int id;
            if (it == ids.end()) { id = (int)classes.size(); ids.emplace(k, id); classes.push_back(pc); }
            else id = it->second;
            if (per_col) tmap[l * M + c] = id;
            else for (size_t cc = 0; cc < M; ++cc) tmap[l * M + cc] = id;
        }
    }
    if (classes.empty()) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_build_sfcc_tables: no layer uses an SFCC with the LUT solver");
    if (classes.size() > 65536) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_build_sfcc_tables: %zu distinct soil classes (limit 65536): use CGB_SOLVER_NEWTON for "
                                         "fully per-column parameters", classes.size());
    auto G = [&](const char* n) { return h->params[n].v[0]; };
    PresolverConsts K{};
    K.ch_w = G("ch_w"); K.ch_i = G("ch_i"); K.ch_a = G("ch_a"); K.ch_m = G("ch_m"); K.ch_o = G("ch_o"); K.L = G("L");
    K.g = 9.80665; K.Lsl = 3.34e5; K.Tmin = Tmin; K.errtol = errtol; K.Tmax = Tmax;
    const int nc = (int)classes.size();
    int cap = 4096;
    std::vector<double> Hk, Tk;
    std::vector<int> nk(nc);
    for (;;) {
        K.cap = cap;
        PresolverClass* d_cls = nullptr; double *d_H = nullptr, *d_T = nullptr; int* d_nk = nullptr;
        cudaError_t e = cudaMalloc((void**)&d_cls, nc * sizeof(PresolverClass));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_H, (size_t)nc * cap * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_T, (size_t)nc * cap * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc((void**)&d_nk, nc * sizeof(int));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_cls, classes.data(), nc * sizeof(PresolverClass), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) {
            k_presolver_build<<<(nc + 63) / 64, 64, 0, h->stream>>>(d_cls, nc, K, d_H, d_T, d_nk);
            h->launches++;
            e = cudaGetLastError();
        }
        Hk.resize((size_t)nc * cap); Tk.resize((size_t)nc * cap);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nk.data(), d_nk, nc * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(Hk.data(), d_H, (size_t)nc * cap * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(Tk.data(), d_T, (size_t)nc * cap * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
        cudaError_t e2 = cudaStreamSynchronize(h->stream);
        cudaFree(d_cls); cudaFree(d_H); cudaFree(d_T); cudaFree(d_nk);
        if (e != cudaSuccess || e2 != cudaSuccess) CGB_FAIL(h, CGB_ERR_CUDA, "cgb_build_sfcc_tables: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
        bool overflow = false;
        for (int i = 0; i < nc; ++i) overflow = overflow || nk[i] < 0;
        if (!overflow) break;
        if (cap >= (1 << 18)) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_build_sfcc_tables: a table needs more than %d knots", cap);
        cap *= 4;
    }
    h->tables.clear();
    for (int i = 0; i < nc; ++i) {
        if (nk[i] < 2) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_build_sfcc_tables: class %d produced %d knots", i, nk[i]);
        LutTable& t = h->tables[i];
        t.H.assign(Hk.begin() + (size_t)i * cap, Hk.begin() + (size_t)i * cap + nk[i]);
        t.T.assign(Tk.begin() + (size_t)i * cap, Tk.begin() + (size_t)i * cap + nk[i]);
        t.extrap = extrap;
    }
    h->table_map = tmap;
    h->dirty = true;
    if (n_tables_out) *n_tables_out = nc;
    return CGB_OK;
}

// read back one table (after cgb_build_sfcc_tables or cgb_set_sfcc_table): *nk = its length; Hk / Tk (capacity cap) may be NULL
extern "C" int cgb_get_sfcc_table(cgb_handle* h, int32_t table_id, double* Hk, double* Tk, int64_t cap, int64_t* nk)
{
    if (!h || !nk) return CGB_ERR_INVALID;
    auto it = h->tables.find(table_id);
    if (it == h->tables.end()) CGB_FAIL(h, CGB_ERR_INVALID, "no SFCC table %d", table_id);
    *nk = (int64_t)it->second.H.size();
    if (Hk && Tk) {
        if (cap < *nk) CGB_FAIL(h, CGB_ERR_INVALID, "table %d has %lld knots, capacity %lld", table_id, (long long)*nk, (long long)cap);
        std::memcpy(Hk, it->second.H.data(), *nk * sizeof(double));
        std::memcpy(Tk, it->second.T.data(), *nk * sizeof(double));
    }
    return CGB_OK;
}
