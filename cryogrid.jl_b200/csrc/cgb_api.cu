// cgb_api.cu -- C ABI of libcgb200.so (see include/cgb200.h). Host-side bookkeeping + kernel launches.
// There is no CPU path here: without a CUDA device cgb_create fails with CGB_ERR_NO_DEVICE.
#include "cgb_handle.h"
#include "cgb_launch.h"
#include "cgb_init.cuh"

using namespace cgb;

static thread_local std::string g_create_error;

// ------------------------------------------------------------------------------------------------
extern "C" int cgb_abi_version(void) { return CGB_ABI_VERSION; }

extern "C" const char* cgb_last_error(const cgb_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int cgb_create(const cgb_config* cfg, cgb_handle** out)
{
    if (!cfg || !out) { g_create_error = "null argument"; return CGB_ERR_INVALID; }
    *out = nullptr;
    if (cfg->abi_version != CGB_ABI_VERSION) { g_create_error = "ABI version mismatch"; return CGB_ERR_INVALID; }
    if (cfg->n_cells < 2 || cfg->n_columns < 1 || cfg->n_layers < 1 || cfg->n_layers > cfg->n_cells) {
        g_create_error = "need n_cells >= 2, n_columns >= 1, 1 <= n_layers <= n_cells (n_layers == n_cells: Heterogeneous per-cell parameters)";
        return CGB_ERR_INVALID;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device available (libcgb200 has no CPU fallback): ") + cudaGetErrorString(e);
        return CGB_ERR_NO_DEVICE;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_create_error = "device ordinal out of range"; return CGB_ERR_INVALID; }
    e = cudaSetDevice(cfg->device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return CGB_ERR_CUDA; }
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major != 10) {
        g_create_error = "libcgb200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor);
        return CGB_ERR_NO_DEVICE;
    }
    int cpl_needed = (cfg->n_cells + 31) / 32;
    if (cpl_needed > 16) { g_create_error = "n_cells > 512 is not supported by the warp-per-column kernels"; return CGB_ERR_UNSUPPORTED; }
    cgb_handle* h = new cgb_handle();
    h->cfg = *cfg;
    h->N = cfg->n_cells; h->M = cfg->n_columns; h->nl = cfg->n_layers;
    h->nvar = (cfg->physics == CGB_PHYSICS_HEAT_SALT) ? 2 : 1;
    h->num_sms = prop.multiProcessorCount;
    h->fc_kind.assign(h->nl, CGB_FC_FREEWATER);
    h->fc_solver.assign(h->nl, CGB_SOLVER_LUT);
    h->cell_layer.assign(h->N, 0);
    cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&h->stage_ready[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->stage_free[i], cudaEventDisableTiming);
    }
    // defaults: thermal_properties.jl:13-18, Soils/para/simple.jl:37-43, Physics.jl:3-8
    auto setg = [&](const char* n, double v) { h->params[n].v = {v}; h->params[n].per = CGB_PER_GLOBAL; };
    setg("kh_w", 0.57); setg("kh_i", 2.2); setg("kh_a", 0.025); setg("kh_m", 3.8); setg("kh_o", 0.25);
    setg("ch_w", 4.2e6); setg("ch_i", 1.9e6); setg("ch_a", 0.00125e6); setg("ch_m", 2.0e6); setg("ch_o", 2.5e6);
    setg("L", 3.34e8);
    setg("por", 0.5); setg("sat", 1.0); setg("org", 0.0);
    setg("vg_alpha", 1.0); setg("vg_n", 2.0); setg("theta_res", 0.0); setg("Tm", 0.0); setg("beta", 1.0); setg("omega", 1.0);
    setg("nf", 1.0); setg("nt", 1.0); setg("top_offset", 0.0);
    setg("tau", 1.5); setg("ds0", 8.0e-10); setg("benthic_salt", 900.0); setg("salt_dc_max", 5.0e-3); setg("salt_para_por", 0.5);
    setg("salt_surface_state", 0.0);
    // device state
    size_t nu = (size_t)h->nvar * h->N * h->M;
    int rc = 0;
    rc |= dev_alloc(h, &h->dev.u, nu);
    rc |= dev_alloc(h, &h->dev.t, (size_t)h->M);
    rc |= dev_alloc(h, &h->dev.dt, (size_t)h->M);
    rc |= dev_alloc(h, &h->dev.steps, (size_t)h->M);
    rc |= dev_alloc(h, &h->dev.iters, (size_t)h->M);
    rc |= dev_alloc(h, &h->dev.status, (size_t)h->M);
    rc |= dev_alloc(h, &h->dev.counter, 1);
    rc |= dev_alloc(h, &h->d_scratch, (size_t)cgb_handle::kScratch * (h->N + 1) * h->M);
    if (rc) { g_create_error = h->err; cgb_destroy(h); return rc; }
    h->n_state_allocs = h->allocs.size();
    *out = h;
    return CGB_OK;
}

extern "C" void cgb_destroy(cgb_handle* h)
{
    if (!h) return;
    clear_stale_error("cgb_destroy(entry)");
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (void* p : h->allocs) cudaFree(p);
    for (int i = 0; i < 2; ++i) {
        if (h->d_stage[i]) cudaFree(h->d_stage[i]);
        if (i == 0 && h->d_lite) cudaFree(h->d_lite);
        if (i == 0 && h->d_selfull) cudaFree(h->d_selfull);
        if (h->stage_ready[i]) cudaEventDestroy(h->stage_ready[i]);
        if (h->stage_free[i]) cudaEventDestroy(h->stage_free[i]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    clear_stale_error("cgb_destroy(exit)");
    delete h;
}

extern "C" int cgb_set_grid(cgb_handle* h, const double* edges)
{
    if (!h || !edges) return CGB_ERR_INVALID;
    for (int i = 0; i < h->N; ++i)
        if (!(edges[i + 1] > edges[i])) CGB_FAIL(h, CGB_ERR_INVALID, "grid edges must be strictly increasing (edge %d)", i);
    h->edges.assign(edges, edges + h->N + 1);
    h->dirty = true;
    return CGB_OK;
}

extern "C" int cgb_set_layers(cgb_handle* h, const int32_t* cell_layer)
{
    if (!h || !cell_layer) return CGB_ERR_INVALID;
    for (int i = 0; i < h->N; ++i) {
        if (cell_layer[i] < 0 || cell_layer[i] >= h->nl) CGB_FAIL(h, CGB_ERR_INVALID, "cell_layer[%d]=%d out of range", i, cell_layer[i]);
        if (i > 0 && cell_layer[i] < cell_layer[i - 1]) CGB_FAIL(h, CGB_ERR_INVALID, "cell_layer must be non-decreasing");
    }
    h->cell_layer.assign(cell_layer, cell_layer + h->N);
    h->dirty = true;
    return CGB_OK;
}

// Granularities finalize() actually honours, per parameter (a granularity outside this table used to be accepted and then read
// as v[0], i.e. a per-column ensemble silently ran with member 0's value).
static unsigned allowed_granularities(const cgb_handle* h, const char* name)
{
    auto bit = [](int per) { return 1u << per; };
    const unsigned G = bit(CGB_PER_GLOBAL), L = bit(CGB_PER_LAYER), C = bit(CGB_PER_COLUMN), CL = bit(CGB_PER_COLUMN_LAYER), CELL = bit(CGB_PER_CELL);
    const unsigned CC = bit(CGB_PER_CELL_COLUMN);
    const bool salt = h->cfg.physics == CGB_PHYSICS_HEAT_SALT;
    auto is = [&](const char* a) { return std::strcmp(name, a) == 0; };
    if (is("por")) return salt ? (G | CELL | CC) : (G | L | C | CL);     // heat+salt: porosity profile of the single SalineGround layer
    for (const char* n : {"sat", "org", "vg_alpha", "vg_n", "theta_res", "Tm"})
        if (is(n)) return salt ? (G | C) : (G | L | C | CL);
    for (const char* n : {"beta", "omega"})
        if (is(n)) return salt ? G : (G | L | C | CL);
    for (const char* n : {"nf", "nt", "top_offset", "bot_value", "benthic_salt"})
        if (is(n)) return G | C;
    return G;   // thermal properties (kh_*, ch_*, L) and salt properties: one value for the whole batch
}

extern "C" int cgb_set_param(cgb_handle* h, const char* name, const double* values, int64_t n, int32_t per)
{
    if (!h || !name || !values) return CGB_ERR_INVALID;
    auto it = h->params.find(name);
    bool is_bot = std::strcmp(name, "bot_value") == 0;
    if (it == h->params.end() && !is_bot) CGB_FAIL(h, CGB_ERR_INVALID, "unknown parameter '%s'", name);
    int64_t want = -1;
    switch (per) {
    case CGB_PER_GLOBAL: want = 1; break;
    case CGB_PER_LAYER: want = h->nl; break;
    case CGB_PER_COLUMN: want = h->M; break;
    case CGB_PER_COLUMN_LAYER: want = (int64_t)h->nl * h->M; break;
    case CGB_PER_CELL: want = h->N; break;
    case CGB_PER_CELL_COLUMN: want = (int64_t)h->N * h->M; break;
    default: CGB_FAIL(h, CGB_ERR_INVALID, "bad granularity %d", per);
    }
    if (!(allowed_granularities(h, name) & (1u << per)))
        CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "parameter '%s' does not support granularity %d (GLOBAL=0, LAYER=1, COLUMN=2, COLUMN_LAYER=3, CELL=4, CELL_COLUMN=5) "
                 "with this physics; see include/cgb200.h", name, per);
    if (n != want) CGB_FAIL(h, CGB_ERR_INVALID, "parameter '%s': expected %lld values, got %lld", name, (long long)want, (long long)n);
    ParamArray& p = h->params[name];
    p.v.assign(values, values + n);
    p.per = per;
    h->dirty = true;
    return CGB_OK;
}

extern "C" int cgb_set_freezecurve(cgb_handle* h, int32_t layer, int32_t fc_kind, int32_t solver)
{
    if (!h) return CGB_ERR_INVALID;
    if (layer < 0 || layer >= h->nl) CGB_FAIL(h, CGB_ERR_INVALID, "layer %d out of range", layer);
    if (fc_kind < 0 || fc_kind > 3 || solver < 0 || solver > 1) CGB_FAIL(h, CGB_ERR_INVALID, "bad freeze curve / solver");
    h->fc_kind[layer] = fc_kind;
    h->fc_solver[layer] = solver;
    h->dirty = true;
    return CGB_OK;
}

extern "C" int cgb_set_sfcc_table(cgb_handle* h, int32_t table_id, const double* Hk, const double* Tk, int64_t nk, int32_t extrap)
{
    if (!h || !Hk || !Tk) return CGB_ERR_INVALID;
    if (table_id < 0 || table_id >= 65536 || nk < 2) CGB_FAIL(h, CGB_ERR_INVALID, "bad table id or length");
    for (int64_t i = 1; i < nk; ++i)
        if (!(Hk[i] > Hk[i - 1])) CGB_FAIL(h, CGB_ERR_INVALID, "table knots must be strictly increasing");
    LutTable& t = h->tables[table_id];
    t.H.assign(Hk, Hk + nk);
    t.T.assign(Tk, Tk + nk);
    t.extrap = extrap;
    h->dirty = true;
    return CGB_OK;
}

extern "C" int cgb_set_sfcc_table_map(cgb_handle* h, const int32_t* map)
{
    if (!h) return CGB_ERR_INVALID;
    if (map) h->table_map.assign(map, map + (size_t)h->nl * h->M); else h->table_map.clear();
    h->dirty = true;
    return CGB_OK;
}

extern "C" int cgb_set_forcing_table(cgb_handle* h, int32_t which, const double* t, const double* y, int64_t n)
{
    if (!h || !t || !y || which < 0 || which > 1) return CGB_ERR_INVALID;
    if (n < 2) CGB_FAIL(h, CGB_ERR_INVALID, "forcing table needs >= 2 knots");
    for (int64_t i = 1; i < n; ++i)
        if (!(t[i] > t[i - 1])) CGB_FAIL(h, CGB_ERR_INVALID, "forcing times must be strictly increasing");
    auto& f = h->forc[which];
    f.kind = 2; f.t.assign(t, t + n); f.y.assign(y, y + n);
    h->dirty = true;
    return CGB_OK;
}

extern "C" int cgb_set_forcing_periodic(cgb_handle* h, int32_t which, double period, double amplitude, double phase, double level)
{
    if (!h || which < 0 || which > 1) return CGB_ERR_INVALID;
    auto& f = h->forc[which];
    f.kind = 1; f.period = period; f.amp = amplitude; f.phase = phase; f.level = level;
    h->dirty = true;
    return CGB_OK;
}

extern "C" int cgb_set_forcing_constant(cgb_handle* h, int32_t which, double value)
{
    if (!h || which < 0 || which > 1) return CGB_ERR_INVALID;
    auto& f = h->forc[which];
    f.kind = 0; f.value = value;
    h->dirty = true;
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// push the host-side description to the device
// ------------------------------------------------------------------------------------------------
static std::vector<double> expand(const cgb_handle* h, const ParamArray& p, int target /* COLUMN or COLUMN_LAYER */)
{
    size_t M = (size_t)h->M, nl = (size_t)h->nl;
    std::vector<double> out(target == CGB_PER_COLUMN ? M : nl * M);
    if (target == CGB_PER_COLUMN) {
        for (size_t c = 0; c < M; ++c) out[c] = (p.per == CGB_PER_COLUMN) ? p.v[c] : p.v[0];
    } else {
        for (size_t l = 0; l < nl; ++l)
            for (size_t c = 0; c < M; ++c) {
                double v;
                switch (p.per) {
                case CGB_PER_LAYER: v = p.v[l]; break;
                case CGB_PER_COLUMN: v = p.v[c]; break;
                case CGB_PER_COLUMN_LAYER: v = p.v[l * M + c]; break;
                default: v = p.v[0];
                }
                out[l * M + c] = v;
            }
    }
    return out;
}

static int finalize(cgb_handle* h)
{
    if (!h->dirty) return CGB_OK;
    if ((int)h->edges.size() != h->N + 1) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_set_grid has not been called");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    while (h->allocs.size() > h->n_state_allocs) { cudaFree(h->allocs.back()); h->allocs.pop_back(); }
    Dev& d = h->dev;
    const int N = h->N;
    d.N = N; d.n_layers = h->nl; d.nvar = h->nvar; d.M = h->M;
    int rc = 0;
    // grid: Numerics/grid.jl:24-33
    std::vector<double> zc(N), dk(N), idk(N), dxc(std::max(N - 1, 1)), w1(N + 1, 1.0), w2(N + 1, 1.0), g(N + 1, 0.0);
    for (int i = 0; i < N; ++i) { zc[i] = (h->edges[i] + h->edges[i + 1]) / 2.0; dk[i] = h->edges[i + 1] - h->edges[i]; idk[i] = 1.0 / dk[i]; }
    for (int i = 0; i + 1 < N; ++i) dxc[i] = zc[i + 1] - zc[i];
    for (int e = 1; e < N; ++e) { w1[e] = dk[e - 1]; w2[e] = dk[e]; g[e] = (w1[e] + w2[e]) / dxc[e - 1]; }
    double *p_dk, *p_idk, *p_dxc, *p_w1, *p_w2, *p_g, *p_zc;
    rc |= dev_upload(h, &p_dk, dk); rc |= dev_upload(h, &p_idk, idk); rc |= dev_upload(h, &p_dxc, dxc);
    rc |= dev_upload(h, &p_w1, w1); rc |= dev_upload(h, &p_w2, w2); rc |= dev_upload(h, &p_g, g); rc |= dev_upload(h, &p_zc, zc);
    d.zc = p_zc; d.dk = p_dk; d.inv_dk = p_idk; d.dxc = p_dxc; d.eW1 = p_w1; d.eW2 = p_w2; d.eG = p_g;
    d.ctop = 1.0 / (dk[0] / 2.0); d.cbot = 1.0 / (dk[N - 1] / 2.0);
    int* p_cl; rc |= dev_upload(h, &p_cl, h->cell_layer); d.cell_layer = p_cl;
    int first = N - 1;
    while (first > 0 && h->cell_layer[first - 1] == h->cell_layer[N - 1]) --first;
    d.bot_layer_first = first;
    // freeze curves
    std::vector<int> kinds(h->nl);
    bool any_lut = false;
    for (int l = 0; l < h->nl; ++l) {
        kinds[l] = h->fc_kind[l] | (h->fc_solver[l] << 8);
        if (h->fc_kind[l] != CGB_FC_FREEWATER && h->fc_solver[l] == CGB_SOLVER_LUT && h->cfg.physics != CGB_PHYSICS_HEAT_T) any_lut = true;
    }
    int* p_k; rc |= dev_upload(h, &p_k, kinds); d.fc_kind = p_k;
    // LUTs
    int maxid = -1;
    for (auto& kv : h->tables) maxid = std::max(maxid, kv.first);
    std::vector<int> off(std::max(maxid + 1, 1), 0), len(std::max(maxid + 1, 1), 0), ext(std::max(maxid + 1, 1), 0);
    for (auto& kv : h->tables) { len[kv.first] = (int)kv.second.H.size(); ext[kv.first] = kv.second.extrap; }
    if (any_lut) {
        for (int l = 0; l < h->nl; ++l) {
            if (h->fc_kind[l] == CGB_FC_FREEWATER || h->fc_solver[l] != CGB_SOLVER_LUT) continue;
            if (h->table_map.empty()) {
                if (l > maxid || len[l] < 2) CGB_FAIL(h, CGB_ERR_INVALID, "layer %d uses the SFCC LUT solver but table %d was not set", l, l);
            } else {
                for (long long c = 0; c < h->M; ++c) {
                    int id = h->table_map[(size_t)l * h->M + c];
                    if (id < 0 || id > maxid || len[id] < 2) CGB_FAIL(h, CGB_ERR_INVALID, "table_map[%d,%lld]=%d refers to a missing table", l, c, id);
                }
            }
        }
    }
    // knots (H, T, segment slope dT/dH as three arrays: neighbouring knots share 32-byte sectors, which is what the
    // scattered per-lane loads of the kernels want) and the bucket acceleration index of every table (see
    // lut_eval_bucket).  The uniform H-grid has nb = 2^k ~ 8x the knot count buckets (so that a bucket rarely holds more
    // than one knot), bounded by LUT_NB_MAX and by a total index size of 256 MB when there are very many tables.  A
    // bucket is one word: first candidate segment | trips << 24, the candidates being [lo, lo + 2^trips - 1]; every
    // table is padded with +Inf knots so that this range never leaves it.
    std::vector<double> allH, allT, allS;
    std::vector<unsigned> allB;
    std::vector<int> boff(off.size(), 0), nbs(off.size(), LUT_NB_MIN);
    size_t nb_cap = LUT_NB_MAX;
    while (nb_cap > (size_t)LUT_NB_MIN && nb_cap * sizeof(unsigned) * std::max<size_t>(h->tables.size(), 1) > ((size_t)256 << 20)) nb_cap >>= 1;
    static const int nb_factor = getenv("CGB_LUT_NB_FACTOR") ? atoi(getenv("CGB_LUT_NB_FACTOR")) : 8;   // tuning knob
    for (auto& kv : h->tables) {
        const std::vector<double>& H = kv.second.H; const std::vector<double>& T = kv.second.T;
        const int nk = (int)H.size();
        if (nk >= (1 << 24)) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "SFCC table %d has %d knots (limit 2^24-1)", kv.first, nk);
        int nb = LUT_NB_MIN;
        while (nb < nb_factor * nk && nb < (int)nb_cap) nb <<= 1;
        nbs[kv.first] = nb;
        boff[kv.first] = (int)allB.size();
        // the device computes the bucket as (int)((H - Hmin) * (nb / range)): three roundings, i.e. a slop of ~1e-15 range
        // against the boundaries Hmin + b bw used here; widen every bucket by 1e-12 range
        double hmin = H[0], range = H[nk - 1] - H[0], bw = range / nb, eps = 1e-12 * range;
        auto last_le = [&](double x) { return (int)(std::upper_bound(H.begin(), H.end(), x) - H.begin()) - 1; };
        int tmax = 0;
        for (int b = 0; b < nb; ++b) {
            int lo = std::min(std::max(last_le(hmin + b * bw - eps), 0), nk - 1);
            int hi = std::max(std::min(last_le(hmin + (b + 1) * bw + eps), nk - 1), lo);
            int trips = 0;
            while ((1 << trips) < hi - lo + 1) ++trips;
            tmax = std::max(tmax, trips);
            allB.push_back((unsigned)lo | ((unsigned)trips << 24));
        }
        off[kv.first] = (int)allH.size();
        for (int i = 0; i < nk; ++i) {
            int sgm = std::min(i, nk - 2);
            allH.push_back(H[i]); allT.push_back(T[i]);
            allS.push_back((T[sgm + 1] - T[sgm]) / (H[sgm + 1] - H[sgm]));          // knot nk-1: the last segment continued
        }
        for (int i = 0; i < (1 << tmax); ++i) { allH.push_back(INFINITY); allT.push_back(0.0); allS.push_back(0.0); }
    }
    double *p_lh, *p_lt, *p_ls; unsigned* p_lb; int *p_off, *p_len, *p_ext, *p_boff, *p_nb, *p_map = nullptr;
    rc |= dev_upload(h, &p_lh, allH); rc |= dev_upload(h, &p_lt, allT); rc |= dev_upload(h, &p_ls, allS); rc |= dev_upload(h, &p_lb, allB);
    rc |= dev_upload(h, &p_off, off); rc |= dev_upload(h, &p_len, len); rc |= dev_upload(h, &p_ext, ext);
    rc |= dev_upload(h, &p_boff, boff); rc |= dev_upload(h, &p_nb, nbs);
    if (!h->table_map.empty()) rc |= dev_upload(h, &p_map, h->table_map);
    d.lutH = p_lh; d.lutT = p_lt; d.lutS = p_ls; d.lutB = p_lb; d.lut_off = p_off; d.lut_len = p_len; d.lut_extrap = p_ext;
    d.lut_boff = p_boff; d.lut_nb = p_nb; d.table_map = p_map;
    // per (layer, column) parameters
    auto up_cl = [&](const char* name, const double** dst) {
        double* p; std::vector<double> v = expand(h, h->params[name], CGB_PER_COLUMN_LAYER);
        rc |= dev_upload(h, &p, v); *dst = p;
    };
    auto up_c = [&](const char* name, const double** dst) {
        double* p; std::vector<double> v = expand(h, h->params[name], CGB_PER_COLUMN);
        rc |= dev_upload(h, &p, v); *dst = p;
    };
    up_cl("por", &d.por); up_cl("sat", &d.sat); up_cl("org", &d.org);
    {   // the fast FreeWater kernel needs θwi >= 1e-8 everywhere (FreezeCurves.freewater floors θtot there) and a
        // bottom boundary value that is constant in time
        std::vector<double> por = expand(h, h->params["por"], CGB_PER_COLUMN_LAYER), sat = expand(h, h->params["sat"], CGB_PER_COLUMN_LAYER);
        bool ok = true;
        for (size_t i = 0; i < por.size() && ok; ++i) ok = (por[i] * sat[i] >= 1e-8);
        h->fast_ok = ok && (h->params.count("bot_value") || h->forc[1].kind == 0) && h->cfg.bot_kind == CGB_BC_NEUMANN;
    }
    up_cl("vg_alpha", &d.vg_alpha); up_cl("vg_n", &d.vg_n); up_cl("theta_res", &d.theta_res);
    up_cl("Tm", &d.Tm); up_cl("beta", &d.beta); up_cl("omega", &d.omega);
    up_c("nf", &d.nf); up_c("nt", &d.nt); up_c("top_offset", &d.top_offset);
    d.bot_value = nullptr;
    if (h->params.count("bot_value")) up_c("bot_value", &d.bot_value);
    auto G = [&](const char* n) { return h->params[n].v[0]; };
    d.sk_w = std::sqrt(G("kh_w")); d.sk_i = std::sqrt(G("kh_i")); d.sk_a = std::sqrt(G("kh_a"));
    d.sk_m = std::sqrt(G("kh_m")); d.sk_o = std::sqrt(G("kh_o"));
    d.ch_w = G("ch_w"); d.ch_i = G("ch_i"); d.ch_a = G("ch_a"); d.ch_m = G("ch_m"); d.ch_o = G("ch_o");
    d.L = G("L");
    d.g = 9.80665; d.Lsl = 3.34e5; d.R = 8.314459; d.rho_w = 1000.0;   // Physics.jl:3-8 ; FreezeCurves defaults
    // boundary conditions
    d.top_kind = h->cfg.top_kind; d.bot_kind = h->cfg.bot_kind; d.use_nfactor = h->cfg.use_nfactor; d.limiter = h->cfg.limiter;
    for (int w = 0; w < 2; ++w) {
        Forcing F{};
        auto& f = h->forc[w];
        F.kind = f.kind; F.value = f.value; F.period = f.period; F.amplitude = f.amp; F.phase = f.phase; F.level = f.level;
        F.n = (int)f.t.size(); F.t = nullptr; F.y = nullptr;
        if (f.kind == 2) {
            double *pt, *py;
            rc |= dev_upload(h, &pt, f.t); rc |= dev_upload(h, &py, f.y);
            F.t = pt; F.y = py;
            F.inv_step = (double)(f.t.size() - 1) / (f.t.back() - f.t.front());
            if (w == 0) {
                std::vector<double> inv(f.t.size() - 1);
                for (size_t k = 0; k + 1 < f.t.size(); ++k) inv[k] = 1.0 / (f.t[k + 1] - f.t[k]);
                double* pi; rc |= dev_upload(h, &pi, inv); d.top_inv_dt = pi;
            }
        } else if (w == 0) d.top_inv_dt = nullptr;
        (w == 0 ? d.top : d.bot) = F;
    }
    d.tform = (h->cfg.physics == CGB_PHYSICS_HEAT_T);
    if (d.tform) {
        if (h->cfg.stepper != CGB_STEPPER_EULER || h->cfg.limiter != CGB_LIMITER_CFL)
            CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "HeatBalance(:T) runs with CGEuler and the CFL limiter (heat_types.jl:17)");
        for (int l = 0; l < h->nl; ++l)
            if (h->fc_kind[l] != CGB_FC_DALLAMICO && h->fc_kind[l] != CGB_FC_PAINTERKARRA)
                CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "HeatBalance(:T): layer %d needs an SFCC (FreeWater has no temperature form, soil_heat.jl:100-113)", l);
    }
    d.maxdelta = h->cfg.maxdelta; d.courant = h->cfg.courant; d.dtmin = h->cfg.dtmin; d.dtmax = h->cfg.dtmax;
    d.lite_miniters = h->cfg.lite_miniters; d.lite_maxiters = h->cfg.lite_maxiters; d.lite_tol = h->cfg.lite_tol;
    if (rc) return rc;
    // heat+salt extras
    rc = cgb_salt_finalize(h, zc);
    if (rc) return rc;
    rc = cgb_sfcc_fast_prepare(h);           // SFCC single-class fast path: eligibility + shared-memory tables
    if (rc) return rc;
    rc = cgb_newton_tab_prepare(h);          // per-column curve tables of the Newton kernel: HBM cache
    if (rc) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->dirty = false;
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------
__global__ void k_fill_state(double* t, double* dt, long long* steps, long long* iters, int* status, long long M, double t0, double dt0)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < M) { t[i] = t0; dt[i] = dt0; steps[i] = 0; iters[i] = 0; status[i] = 0; }
}

extern "C" int cgb_set_state(cgb_handle* h, const double* u, double t0)
{
    clear_stale_error("cgb_set_state");
    if (!h || !u) return CGB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    size_t nu = (size_t)h->nvar * h->N * h->M;
    CUDA_TRY(h, cudaMemcpyAsync(h->dev.u, u, nu * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    int thr = 256;
    k_fill_state<<<(unsigned)((h->M + thr - 1) / thr), thr, 0, h->stream>>>(h->dev.t, h->dev.dt, h->dev.steps, h->dev.iters, h->dev.status, h->M, t0, h->cfg.dt0);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_state = true;
    h->t_front = t0;
    return CGB_OK;
}

extern "C" int cgb_restore_state(cgb_handle* h, const double* u, const double* t, const double* dt, const int64_t* steps)
{
    clear_stale_error("cgb_restore_state");
    if (!h || !u || !t || !dt) return CGB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    const size_t M = (size_t)h->M, nu = (size_t)h->nvar * h->N * M;
    double tmax = t[0];
    for (size_t c = 0; c < M; ++c) {
        if (!(dt[c] > 0.0) || !std::isfinite(t[c])) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_restore_state: column %zu has t = %g, dt = %g", c, t[c], dt[c]);
        tmax = std::max(tmax, t[c]);
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->dev.u, u, nu * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->dev.t, t, M * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(h->dev.dt, dt, M * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (steps) CUDA_TRY(h, cudaMemcpyAsync(h->dev.steps, steps, M * sizeof(long long), cudaMemcpyHostToDevice, h->stream));
    else CUDA_TRY(h, cudaMemsetAsync(h->dev.steps, 0, M * sizeof(long long), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->dev.iters, 0, M * sizeof(long long), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->dev.status, 0, M * sizeof(int), h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_state = true;
    h->t_front = tmax;
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// on-device initial conditions
// ------------------------------------------------------------------------------------------------
static int finish_init(cgb_handle* h, double t0)
{
    int thr = 256;
    k_fill_state<<<(unsigned)((h->M + thr - 1) / thr), thr, 0, h->stream>>>(h->dev.t, h->dev.dt, h->dev.steps, h->dev.iters, h->dev.status, h->M, t0, h->cfg.dt0);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    h->have_state = true;
    h->t_front = t0;
    return CGB_OK;
}

extern "C" int cgb_init_interp(cgb_handle* h, const double* zk, const double* vk, int64_t nk, int32_t per, double t0)
{
    if (!h || !zk || !vk || nk < 1) return CGB_ERR_INVALID;
    clear_stale_error("cgb_init_interp");
    if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_init_interp: heat physics only");
    if (per != CGB_PER_GLOBAL && per != CGB_PER_COLUMN) CGB_FAIL(h, CGB_ERR_INVALID, "profile values are GLOBAL (nk) or COLUMN (nk*M)");
    for (int64_t k = 1; k < nk; ++k)
        if (!(zk[k] > zk[k - 1])) CGB_FAIL(h, CGB_ERR_INVALID, "profile depths must be strictly increasing");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    size_t nv = (size_t)nk * (per == CGB_PER_COLUMN ? (size_t)h->M : 1);
    double *dz = nullptr, *dv = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&dz, nk * sizeof(double)));
    if (cudaMalloc((void**)&dv, nv * sizeof(double)) != cudaSuccess) { cudaFree(dz); CGB_FAIL(h, CGB_ERR_ALLOC, "profile upload"); }
    cudaMemcpyAsync(dz, zk, nk * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    cudaMemcpyAsync(dv, vk, nv * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    dim3 grid((unsigned)((h->M + 255) / 256), (unsigned)h->N);
    k_init_interp<<<grid, 256, 0, h->stream>>>(h->dev, h->dev.zc, dz, dv, (int)nk, per == CGB_PER_COLUMN);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(h->stream);
    cudaFree(dz); cudaFree(dv);
    if (e != cudaSuccess) CGB_FAIL(h, CGB_ERR_CUDA, "k_init_interp: %s", cudaGetErrorString(e));
    return finish_init(h, t0);
}

extern "C" int cgb_init_steadystate(cgb_handle* h, const double* T0, int32_t per, double Qgeo, int32_t maxiters, double thresh, double t0)
{
    if (!h || !T0) return CGB_ERR_INVALID;
    clear_stale_error("cgb_init_steadystate");
    if (h->cfg.physics != CGB_PHYSICS_HEAT) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_init_steadystate: heat physics only");
    if (per != CGB_PER_GLOBAL && per != CGB_PER_COLUMN) CGB_FAIL(h, CGB_ERR_INVALID, "T0 is GLOBAL (1) or COLUMN (M)");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    size_t n0 = per == CGB_PER_COLUMN ? (size_t)h->M : 1;
    double* d0 = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&d0, n0 * sizeof(double)));
    cudaMemcpyAsync(d0, T0, n0 * sizeof(double), cudaMemcpyHostToDevice, h->stream);
    double* Ts = h->d_scratch;                                   // two of the (N+1)*M scratch slots
    double* kcs = h->d_scratch + (size_t)(h->N + 1) * h->M;
    k_init_steadystate<<<(unsigned)((h->M + 127) / 128), 128, 0, h->stream>>>(h->dev, h->dev.zc, d0, per == CGB_PER_COLUMN, Qgeo, maxiters, thresh, Ts, kcs);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    cudaStreamSynchronize(h->stream);
    cudaFree(d0);
    if (e != cudaSuccess) CGB_FAIL(h, CGB_ERR_CUDA, "k_init_steadystate: %s", cudaGetErrorString(e));
    return finish_init(h, t0);
}

extern "C" int cgb_synchronize(cgb_handle* h)
{
    clear_stale_error("cgb_synchronize");
    if (!h) return CGB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    return CGB_OK;
}

extern "C" int cgb_get_state(cgb_handle* h, double* u, double* t, double* dt)
{
    clear_stale_error("cgb_get_state");
    if (!h) return CGB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    size_t nu = (size_t)h->nvar * h->N * h->M;
    if (u) CUDA_TRY(h, cudaMemcpyAsync(u, h->dev.u, nu * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (t) CUDA_TRY(h, cudaMemcpyAsync(t, h->dev.t, h->M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (dt) CUDA_TRY(h, cudaMemcpyAsync(dt, h->dev.dt, h->M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return CGB_OK;
}

extern "C" int cgb_get_status(cgb_handle* h, int32_t* status, int64_t* steps)
{
    if (!h) return CGB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    if (status) CUDA_TRY(h, cudaMemcpyAsync(status, h->dev.status, h->M * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    if (steps) CUDA_TRY(h, cudaMemcpyAsync(steps, h->dev.steps, h->M * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return CGB_OK;
}

extern "C" int cgb_get_stats(cgb_handle* h, cgb_stats* out)
{
    clear_stale_error("cgb_get_stats");
    if (!h || !out) return CGB_ERR_INVALID;
    size_t M = (size_t)h->M;
    std::vector<long long> steps(M), iters(M);
    std::vector<int> status(M);
    std::vector<double> dt(M);
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaMemcpyAsync(steps.data(), h->dev.steps, M * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(iters.data(), h->dev.iters, M * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(status.data(), h->dev.status, M * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaMemcpyAsync(dt.data(), h->dev.dt, M * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    cgb_stats s{};
    s.n_columns = h->M; s.n_cells = h->N;
    s.min_steps = steps[0]; s.max_steps = steps[0]; s.min_dt = dt[0]; s.max_dt = dt[0];
    for (size_t c = 0; c < M; ++c) {
        s.column_steps += steps[c]; s.lite_iterations += iters[c]; s.status_or |= status[c];
        s.min_steps = std::min<int64_t>(s.min_steps, steps[c]); s.max_steps = std::max<int64_t>(s.max_steps, steps[c]);
        s.min_dt = std::min(s.min_dt, dt[c]); s.max_dt = std::max(s.max_dt, dt[c]);
    }
    s.cell_steps = s.column_steps * h->N;
    s.kernel_launches = h->launches;
    *out = s;
    return CGB_OK;
}

extern "C" int cgb_device_state(cgb_handle* h, void** u_dev, void** t_dev, void** dt_dev)
{
    if (!h) return CGB_ERR_INVALID;
    if (u_dev) *u_dev = h->dev.u;
    if (t_dev) *t_dev = h->dev.t;
    if (dt_dev) *dt_dev = h->dev.dt;
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel dispatch
// ------------------------------------------------------------------------------------------------
static bool all_freewater(const cgb_handle* h)
{
    for (int l = 0; l < h->nl; ++l)
        if (h->fc_kind[l] != CGB_FC_FREEWATER) return false;
    return true;
}

static int launch_advance_inner(cgb_handle* h, double t_target, double* saveT, double* saveW);

// saveT / saveW (device pointers, N*M, may be null): where the stepping kernel leaves the diagnostics the reference would
// SAVE at t_target -- the cache of the last RHS evaluation of the interval (see cgb_solve_saveat).
static int launch_advance(cgb_handle* h, double t_target, double* saveT = nullptr, double* saveW = nullptr)
{
    h->t_front = std::max(h->t_front, t_target);
    if (!h->profiling) return launch_advance_inner(h, t_target, saveT, saveW);
    cudaEvent_t a, b;
    CUDA_TRY(h, cudaEventCreate(&a));
    CUDA_TRY(h, cudaEventCreate(&b));
    CUDA_TRY(h, cudaEventRecord(a, h->stream));
    int rc = launch_advance_inner(h, t_target, saveT, saveW);
    CUDA_TRY(h, cudaEventRecord(b, h->stream));
    h->prof_events.emplace_back(a, b);
    return rc;
}

static bool fast_path(const cgb_handle* h)
{
    return h->cfg.physics == CGB_PHYSICS_HEAT && h->cfg.stepper == CGB_STEPPER_EULER && all_freewater(h) &&
           h->cfg.limiter == CGB_LIMITER_MAXDELTA && h->fast_ok;
}

static bool lite_thread_variant()
{
    // "Thomas per thread or partition/PCR per warp": the warp variant is the default; CGB_LITE_VARIANT=thread selects
    // the reference-order Thomas sweep (re-read on every launch so that tests can switch in-process)
    const char* v = getenv("CGB_LITE_VARIANT");
    return v && std::strcmp(v, "thread") == 0;
}

// solver mode of the general explicit kernel: 1 = every layer SFCC + LUT (4: and every n == 2), 2 = every layer SFCC + Newton,
// 3 = every layer FreeWater, 0 = mixed
static int euler_mode(const cgb_handle* h)
{
    if (h->cfg.physics == CGB_PHYSICS_HEAT_T) return 5;   // temperature form: θw straight from the curve, du = dH / dHdT
    bool all_lut = true, all_newton = true;
    for (int l = 0; l < h->nl; ++l) {
        bool sf = h->fc_kind[l] != CGB_FC_FREEWATER;
        all_lut &= sf && h->fc_solver[l] == CGB_SOLVER_LUT;
        all_newton &= sf && h->fc_solver[l] == CGB_SOLVER_NEWTON;
    }
    if (all_lut) {
        // mode 4: every van Genuchten n is exactly 2 (the value of the reference's SFCC example): rsqrt closed form
        bool n2 = true;
        auto it = h->params.find("vg_n");
        if (it == h->params.end() || it->second.v.empty()) n2 = false;
        else for (double v : it->second.v) n2 = n2 && (v == 2.0);
        return n2 ? 4 : 1;
    }
    return all_newton ? 2 : all_freewater(h) ? 3 : 0;
}

// the stepping kernels that carry a column through several save times per launch (device-side saveat)
static bool multi_target_path(const cgb_handle* h)
{
    if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) return false;
    if (h->cfg.stepper == CGB_STEPPER_LITE_IMPLICIT) return !lite_thread_variant();
    return true;
}

// n save times in one launch; the diagnostics saved at targets[k] go to saveT/saveW + k*save_stride
static int launch_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride)
{
    if (h->cfg.stepper == CGB_STEPPER_LITE_IMPLICIT) return cgb_launch_lite_warp_multi(h, targets, n, saveT, saveW, save_stride);
    if (fast_path(h)) return cgb_launch_fast_multi(h, targets, n, saveT, saveW, save_stride);
    if (h->sfcc_fast_ok) return cgb_launch_sfcc_fast_multi(h, targets, n, saveT, saveW, save_stride);
    if (cgb_newton_tab_ok(h)) return cgb_launch_newton_tab_multi(h, targets, n, saveT, saveW, save_stride);
    Diag D{};
    D.T = saveT; D.thw = saveW;
    return cgb_launch_euler_step_multi(h, euler_mode(h), targets, n, D, save_stride);
}

static int launch_advance_inner(cgb_handle* h, double t_target, double* saveT, double* saveW)
{
    Diag D{};
    D.T = saveT; D.thw = saveW;
    if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) return cgb_salt_launch(h, t_target, false, D, nullptr);
    if (h->cfg.stepper == CGB_STEPPER_LITE_IMPLICIT && lite_thread_variant()) return cgb_launch_lite_thread(h, t_target, false, D);
    return launch_multi(h, &t_target, 1, saveT, saveW, 0);
}

static int launch_diag(cgb_handle* h, double t, const Diag& D)
{
    if (h->cfg.stepper == CGB_STEPPER_LITE_IMPLICIT) return cgb_launch_lite_thread(h, t, true, D);
    return cgb_launch_euler(h, 0, true, t, D);
}

extern "C" int cgb_advance_to(cgb_handle* h, double t_target)
{
    clear_stale_error("cgb_advance_to");
    if (!h) return CGB_ERR_INVALID;
    if (!h->have_state) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_set_state has not been called");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    return launch_advance(h, t_target);
}

// diag variables of the heat path
static int diag_slot(cgb_handle* h, const char* var, Diag& D, double** ptr, size_t* count)
{
    size_t NM = (size_t)h->N * h->M, EM = (size_t)(h->N + 1) * h->M;
    double* s = h->d_scratch;
    *count = NM;
    std::string v(var);
    if (v == "T") D.T = s; else if (v == "thw") D.thw = s; else if (v == "C") D.C = s; else if (v == "dHdT") D.dHdT = s;
    else if (v == "dthwdT") D.dthwdT = s; else if (v == "kc") D.kc = s; else if (v == "dH") D.dH = s;
    else if (v == "dT") D.dT = s; else if (v == "Henth") D.Hd = s;
    else if (v == "k") { D.k = s; *count = EM; } else if (v == "jH") { D.jH = s; *count = EM; }
    else if (v == "DT_an") D.an = s; else if (v == "DT_as") D.as = s; else if (v == "DT_ap") D.ap = s; else if (v == "DT_bp") D.bp = s;
    else CGB_FAIL(h, CGB_ERR_INVALID, "unknown diagnostic '%s'", var);
    *ptr = s;
    return CGB_OK;
}

extern "C" int cgb_get_diag(cgb_handle* h, const char* var, double t, double* out)
{
    clear_stale_error("cgb_get_diag");
    if (!h || !var || !out) return CGB_ERR_INVALID;
    if (!h->have_state) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_set_state has not been called");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) return cgb_salt_get_diag(h, var, t, out);
    Diag D{};
    double* p; size_t cnt;
    rc = diag_slot(h, var, D, &p, &cnt);
    if (rc) return rc;
    rc = launch_diag(h, t, D);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(out, p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return CGB_OK;
}

extern "C" int cgb_rhs(cgb_handle* h, double t, double* du_out)
{
    if (!h || !du_out) return CGB_ERR_INVALID;
    if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) {
        int rc = cgb_get_diag(h, "dT", t, du_out);
        if (rc) return rc;
        return cgb_get_diag(h, "dc", t, du_out + (size_t)h->N * h->M);
    }
    if (h->cfg.stepper == CGB_STEPPER_LITE_IMPLICIT) {   // heat_implicit.jl:106: du == 0
        std::memset(du_out, 0, sizeof(double) * (size_t)h->N * h->M);
        return CGB_OK;
    }
    return cgb_get_diag(h, h->cfg.physics == CGB_PHYSICS_HEAT_T ? "dT" : "dH", t, du_out);
}

// ------------------------------------------------------------------------------------------------
// solve(prob; saveat): advance + save, D2H overlapped on a second stream
// ------------------------------------------------------------------------------------------------
// Parses "H,T,thw,T_now,thw_now" and runs one save interval: advance to t with the reference-verbatim diagnostics (cache of
// the last RHS evaluation) fused into the stepping kernel, plus a DIAG launch only for *_now variables or zero-step saves.
struct SaveVars {
    std::vector<std::string> names;
    bool T = false, W = false, Tnow = false, Wnow = false;
};

static int parse_savevars(cgb_handle* h, const char* vars, SaveVars& sv)
{
    std::string s(vars), cur;
    for (char c : s) { if (c == ',') { if (!cur.empty()) sv.names.push_back(cur); cur.clear(); } else if (c != ' ') cur.push_back(c); }
    if (!cur.empty()) sv.names.push_back(cur);
    for (auto& n : sv.names) {
        if (n == "T") sv.T = true; else if (n == "thw") sv.W = true; else if (n == "T_now") sv.Tnow = true; else if (n == "thw_now") sv.Wnow = true;
        else if (n != "H") CGB_FAIL(h, CGB_ERR_INVALID, "saved variables are H, T, thw, T_now, thw_now (got '%s')", n.c_str());
    }
    return CGB_OK;
}

// fills bufT/bufW (verbatim) and nowT/nowW (recomputed at the saved state); any pointer may be null when not needed
static int advance_and_save(cgb_handle* h, double t, const SaveVars& sv, double* bufT, double* bufW, double* nowT, double* nowW)
{
    bool moved = t > h->t_front;          // otherwise no step is taken: the cache IS the state at t (initial save, basic_solvers.jl:33-37)
    int rc = launch_advance(h, t, (sv.T && moved) ? bufT : nullptr, (sv.W && moved) ? bufW : nullptr);
    if (rc) return rc;
    Diag D{};
    if (sv.Tnow) D.T = nowT;
    if (sv.Wnow) D.thw = nowW;
    bool need = sv.Tnow || sv.Wnow;
    if (!moved) {
        if (sv.T) { if (sv.Tnow) { /* copied below */ } else D.T = bufT; need = true; }
        if (sv.W) { if (sv.Wnow) { } else D.thw = bufW; need = true; }
    }
    if (need) { rc = launch_diag(h, t, D); if (rc) return rc; }
    size_t NM = (size_t)h->N * h->M;
    if (!moved && sv.T && sv.Tnow) CUDA_TRY(h, cudaMemcpyAsync(bufT, nowT, NM * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    if (!moved && sv.W && sv.Wnow) CUDA_TRY(h, cudaMemcpyAsync(bufW, nowW, NM * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    return CGB_OK;
}

// (Re)allocates both staging buffers for `elems` doubles each.  stage_elems is reset BEFORE anything is freed and set only after both
// allocations succeeded, so a failed allocation can never leave a stale size behind (a later, smaller request would otherwise pass
// the size check and launch kernels on null buffers).
static int ensure_staging(cgb_handle* h, size_t elems)
{
    if (h->stage_elems >= elems && h->d_stage[0] && h->d_stage[1]) return CGB_OK;
    h->stage_elems = 0;
    for (int i = 0; i < 2; ++i) {
        if (h->d_stage[i]) cudaFree(h->d_stage[i]);
        h->d_stage[i] = nullptr;
    }
    for (int i = 0; i < 2; ++i) {
        cudaError_t e = cudaMalloc((void**)&h->d_stage[i], elems * sizeof(double));
        if (e != cudaSuccess) {
            cudaGetLastError();
            if (i == 1) { cudaFree(h->d_stage[0]); h->d_stage[0] = nullptr; }
            h->d_stage[i] = nullptr;
            CGB_FAIL(h, CGB_ERR_ALLOC, "staging buffer (%zu bytes): %s", elems * sizeof(double), cudaGetErrorString(e));
        }
    }
    h->stage_elems = elems;
    return CGB_OK;
}

// heat+salt: saved variables are the prognostic T, c and any cell diagnostic of the salt path (thw, C, dHdT, kc, H, dT, dc, ...), all
// evaluated AT the saved state (the reference's SalineGround example saves T, θw, k, dₛ: examples/heat_sfcc_salt_constantbc.jl:20)
static int salt_solve_saveat(cgb_handle* h, const double* tsave, int64_t nsave, const char* vars, double* out)
{
    std::vector<std::string> names;
    {
        std::string s(vars), cur;
        for (char c : s) { if (c == ',') { if (!cur.empty()) names.push_back(cur); cur.clear(); } else if (c != ' ') cur.push_back(c); }
        if (!cur.empty()) names.push_back(cur);
    }
    if (names.empty()) CGB_FAIL(h, CGB_ERR_INVALID, "no variables to save");
    for (auto& n : names)
        if (n == "k" || n == "ds" || n == "jH" || n == "jc") CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_solve_saveat (heat+salt): cell variables only, got '%s'", n.c_str());
    const size_t NM = (size_t)h->N * h->M, per_save = NM * names.size();
    int rc = ensure_staging(h, per_save);
    if (rc) return rc;
    for (int64_t s = 0; s < nsave; ++s) {
        const int b = (int)(s & 1);
        rc = launch_advance(h, tsave[s]);
        if (rc) return rc;
        if (s >= 2) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->stage_free[b], 0));
        for (size_t v = 0; v < names.size(); ++v) {
            double* dst = h->d_stage[b] + v * NM;
            if (names[v] == "T") CUDA_TRY(h, cudaMemcpyAsync(dst, h->dev.u, NM * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            else if (names[v] == "c") CUDA_TRY(h, cudaMemcpyAsync(dst, h->dev.u + NM, NM * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
            else { rc = cgb_salt_diag_to(h, names[v].c_str(), tsave[s], dst); if (rc) return rc; }
        }
        CUDA_TRY(h, cudaEventRecord(h->stage_ready[b], h->stream));
        CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->stage_ready[b], 0));
        CUDA_TRY(h, cudaMemcpyAsync(out + (size_t)s * per_save, h->d_stage[b], per_save * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
        CUDA_TRY(h, cudaEventRecord(h->stage_free[b], h->copy_stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    return CGB_OK;
}

extern "C" int cgb_solve_saveat(cgb_handle* h, const double* tsave, int64_t nsave, const char* vars, double* out)
{
    if (!h || !tsave || !vars || !out || nsave < 1) return CGB_ERR_INVALID;
    clear_stale_error("cgb_solve_saveat");
    if (!h->have_state) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_set_state has not been called");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) return salt_solve_saveat(h, tsave, nsave, vars, out);
    SaveVars sv;
    rc = parse_savevars(h, vars, sv);
    if (rc) return rc;
    if (sv.names.empty()) CGB_FAIL(h, CGB_ERR_INVALID, "no variables to save");
    size_t NM = (size_t)h->N * h->M, per_save = NM * sv.names.size();
    rc = ensure_staging(h, per_save);
    if (rc) return rc;
    // Device-side saveat: up to `chunk` save times per launch -- a column is carried through all of them in registers, the
    // saved fields of the chunk land contiguously in one staging buffer and leave in one copy.
    bool chunked = multi_target_path(h) && !sv.Tnow && !sv.Wnow;
    for (auto& n : sv.names) chunked = chunked && n != "H";   // H is the live state: it must be copied out after every save
    for (int64_t s = 1; chunked && s < nsave; ++s) chunked = tsave[s] > tsave[s - 1];
    if (chunked && nsave >= 2 && tsave[0] > h->t_front) {
        static const int chunk_max = [] { const char* e = getenv("CGB_SAVE_CHUNK"); int v = e ? atoi(e) : 8; return v >= 1 ? v : 8; }();   // developer knob
        int chunk = (int)std::min<int64_t>(std::min<int64_t>(chunk_max, cgb_fast_max_targets()), nsave);
        // 2 x chunk x per_save doubles of staging: halve the chunk when the device cannot hold that (10^6 columns: GBs per save)
        while ((rc = ensure_staging(h, per_save * chunk)) != CGB_OK && chunk > 1) chunk = std::max(1, chunk / 2);
        if (rc) return rc;
        int64_t c = 0;
        for (int64_t s0 = 0; s0 < nsave; ++c) {
            // chunks shrink towards the end so that the last (un-overlapped) device-to-host copy is a single save
            int n = (int)std::min<int64_t>(chunk, std::max<int64_t>(1, (nsave - s0 + 1) / 2)), b = (int)(c & 1);
            if (c >= 2) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->stage_free[b], 0));
            double *bT = nullptr, *bW = nullptr;
            for (size_t v = 0; v < sv.names.size(); ++v) {
                if (sv.names[v] == "T") bT = h->d_stage[b] + v * NM; else bW = h->d_stage[b] + v * NM;
            }
            // a column that stops early (forcing range, NaN, bad dt: status bits) leaves its remaining save slots unwritten: make them NaN
            // (all-ones bytes) instead of stale data from an earlier chunk; callers must still check cgb_get_status
            CUDA_TRY(h, cudaMemsetAsync(h->d_stage[b], 0xFF, (size_t)n * per_save * sizeof(double), h->stream));
            cudaEvent_t pa = nullptr, pb = nullptr;
            if (h->profiling) { cudaEventCreate(&pa); cudaEventCreate(&pb); cudaEventRecord(pa, h->stream); }
            rc = launch_multi(h, tsave + s0, n, bT, bW, (long long)per_save);
            if (h->profiling) { cudaEventRecord(pb, h->stream); h->prof_events.emplace_back(pa, pb); }
            if (rc) return rc;
            h->t_front = tsave[s0 + n - 1];
            CUDA_TRY(h, cudaEventRecord(h->stage_ready[b], h->stream));
            CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->stage_ready[b], 0));
            CUDA_TRY(h, cudaMemcpyAsync(out + (size_t)s0 * per_save, h->d_stage[b], (size_t)n * per_save * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
            CUDA_TRY(h, cudaEventRecord(h->stage_free[b], h->copy_stream));
            s0 += n;
        }
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
        return CGB_OK;
    }
    for (int64_t s = 0; s < nsave; ++s) {
        int b = (int)(s & 1);
        // the stepping kernel writes straight into staging buffer b: it must have been drained by the copy of save s-2
        if (s >= 2) CUDA_TRY(h, cudaStreamWaitEvent(h->stream, h->stage_free[b], 0));
        double *bT = nullptr, *bW = nullptr, *nT = nullptr, *nW = nullptr;
        for (size_t v = 0; v < sv.names.size(); ++v) {
            double* dst = h->d_stage[b] + v * NM;
            if (sv.names[v] == "T") bT = dst; else if (sv.names[v] == "thw") bW = dst;
            else if (sv.names[v] == "T_now") nT = dst; else if (sv.names[v] == "thw_now") nW = dst;
        }
        rc = advance_and_save(h, tsave[s], sv, bT, bW, nT, nW);
        if (rc) return rc;
        for (size_t v = 0; v < sv.names.size(); ++v)
            if (sv.names[v] == "H") CUDA_TRY(h, cudaMemcpyAsync(h->d_stage[b] + v * NM, h->dev.u, NM * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(h, cudaEventRecord(h->stage_ready[b], h->stream));
        CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->stage_ready[b], 0));
        CUDA_TRY(h, cudaMemcpyAsync(out + (size_t)s * per_save, h->d_stage[b], per_save * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
        CUDA_TRY(h, cudaEventRecord(h->stage_free[b], h->copy_stream));
    }
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->copy_stream));
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// device-side output path: save selected cells only + sub-grid thaw depth per column
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_rows(const double* __restrict__ src, const int* __restrict__ rows, int nsel, long long M, double* __restrict__ dst)
{
    long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (col < M && k < nsel) dst[(size_t)k * M + col] = src[(size_t)rows[k] * M + col];
}

// Diagnostics.thawdepth (Diagnostics/permafrost.jl:38-59): first thawed / first frozen cell, linear in between.
__global__ void __launch_bounds__(128) k_thawdepth(const double* __restrict__ T, const double* __restrict__ zc, int N, long long M, double* __restrict__ out)
{
    long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (col >= M) return;
    int i_frozen = -1, i_thawed = -1;
    double Tf = 0.0, Tt = 0.0;
    // eight cells per round: the loads of a round are independent (one thread per column is latency bound otherwise: a batch of
    // 10^4 columns puts ~70 threads on an SM, and a scan that exits after every cell pays one DRAM round trip per cell)
    for (int i0 = 0; i0 < N && (i_frozen < 0 || i_thawed < 0); i0 += 8) {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (i0 + k < N) ? __ldcs(T + (size_t)(i0 + k) * M + col) : 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (i0 + k < N) {
                if (i_frozen < 0 && v[k] <= 0.0) { i_frozen = i0 + k; Tf = v[k]; }
                if (i_thawed < 0 && v[k] > 0.0) { i_thawed = i0 + k; Tt = v[k]; }
            }
        }
    }
    double td;
    if (i_thawed >= 0 && i_frozen >= 0 && i_thawed < i_frozen) {
        double z1 = zc[i_thawed], z2 = zc[i_frozen];
        td = (0.0 - Tt) * (z2 - z1) / (Tf - Tt) + z1;          // solve_depth_linear, Diagnostics.jl:16
    } else if (i_frozen < 0) td = zc[N - 1];
    else td = zc[0];
    out[col] = td;
}

extern "C" int cgb_solve_saveat_select(cgb_handle* h, const double* tsave, int64_t nsave, const char* vars, const int32_t* cell_index,
                                       int32_t nsel, double* out, double* thawdepth_out)
{
    if (!h || !tsave || !vars || nsave < 1 || (nsel > 0 && (!cell_index || !out)) || (nsel <= 0 && !thawdepth_out)) return CGB_ERR_INVALID;
    clear_stale_error("cgb_solve_saveat_select");
    if (!h->have_state) CGB_FAIL(h, CGB_ERR_INVALID, "no state: call cgb_set_state or cgb_init_* first");
    if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_solve_saveat_select: heat physics only");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    SaveVars sv;
    if (nsel > 0) { rc = parse_savevars(h, vars, sv); if (rc) return rc; }
    std::vector<std::string>& names = sv.names;
    for (int k = 0; k < nsel; ++k)
        if (cell_index[k] < 0 || cell_index[k] >= h->N) CGB_FAIL(h, CGB_ERR_INVALID, "cell_index[%d] = %d out of range", k, cell_index[k]);
    size_t M = (size_t)h->M, NM = (size_t)h->N * M, SM = (size_t)std::max(nsel, 0) * M;
    size_t per_save = names.size() * SM + (thawdepth_out ? M : 0);
    rc = ensure_staging(h, per_save);
    if (rc) return rc;
    int* d_rows = nullptr;
    if (nsel > 0) {
        CUDA_TRY(h, cudaMalloc((void**)&d_rows, nsel * sizeof(int)));
        CUDA_TRY(h, cudaMemcpyAsync(d_rows, cell_index, nsel * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    // full-depth fields land in the four (N+1)*M slots of the diag scratch, the selected rows are gathered from there
    double* fullT = h->d_scratch;
    double* fullW = h->d_scratch + (size_t)(h->N + 1) * M;
    double* nowT = h->d_scratch + 2 * (size_t)(h->N + 1) * M;
    double* nowW = h->d_scratch + 3 * (size_t)(h->N + 1) * M;
    SaveVars need = sv;
    need.T |= thawdepth_out != nullptr;           // thaw depth is computed from the saved T, as Diagnostics.thawdepth(out.T) does
    // Device-side saveat for the selected output too: up to 8 save times per stepping launch (a column stays in registers across
    // them), the full fields of the chunk stay on the device, only the selected rows and thaw depths of the chunk are copied out
    // (two copies per chunk).  Same conditions as in cgb_solve_saveat.
    bool chunked = multi_target_path(h) && !sv.Tnow && !sv.Wnow && nsave >= 2 && tsave[0] > h->t_front;
    for (auto& n : names) chunked = chunked && n != "H";
    for (int64_t s = 1; chunked && s < nsave; ++s) chunked = tsave[s] > tsave[s - 1];
    if (chunked) {
        const size_t vars_save = names.size() * SM;
        const int nfull = (need.T ? 1 : 0) + (need.W ? 1 : 0);
        static const int chunk_max = [] { const char* e = getenv("CGB_SAVE_CHUNK"); int v = e ? atoi(e) : 8; return v >= 1 ? v : 8; }();   // developer knob
        int chunk = (int)std::min<int64_t>(std::min<int64_t>(chunk_max, cgb_fast_max_targets()), nsave);
        while (chunk > 1 && (size_t)chunk * NM * nfull * sizeof(double) > ((size_t)4 << 30)) chunk /= 2;
        rc = ensure_staging(h, (size_t)chunk * per_save);
        if (rc) { cudaFree(d_rows); return rc; }
        const size_t full_elems = (size_t)chunk * NM * std::max(nfull, 1);
        if (h->selfull_elems < full_elems) {
            if (h->d_selfull) cudaFree(h->d_selfull);
            h->d_selfull = nullptr; h->selfull_elems = 0;
            if (cudaMalloc((void**)&h->d_selfull, full_elems * sizeof(double)) != cudaSuccess) { cudaGetLastError(); cudaFree(d_rows); CGB_FAIL(h, CGB_ERR_ALLOC, "full-field chunk buffer (%zu bytes)", full_elems * sizeof(double)); }
            h->selfull_elems = full_elems;
        }
        double* fT = need.T ? h->d_selfull : nullptr;
        double* fW = need.W ? h->d_selfull + (need.T ? (size_t)chunk * NM : 0) : nullptr;
        dim3 grid((unsigned)((M + 255) / 256), (unsigned)std::max(nsel, 1));
        int64_t c = 0;
        for (int64_t s0 = 0; s0 < nsave; ++c) {
            const int n = (int)std::min<int64_t>(chunk, std::max<int64_t>(1, (nsave - s0 + 1) / 2)), b = (int)(c & 1);
            // columns that stop early leave their remaining save slots unwritten: NaN (all-ones bytes), not stale data
            cudaMemsetAsync(h->d_selfull, 0xFF, (size_t)chunk * NM * std::max(nfull, 1) * sizeof(double), h->stream);
            rc = launch_multi(h, tsave + s0, n, fT, fW, (long long)NM);
            if (rc) { cudaFree(d_rows); return rc; }
            h->t_front = tsave[s0 + n - 1];
            if (c >= 2) cudaStreamWaitEvent(h->stream, h->stage_free[b], 0);
            for (int k = 0; k < n; ++k) {
                for (size_t v = 0; v < names.size(); ++v) {
                    const double* src = (names[v] == "T" ? fT : fW) + (size_t)k * NM;
                    k_gather_rows<<<grid, 256, 0, h->stream>>>(src, d_rows, nsel, (long long)M, h->d_stage[b] + (size_t)k * vars_save + v * SM);
                    h->launches++;
                }
                if (thawdepth_out) {
                    k_thawdepth<<<(unsigned)((M + 127) / 128), 128, 0, h->stream>>>(fT + (size_t)k * NM, h->dev.zc, h->N, (long long)M,
                                                                                   h->d_stage[b] + (size_t)chunk * vars_save + (size_t)k * M);
                    h->launches++;
                }
            }
            cudaEventRecord(h->stage_ready[b], h->stream);
            cudaStreamWaitEvent(h->copy_stream, h->stage_ready[b], 0);
            if (!names.empty())
                cudaMemcpyAsync(out + (size_t)s0 * vars_save, h->d_stage[b], (size_t)n * vars_save * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream);
            if (thawdepth_out)
                cudaMemcpyAsync(thawdepth_out + (size_t)s0 * M, h->d_stage[b] + (size_t)chunk * vars_save, (size_t)n * M * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream);
            cudaEventRecord(h->stage_free[b], h->copy_stream);
            s0 += n;
        }
        cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->copy_stream);
        cudaFree(d_rows);
        if (e1 != cudaSuccess || e2 != cudaSuccess) CGB_FAIL(h, CGB_ERR_CUDA, "cgb_solve_saveat_select: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        return CGB_OK;
    }
    for (int64_t s = 0; s < nsave; ++s) {
        int b = (int)(s & 1);
        rc = advance_and_save(h, tsave[s], need, fullT, fullW, nowT, nowW);
        if (rc) { cudaFree(d_rows); return rc; }
        if (s >= 2) cudaStreamWaitEvent(h->stream, h->stage_free[b], 0);
        dim3 grid((unsigned)((M + 255) / 256), (unsigned)std::max(nsel, 1));
        for (size_t v = 0; v < names.size(); ++v) {
            const double* src = names[v] == "H" ? h->dev.u : names[v] == "T" ? fullT : names[v] == "thw" ? fullW : names[v] == "T_now" ? nowT : nowW;
            k_gather_rows<<<grid, 256, 0, h->stream>>>(src, d_rows, nsel, (long long)M, h->d_stage[b] + v * SM);
            h->launches++;
        }
        if (thawdepth_out) {
            k_thawdepth<<<(unsigned)((M + 127) / 128), 128, 0, h->stream>>>(fullT, h->dev.zc, h->N, (long long)M, h->d_stage[b] + names.size() * SM);
            h->launches++;
        }
        cudaEventRecord(h->stage_ready[b], h->stream);
        cudaStreamWaitEvent(h->copy_stream, h->stage_ready[b], 0);
        if (!names.empty())
            cudaMemcpyAsync(out + (size_t)s * names.size() * SM, h->d_stage[b], names.size() * SM * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream);
        if (thawdepth_out)
            cudaMemcpyAsync(thawdepth_out + (size_t)s * M, h->d_stage[b] + names.size() * SM, M * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream);
        cudaEventRecord(h->stage_free[b], h->copy_stream);
    }
    cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->copy_stream);
    cudaFree(d_rows);
    (void)NM;
    if (e1 != cudaSuccess || e2 != cudaSuccess) CGB_FAIL(h, CGB_ERR_CUDA, "cgb_solve_saveat_select: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// permafrost diagnostics reduced on the device over the saves of one call (Diagnostics/permafrost.jl:7-113): the saved T fields
// never leave the GPU, only M-sized results do
// ------------------------------------------------------------------------------------------------
// running reductions after every save: per-cell max / min of the saved T, per-column max thaw depth (ALT) and the depth-mean T
__global__ void __launch_bounds__(128) k_permafrost_update(const double* __restrict__ T, const double* __restrict__ zc, int N, long long M, int first,
                                                          double Tmelt, int i_lo, int i_hi, double* __restrict__ Tmax, double* __restrict__ Tmin,
                                                          double* __restrict__ alt, double* __restrict__ magt_out)
{
    long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (col >= M) return;
    int i_frozen = -1, i_thawed = -1;
    double Tf = 0.0, Tt = 0.0, sum = 0.0;
    for (int i = 0; i < N; ++i) {
        const size_t o = (size_t)i * M + col;
        const double v = T[o];
        if (i_frozen < 0 && v <= Tmelt) { i_frozen = i; Tf = v; }
        if (i_thawed < 0 && v > Tmelt) { i_thawed = i; Tt = v; }
        Tmax[o] = first ? v : fmax(Tmax[o], v);
        Tmin[o] = first ? v : fmin(Tmin[o], v);
        if (i >= i_lo && i <= i_hi) sum += v;
    }
    double td;                                                     // thawdepth, permafrost.jl:38-59
    if (i_thawed >= 0 && i_frozen >= 0 && i_thawed < i_frozen) td = (Tmelt - Tt) * (zc[i_frozen] - zc[i_thawed]) / (Tf - Tt) + zc[i_thawed];
    else if (i_frozen < 0) td = zc[N - 1];
    else td = zc[0];
    alt[col] = first ? td : fmax(alt[col], td);                    // active_layer_thickness, permafrost.jl:66-72 (max over the window)
    if (magt_out) magt_out[col] = (i_hi >= i_lo) ? sum / (double)(i_hi - i_lo + 1) : nan("");   // mean_annual_ground_temperature, :108-113
}

// final pass: permafrost table / base (permafrost.jl:80-101, verbatim incl. the index arithmetic of permafrostbase) and the depth of
// zero annual amplitude (:7-30) from the per-cell window max / min
__global__ void __launch_bounds__(128) k_permafrost_final(const double* __restrict__ Tmax, const double* __restrict__ Tmin, const double* __restrict__ zc, int N,
                                                         long long M, double Tmelt, double threshold, double* __restrict__ table, double* __restrict__ base,
                                                         double* __restrict__ zaa)
{
    long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (col >= M) return;
    int first_frozen = 0, last_frozen = -1, i_amp = -1;            // argmax of an all-false mask is the first index
    bool found = false;
    for (int i = 0; i < N; ++i) {
        const size_t o = (size_t)i * M + col;
        const bool fr = Tmax[o] < Tmelt;
        if (fr && !found) { first_frozen = i; found = true; }
        if (fr) last_frozen = i;
        if (i_amp < 0 && (Tmax[o] - Tmin[o]) < threshold) i_amp = i;
    }
    if (table) table[col] = zc[first_frozen];
    if (base) {
        // z_inds = N .- argmax(reverse(mask)): r = 1-based position of the deepest frozen cell counted from the bottom (1 if none),
        // 1-based index N - r, i.e. ONE CELL ABOVE the deepest frozen cell (reproduced verbatim; index 0 cannot occur in Julia)
        int r = (last_frozen >= 0) ? (N - last_frozen) : 1;
        int idx1 = N - r;
        base[col] = zc[idx1 >= 1 ? idx1 - 1 : 0];
    }
    if (zaa) {
        if (i_amp > 0) {
            const size_t a = (size_t)(i_amp - 1) * M + col, b = (size_t)i_amp * M + col;
            const double A1 = Tmax[a] - Tmin[a], A2 = Tmax[b] - Tmin[b];
            zaa[col] = (threshold - A1) * (zc[i_amp] - zc[i_amp - 1]) / (A2 - A1) + zc[i_amp - 1];      // solve_depth_linear, Diagnostics.jl:16
        } else zaa[col] = (i_amp == 0) ? zc[0] : INFINITY;
    }
}

extern "C" int cgb_solve_saveat_permafrost(cgb_handle* h, const double* tsave, int64_t nsave, double Tmelt, double zaa_threshold, double magt_upper,
                                           double magt_lower, double* alt, double* pf_table, double* pf_base, double* zaa, double* magt)
{
    if (!h || !tsave || nsave < 1) return CGB_ERR_INVALID;
    clear_stale_error("cgb_solve_saveat_permafrost");
    if (!h->have_state) CGB_FAIL(h, CGB_ERR_INVALID, "no state: call cgb_set_state or cgb_init_* first");
    if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "cgb_solve_saveat_permafrost: heat physics only");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    const size_t M = (size_t)h->M, NM1 = (size_t)(h->N + 1) * M;
    double* fullT = h->d_scratch;                       // saved T of the current save (reference-verbatim "T")
    double* Tmax = h->d_scratch + 2 * NM1;
    double* Tmin = h->d_scratch + 3 * NM1;
    // cells with magt_upper <= z <= magt_lower (Z(Between(lower, upper)) on the cell midpoints)
    int i_lo = h->N, i_hi = -1;
    for (int i = 0; i < h->N; ++i) {
        double z = (h->edges[i] + h->edges[i + 1]) / 2.0;
        if (z >= magt_upper && z <= magt_lower) { i_lo = std::min(i_lo, i); i_hi = std::max(i_hi, i); }
    }
    double* d_small = nullptr;                          // alt, table, base, zaa [M each] + magt [nsave*M]
    const size_t n_small = 4 * M + (magt ? (size_t)nsave * M : 0);
    CUDA_TRY(h, cudaMalloc((void**)&d_small, n_small * sizeof(double)));
    double *d_alt = d_small, *d_table = d_small + M, *d_base = d_small + 2 * M, *d_zaa = d_small + 3 * M, *d_magt = magt ? d_small + 4 * M : nullptr;
    SaveVars need; need.T = true; need.names.push_back("T");
    const unsigned grid = (unsigned)((M + 127) / 128);
    for (int64_t s = 0; s < nsave; ++s) {
        rc = advance_and_save(h, tsave[s], need, fullT, nullptr, nullptr, nullptr);
        if (rc) { cudaFree(d_small); return rc; }
        k_permafrost_update<<<grid, 128, 0, h->stream>>>(fullT, h->dev.zc, h->N, (long long)M, s == 0, Tmelt, i_lo, i_hi, Tmax, Tmin, d_alt,
                                                         d_magt ? d_magt + (size_t)s * M : nullptr);
        h->launches++;
    }
    k_permafrost_final<<<grid, 128, 0, h->stream>>>(Tmax, Tmin, h->dev.zc, h->N, (long long)M, Tmelt, zaa_threshold, d_table, d_base, d_zaa);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    auto back = [&](double* dst, const double* src, size_t n) { if (dst && e == cudaSuccess) e = cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream); };
    back(alt, d_alt, M); back(pf_table, d_table, M); back(pf_base, d_base, M); back(zaa, d_zaa, M); back(magt, d_magt, (size_t)nsave * M);
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    cudaFree(d_small);
    if (e != cudaSuccess || e2 != cudaSuccess) CGB_FAIL(h, CGB_ERR_CUDA, "cgb_solve_saveat_permafrost: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// ensemble partial sums: one block per cell, coalesced over columns
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ensemble_partial(const double* __restrict__ x, long long M, int N, double* __restrict__ out)
{
    int cell = blockIdx.x;
    const double* row = x + (size_t)cell * M;
    double s = 0.0, q = 0.0;
    for (long long c = threadIdx.x; c < M; c += blockDim.x) { double v = row[c]; s += v; q = fma(v, v, q); }
    __shared__ double ss[8], sq[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(FULL, s, o); q += __shfl_xor_sync(FULL, q, o); }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0;
        for (int w = 0; w < 8; ++w) { a += ss[w]; b += sq[w]; }
        out[cell] = a; out[N + cell] = b;
    }
}

extern "C" int cgb_ensemble_partial_device(cgb_handle* h, const char* var, double t, void* out_dev)
{
    clear_stale_error("cgb_ensemble_partial_device");
    if (!h || !var || !out_dev) return CGB_ERR_INVALID;
    if (!h->have_state) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_set_state has not been called");
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    const double* src = nullptr;
    if (std::strcmp(var, "H") == 0 && h->cfg.physics == CGB_PHYSICS_HEAT) src = h->dev.u;
    else {
        if (h->cfg.physics == CGB_PHYSICS_HEAT_SALT) CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "ensemble moments: heat only");
        Diag D{}; double* p; size_t cnt;
        rc = diag_slot(h, var, D, &p, &cnt);
        if (rc) return rc;
        if (cnt != (size_t)h->N * h->M) CGB_FAIL(h, CGB_ERR_INVALID, "ensemble moments need a cell variable");
        rc = launch_diag(h, t, D);
        if (rc) return rc;
        src = p;
    }
    k_ensemble_partial<<<h->N, 256, 0, h->stream>>>(src, h->M, h->N, static_cast<double*>(out_dev));
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return CGB_OK;
}

extern "C" int cgb_ensemble_partial(cgb_handle* h, const char* var, double t, double* sum, double* sumsq)
{
    if (!h || !sum || !sumsq) return CGB_ERR_INVALID;
    double* tmp = h->d_scratch + (size_t)(cgb_handle::kScratch - 1) * (h->N + 1) * h->M;   // last scratch slot
    int rc = cgb_ensemble_partial_device(h, var, t, tmp);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpy(sum, tmp, h->N * sizeof(double), cudaMemcpyDeviceToHost));
    CUDA_TRY(h, cudaMemcpy(sumsq, tmp + h->N, h->N * sizeof(double), cudaMemcpyDeviceToHost));
    return CGB_OK;
}

// ------------------------------------------------------------------------------------------------
// measurement
// ------------------------------------------------------------------------------------------------
extern "C" int cgb_profile_enable(cgb_handle* h, int32_t on)
{
    if (!h) return CGB_ERR_INVALID;
    h->profiling = on != 0;
    return CGB_OK;
}

extern "C" int cgb_profile_read(cgb_handle* h, int64_t* n_launches, double* total_ms)
{
    if (!h || !n_launches || !total_ms) return CGB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    for (auto& pr : h->prof_events) {
        float ms = 0.f;
        CUDA_TRY(h, cudaEventElapsedTime(&ms, pr.first, pr.second));
        tot += ms;
        cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
    }
    *n_launches = (int64_t)h->prof_events.size();
    *total_ms = tot;
    h->prof_events.clear();
    return CGB_OK;
}

// self-test of the reciprocal sequences used by the kernels (accuracy is asserted in tests/test_gpu_explicit.py)
__global__ void k_selftest_rcp(const double* __restrict__ x, long long n, double* raw, double* n2, double* c3)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { raw[i] = rcp_raw(x[i]); n2[i] = rcp_fast(x[i]); c3[i] = rcp_fast3(x[i]); }
}

extern "C" int cgb_selftest_rcp(const double* x, int64_t n, double* out_raw, double* out_newton2, double* out_cubic)
{
    if (!x || n < 1 || !out_raw || !out_newton2 || !out_cubic) return CGB_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return CGB_ERR_NO_DEVICE;
    double* d = nullptr;
    if (cudaMalloc((void**)&d, 4 * n * sizeof(double)) != cudaSuccess) return CGB_ERR_ALLOC;
    cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
    k_selftest_rcp<<<(unsigned)((n + 255) / 256), 256>>>(d, n, d + n, d + 2 * n, d + 3 * n);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out_raw, d + n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(out_newton2, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(out_cubic, d + 3 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? CGB_OK : CGB_ERR_CUDA;
}

// self-test of the curve table of the SFCC single-class fast path: θw(T) as the stepping kernel evaluates it
extern "C" int cgb_selftest_sfcc_curve(cgb_handle* h, const double* T, int64_t n, double* thw_out)
{
    if (!h || !T || !thw_out || n < 1) return CGB_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->cfg.device));
    int rc = finalize(h);
    if (rc) return rc;
    return cgb_sfcc_fast_selftest(h, T, (long long)n, thw_out);
}

// self-test of the transcendental replacements used by the van Genuchten powers (tests/test_gpu_sfcc.py)
__global__ void k_selftest_explog(const double* __restrict__ x, long long n, double* ex, double* lg, int tab)
{
    vgtab_init();
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) {
        if (tab) { ex[i] = exp_tab(x[i]); lg[i] = log_tab(fabs(x[i])); }
        else { ex[i] = exp_fast(x[i]); lg[i] = log_pos(fabs(x[i])); }
    }
}

static int selftest_explog(const double* x, int64_t n, double* out_exp, double* out_log, int tab);
extern "C" int cgb_selftest_explog(const double* x, int64_t n, double* out_exp, double* out_log) { return selftest_explog(x, n, out_exp, out_log, 0); }
// the table-driven variants used by the stepping kernels (exp_tab / log_tab, csrc/cgb_common.cuh)
extern "C" int cgb_selftest_explog_tab(const double* x, int64_t n, double* out_exp, double* out_log) { return selftest_explog(x, n, out_exp, out_log, 1); }
static int selftest_explog(const double* x, int64_t n, double* out_exp, double* out_log, int tab)
{
    if (!x || n < 1 || !out_exp || !out_log) return CGB_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return CGB_ERR_NO_DEVICE;
    double* d = nullptr;
    if (cudaMalloc((void**)&d, 3 * n * sizeof(double)) != cudaSuccess) return CGB_ERR_ALLOC;
    cudaMemcpy(d, x, n * sizeof(double), cudaMemcpyHostToDevice);
    k_selftest_explog<<<(unsigned)((n + 255) / 256), 256>>>(d, n, d + n, d + 2 * n, tab);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out_exp, d + n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaMemcpy(out_log, d + 2 * n, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? CGB_OK : CGB_ERR_CUDA;
}
