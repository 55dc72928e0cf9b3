// cgb_tu_newton_tab.cu -- eligibility and launcher of k_heat_newton_tab (cgb_euler_newton_tab.cuh): SFCC columns with per-column
// parameters through the on-device Newton solver, one layer, MaxDelta limiter (BASELINE cfg4).
#include "cgb_launch.h"
#include "cgb_euler_newton_tab.cuh"
using namespace cgb;

#ifndef CGB_NT_WARPS
#define CGB_NT_WARPS 8
#endif
constexpr int kNtWarps = CGB_NT_WARPS;

bool cgb_newton_tab_ok(const cgb_handle* h)
{
    const char* e = getenv("CGB_NEWTON_TAB");
    if (e && atoi(e) == 0) return false;                   // A/B knob: the general Newton kernel k_heat_step<.,2>
    if (h->cfg.physics != CGB_PHYSICS_HEAT || h->cfg.stepper != CGB_STEPPER_EULER || h->cfg.limiter != CGB_LIMITER_MAXDELTA) return false;
    if (h->nl != 1 || h->fc_solver[0] != CGB_SOLVER_NEWTON) return false;
    if (h->fc_kind[0] != CGB_FC_DALLAMICO && h->fc_kind[0] != CGB_FC_PAINTERKARRA) return false;
    return h->cfg.bot_kind == CGB_BC_NEUMANN && (h->params.count("bot_value") || h->forc[1].kind == 0);
}

// Room for every column's fitted curve table (NT_ROWS x 8 doubles) in HBM, when that is a small part of what is free: the
// first launch after the description changed fits and stores, later launches load.  CGB_NT_CACHE=0: fit in every launch.
int cgb_newton_tab_prepare(cgb_handle* h)
{
    h->nt_cache = nullptr; h->nt_cache_valid = false;
    if (!cgb_newton_tab_ok(h)) return CGB_OK;
    const char* e = getenv("CGB_NT_CACHE");
    if (e && atoi(e) == 0) return CGB_OK;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return CGB_OK; }
    const size_t need = (size_t)h->M * NT_ROWS * 8 * sizeof(double);
    if (need > free_b / 4) return CGB_OK;
    return dev_alloc(h, &h->nt_cache, (size_t)h->M * NT_ROWS * 8);
}

template <int CPL, bool SAVE>
static int launch_nt_s(cgb_handle* h, const FastTargets& tg, double* saveT, double* saveW)
{
    auto kern = k_heat_newton_tab<CPL, kNtWarps, SAVE>;
    const size_t smem = (size_t)kNtWarps * NT_ROWS * NT_STRIDE * sizeof(double);
    CUDA_TRY(h, cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = std::min<long long>((h->M + kNtWarps - 1) / kNtWarps, (long long)h->num_sms);
    CUDA_TRY(h, cudaMemsetAsync(h->dev.counter, 0, sizeof(unsigned long long), h->stream));
    kern<<<(unsigned)grid, kNtWarps * 32, smem, h->stream>>>(h->dev, tg, saveT, saveW, h->nt_cache, h->nt_cache_valid ? 1 : 0);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    if (h->nt_cache) h->nt_cache_valid = true;
    return CGB_OK;
}

template <int CPL>
static int launch_nt_t(cgb_handle* h, const FastTargets& tg, double* saveT, double* saveW)
{
    if (saveT || saveW) return launch_nt_s<CPL, true>(h, tg, saveT, saveW);
    return launch_nt_s<CPL, false>(h, tg, nullptr, nullptr);
}

int cgb_launch_newton_tab_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride)
{
    if (n < 1 || n > FAST_MAX_TARGETS) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_launch_newton_tab_multi: 1..%d targets", FAST_MAX_TARGETS);
    FastTargets tg{};
    for (int k = 0; k < n; ++k) tg.t[k] = targets[k];
    tg.n = n; tg.save_stride = save_stride;
    int cpl = (h->N + 31) / 32;
    switch (cpl) {
    case 1: return launch_nt_t<1>(h, tg, saveT, saveW);
    case 2: return launch_nt_t<2>(h, tg, saveT, saveW);
    case 3: case 4: return launch_nt_t<4>(h, tg, saveT, saveW);
    case 5: case 6: return launch_nt_t<6>(h, tg, saveT, saveW);
    case 7: return launch_nt_t<7>(h, tg, saveT, saveW);
    case 8: return launch_nt_t<8>(h, tg, saveT, saveW);
    case 9: return launch_nt_t<9>(h, tg, saveT, saveW);
    case 10: return launch_nt_t<10>(h, tg, saveT, saveW);
    case 11: case 12: return launch_nt_t<12>(h, tg, saveT, saveW);
    default: return launch_nt_t<16>(h, tg, saveT, saveW);
    }
}
