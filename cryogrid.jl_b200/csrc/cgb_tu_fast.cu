// cgb_tu_fast.cu -- launcher of the FreeWater fast kernel (cgb_euler_fast.cuh).
#include "cgb_launch.h"
#include "cgb_euler_fast.cuh"
using namespace cgb;

// Tuned on B200 (profiles/): 3 resident CTAs per SM (168 registers) and the cubic reciprocal.  Re-measured with the
// x3-unrolled step loop (cell-steps/s on cfg2): MINB 3 + cubic 3.62e11, MINB 3 + two Newton steps 3.47e11, MINB 2 (230 registers)
// 3.54e11, MINB 4 (128 registers, 224 B spills) 3.53e11.  The macros exist for such A/B builds (build.py: CGB_NVCC_EXTRA).
#ifndef CGB_FAST_MINB
#define CGB_FAST_MINB 3
#endif
#ifndef CGB_FAST_CUBIC
#define CGB_FAST_CUBIC true
#endif
constexpr int kFastMinB = CGB_FAST_MINB;
constexpr bool kFastCubic = CGB_FAST_CUBIC;

template <int CPL, bool SAVE>
static int launch_fast_s(cgb_handle* h, const FastTargets& tg, double* saveT, double* saveW)
{
    auto kern = k_heat_euler_fast<CPL, kFastMinB, kFastCubic, SAVE>;
    size_t smem = (size_t)(4 * h->nl * FC_STRIDE) * sizeof(double);
    int blocks_per_sm = 0;
    CUDA_TRY(h, cgb_prepare_kernel(h, (const void*)kern, smem, &blocks_per_sm));
    long long tiles = (h->M + 3) / 4;
    long long grid = std::min<long long>(tiles, (long long)h->num_sms * blocks_per_sm);
    CUDA_TRY(h, cudaMemsetAsync(h->dev.counter, 0, sizeof(unsigned long long), h->stream));
    kern<<<(unsigned)grid, 128, smem, h->stream>>>(h->dev, tg, saveT, saveW);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return CGB_OK;
}

template <int CPL>
static int launch_fast_t(cgb_handle* h, const FastTargets& tg, double* saveT, double* saveW)
{
    // the saving instantiation is separate so that the plain stepping kernel carries no extra registers or branches
    if (saveT || saveW) return launch_fast_s<CPL, true>(h, tg, saveT, saveW);
    return launch_fast_s<CPL, false>(h, tg, nullptr, nullptr);
}

int cgb_launch_fast(cgb_handle* h, double t_target, double* saveT, double* saveW)
{
    return cgb_launch_fast_multi(h, &t_target, 1, saveT, saveW, 0);
}

int cgb_fast_max_targets(void) { return FAST_MAX_TARGETS; }

int cgb_launch_fast_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride)
{
    if (n < 1 || n > FAST_MAX_TARGETS) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_launch_fast_multi: 1..%d targets", FAST_MAX_TARGETS);
    FastTargets tg{};
    for (int k = 0; k < n; ++k) tg.t[k] = targets[k];
    tg.n = n; tg.save_stride = save_stride;
    const double t_target = 0.0; (void)t_target;
    int cpl = (h->N + 31) / 32;
    switch (cpl) {
    case 1: return launch_fast_t<1>(h, tg, saveT, saveW);
    case 2: return launch_fast_t<2>(h, tg, saveT, saveW);
    case 3: case 4: return launch_fast_t<4>(h, tg, saveT, saveW);
    case 5: case 6: return launch_fast_t<6>(h, tg, saveT, saveW);
    case 7: return launch_fast_t<7>(h, tg, saveT, saveW);
    case 8: return launch_fast_t<8>(h, tg, saveT, saveW);
    case 9: return launch_fast_t<9>(h, tg, saveT, saveW);
    case 10: return launch_fast_t<10>(h, tg, saveT, saveW);
    case 11: case 12: return launch_fast_t<12>(h, tg, saveT, saveW);
    default: return launch_fast_t<16>(h, tg, saveT, saveW);
    }
}
