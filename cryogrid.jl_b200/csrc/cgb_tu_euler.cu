// cgb_tu_euler.cu -- stepping instantiations of the general explicit path (solver modes 0..3): k_heat_step.
#include "cgb_launch.h"
#include "cgb_euler_step.cuh"
using namespace cgb;

template <int CPL, int MODE>
static int launch_step_t(cgb_handle* h, const FastTargets& tg, const Diag& D)
{
    auto kern = k_heat_step<CPL, MODE>;
    size_t smem = (size_t)(5 * CPL * 32 + 4 * h->nl * LC_STRIDE) * sizeof(double);
    int blocks_per_sm = 0;
    CUDA_TRY(h, cgb_prepare_kernel(h, (const void*)kern, smem, &blocks_per_sm));
    long long tiles = (h->M + 3) / 4;
    long long grid = std::min<long long>(tiles, (long long)h->num_sms * blocks_per_sm);
    CUDA_TRY(h, cudaMemsetAsync(h->dev.counter, 0, sizeof(unsigned long long), h->stream));
    kern<<<(unsigned)grid, 128, smem, h->stream>>>(h->dev, tg, D);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return CGB_OK;
}

template <int MODE>
static int launch_step_c(cgb_handle* h, const FastTargets& tg, const Diag& D)
{
    int cpl = (h->N + 31) / 32;
    switch (cpl) {
    case 1: return launch_step_t<1, MODE>(h, tg, D);
    case 2: return launch_step_t<2, MODE>(h, tg, D);
    case 3: case 4: return launch_step_t<4, MODE>(h, tg, D);
    case 5: case 6: return launch_step_t<6, MODE>(h, tg, D);
    case 7: return launch_step_t<7, MODE>(h, tg, D);
    case 8: return launch_step_t<8, MODE>(h, tg, D);
    case 9: return launch_step_t<9, MODE>(h, tg, D);
    case 10: return launch_step_t<10, MODE>(h, tg, D);
    case 11: case 12: return launch_step_t<12, MODE>(h, tg, D);
    default: return launch_step_t<16, MODE>(h, tg, D);
    }
}

int cgb_launch_euler_step_multi(cgb_handle* h, int mode, const double* targets, int n, const Diag& D, long long save_stride)
{
    if (n < 1 || n > FAST_MAX_TARGETS) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_launch_euler_step_multi: 1..%d targets", FAST_MAX_TARGETS);
    FastTargets tg{};
    for (int k = 0; k < n; ++k) tg.t[k] = targets[k];
    tg.n = n; tg.save_stride = save_stride;
    switch (mode) {
    case 1: return launch_step_c<1>(h, tg, D);
    case 2: return launch_step_c<2>(h, tg, D);
    case 3: return launch_step_c<3>(h, tg, D);
    case 4: return launch_step_c<4>(h, tg, D);
    case 5: return launch_step_c<5>(h, tg, D);
    default: return launch_step_c<0>(h, tg, D);
    }
}

int cgb_launch_euler_step(cgb_handle* h, int mode, double t_target, const Diag& D)
{
    return cgb_launch_euler_step_multi(h, mode, &t_target, 1, D, 0);
}

int cgb_launch_euler(cgb_handle* h, int mode, bool diag, double t_target, const Diag& D)
{
    return diag ? cgb_launch_euler_diag(h, t_target, D) : cgb_launch_euler_step(h, mode, t_target, D);
}
