// shared launcher template of the general explicit kernel (included by cgb_tu_euler.cu and cgb_tu_euler_diag.cu)
#pragma once
#include "cgb_launch.h"
#include "cgb_euler.cuh"
using namespace cgb;

template <int CPL, int MODE, bool DIAG>
static int launch_euler_t(cgb_handle* h, double t_target, const Diag& D)
{
    auto kern = k_heat_euler<CPL, MODE, DIAG>;
    size_t smem = (size_t)(5 * CPL * 32 + 4 * h->nl * LC_STRIDE) * sizeof(double);
    int blocks_per_sm = 0;
    CUDA_TRY(h, cgb_prepare_kernel(h, (const void*)kern, smem, &blocks_per_sm));
    long long tiles = (h->M + 3) / 4;
    long long grid = std::min<long long>(tiles, (long long)h->num_sms * blocks_per_sm);
    CUDA_TRY(h, cudaMemsetAsync(h->dev.counter, 0, sizeof(unsigned long long), h->stream));
    kern<<<(unsigned)grid, 128, smem, h->stream>>>(h->dev, t_target, D);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return CGB_OK;
}

template <int MODE, bool DIAG>
static int launch_euler_c(cgb_handle* h, double t_target, const Diag& D)
{
    int cpl = (h->N + 31) / 32;
    switch (cpl) {
    case 1: return launch_euler_t<1, MODE, DIAG>(h, t_target, D);
    case 2: return launch_euler_t<2, MODE, DIAG>(h, t_target, D);
    case 3: case 4: return launch_euler_t<4, MODE, DIAG>(h, t_target, D);
    case 5: case 6: return launch_euler_t<6, MODE, DIAG>(h, t_target, D);
    case 7: return launch_euler_t<7, MODE, DIAG>(h, t_target, D);
    case 8: return launch_euler_t<8, MODE, DIAG>(h, t_target, D);
    case 9: return launch_euler_t<9, MODE, DIAG>(h, t_target, D);
    case 10: return launch_euler_t<10, MODE, DIAG>(h, t_target, D);
    case 11: case 12: return launch_euler_t<12, MODE, DIAG>(h, t_target, D);
    default: return launch_euler_t<16, MODE, DIAG>(h, t_target, D);
    }
}

