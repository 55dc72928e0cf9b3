// cgb_tu_euler_diag.cu -- the diagnostics kernel of the explicit path (one RHS evaluation, all diagnostics).
#include "cgb_launch.h"
#include "cgb_euler.cuh"
using namespace cgb;

template <int CPL>
static int launch_diag_t(cgb_handle* h, double t, const Diag& D)
{
    auto kern = k_heat_diag<CPL>;
    size_t smem = (size_t)(4 * CPL * 32 + 4 * h->nl * LC_STRIDE) * sizeof(double);
    int blocks_per_sm = 0;
    CUDA_TRY(h, cgb_prepare_kernel(h, (const void*)kern, smem, &blocks_per_sm));
    long long grid = std::min<long long>((h->M + 3) / 4, (long long)h->num_sms * blocks_per_sm);
    CUDA_TRY(h, cudaMemsetAsync(h->dev.counter, 0, sizeof(unsigned long long), h->stream));
    kern<<<(unsigned)grid, 128, smem, h->stream>>>(h->dev, t, D);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return CGB_OK;
}

int cgb_launch_euler_diag(cgb_handle* h, double t, const Diag& D)
{
    int cpl = (h->N + 31) / 32;
    switch (cpl) {
    case 1: return launch_diag_t<1>(h, t, D);
    case 2: return launch_diag_t<2>(h, t, D);
    case 3: case 4: return launch_diag_t<4>(h, t, D);
    case 5: case 6: return launch_diag_t<6>(h, t, D);
    case 7: return launch_diag_t<7>(h, t, D);
    case 8: return launch_diag_t<8>(h, t, D);
    case 9: return launch_diag_t<9>(h, t, D);
    case 10: return launch_diag_t<10>(h, t, D);
    case 11: case 12: return launch_diag_t<12>(h, t, D);
    default: return launch_diag_t<16>(h, t, D);
    }
}
