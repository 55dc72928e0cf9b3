// cgb_tu_lite_warp.cu -- LiteImplicitEuler, warp-per-column variant.
#include "cgb_launch.h"
#include "cgb_lite_warp.cuh"

int cgb_launch_lite_warp_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride)
{
    if (n < 1 || n > cgb::FAST_MAX_TARGETS) CGB_FAIL(h, CGB_ERR_INVALID, "cgb_launch_lite_warp_multi: 1..%d targets", cgb::FAST_MAX_TARGETS);
    cgb::FastTargets tg{};
    for (int k = 0; k < n; ++k) tg.t[k] = targets[k];
    tg.n = n; tg.save_stride = save_stride;
    return lite_warp_launch(h, tg, saveT, saveW);
}

int cgb_launch_lite_warp(cgb_handle* h, double t_target, double* saveT, double* saveW)
{
    return cgb_launch_lite_warp_multi(h, &t_target, 1, saveT, saveW, 0);
}
