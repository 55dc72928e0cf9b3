// cgb_tu_lite_warp.cu -- LiteImplicitEuler, warp-per-column variant.
#include "cgb_launch.h"
#include "cgb_lite_warp.cuh"

int cgb_launch_lite_warp(cgb_handle* h, double t_target, double* saveT, double* saveW) { return lite_warp_launch(h, t_target, saveT, saveW); }
