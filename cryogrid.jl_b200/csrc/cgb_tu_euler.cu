// cgb_tu_euler.cu -- stepping instantiations of the general explicit kernel (solver modes 0..3).
#include "cgb_tu_euler_common.cuh"

int cgb_launch_euler_step(cgb_handle* h, int mode, double t_target, const Diag& D)
{
    switch (mode) {
    case 1: return launch_euler_c<1, false>(h, t_target, D);
    case 2: return launch_euler_c<2, false>(h, t_target, D);
    case 3: return launch_euler_c<3, false>(h, t_target, D);
    default: return launch_euler_c<0, false>(h, t_target, D);
    }
}

int cgb_launch_euler(cgb_handle* h, int mode, bool diag, double t_target, const Diag& D)
{
    return diag ? cgb_launch_euler_diag(h, t_target, D) : cgb_launch_euler_step(h, mode, t_target, D);
}
