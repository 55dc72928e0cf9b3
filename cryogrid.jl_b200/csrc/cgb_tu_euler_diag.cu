// cgb_tu_euler_diag.cu -- DIAG instantiations of the general explicit kernel (one RHS evaluation, all diagnostics).
#include "cgb_tu_euler_common.cuh"

int cgb_launch_euler_diag(cgb_handle* h, double t, const Diag& D) { return launch_euler_c<0, true>(h, t, D); }
