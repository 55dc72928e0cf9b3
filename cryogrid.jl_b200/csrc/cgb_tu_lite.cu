// cgb_tu_lite.cu -- LiteImplicitEuler, Thomas-per-thread variant + implicit diagnostics.
#include "cgb_launch.h"
#include "cgb_lite.cuh"

int cgb_launch_lite_thread(cgb_handle* h, double t, bool diag, const cgb::Diag& D) { return lite_launch(h, t, diag, D); }
