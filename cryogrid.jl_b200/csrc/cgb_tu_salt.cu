// cgb_tu_salt.cu -- coupled heat + salt kernels.
#include "cgb_launch.h"
#include "cgb_salt.cuh"

int cgb_salt_finalize(cgb_handle* h, const std::vector<double>& zc) { return salt_finalize(h, zc); }
int cgb_salt_launch(cgb_handle* h, double t_target, bool diag, const cgb::Diag& D, void* sd) { return salt_launch(h, t_target, diag, D, sd); }
int cgb_salt_get_diag(cgb_handle* h, const char* var, double t, double* out) { return salt_get_diag(h, var, t, out); }
int cgb_salt_diag_to(cgb_handle* h, const char* var, double t, double* dev_out) { return salt_diag_to(h, var, t, dev_out, nullptr); }
