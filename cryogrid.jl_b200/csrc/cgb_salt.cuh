// cgb_salt.cuh -- heat+salt kernels (placeholder until implemented)
#pragma once
#include "cgb_handle.h"
static int salt_finalize(cgb_handle*, const std::vector<double>&) { return CGB_OK; }
static int salt_launch(cgb_handle* h, double, bool, const cgb::Diag&, void*) { CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "heat+salt not built yet"); }
static int salt_get_diag(cgb_handle* h, const char*, double, double*) { CGB_FAIL(h, CGB_ERR_UNSUPPORTED, "heat+salt not built yet"); }
