// cgb_launch.h -- kernel launchers, one translation unit per kernel family (compiled in parallel by build.py).
#pragma once
#include "cgb_handle.h"

int cgb_launch_fast(cgb_handle* h, double t_target, double* saveT, double* saveW);                 // cgb_tu_fast.cu
int cgb_launch_fast_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride);
int cgb_fast_max_targets(void);
int cgb_launch_euler(cgb_handle* h, int mode, bool diag, double t_target, const cgb::Diag& D);     // cgb_tu_euler.cu
int cgb_launch_euler_step(cgb_handle* h, int mode, double t_target, const cgb::Diag& D);
int cgb_launch_euler_step_multi(cgb_handle* h, int mode, const double* targets, int n, const cgb::Diag& D, long long save_stride);
int cgb_launch_euler_diag(cgb_handle* h, double t, const cgb::Diag& D);                            // cgb_tu_euler_diag.cu
int cgb_launch_lite_thread(cgb_handle* h, double t, bool diag, const cgb::Diag& D);                // cgb_tu_lite.cu
int cgb_launch_lite_warp(cgb_handle* h, double t_target, double* saveT, double* saveW);            // cgb_tu_lite_warp.cu
int cgb_launch_lite_warp_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride);
int cgb_salt_finalize(cgb_handle* h, const std::vector<double>& zc);                               // cgb_tu_salt.cu
int cgb_salt_launch(cgb_handle* h, double t_target, bool diag, const cgb::Diag& D, void* salt_diag);
int cgb_salt_get_diag(cgb_handle* h, const char* var, double t, double* out);
int cgb_salt_diag_to(cgb_handle* h, const char* var, double t, double* dev_out);   // cell variables only: N*M doubles
int cgb_sfcc_fast_prepare(cgb_handle* h);                                                          // cgb_tu_sfcc_fast.cu (called by finalize)
int cgb_launch_sfcc_fast_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride);
int cgb_sfcc_fast_selftest(cgb_handle* h, const double* T, long long n, double* out);
bool cgb_newton_tab_ok(const cgb_handle* h);
int cgb_newton_tab_prepare(cgb_handle* h);                                                       // cgb_tu_newton_tab.cu
int cgb_launch_newton_tab_multi(cgb_handle* h, const double* targets, int n, double* saveT, double* saveW, long long save_stride);
