// cgb_handle.h -- host-side handle of libcgb200 (shared by the API translation unit and kernel launchers).
#pragma once
#include "../../include/cgb200.h"
#include "cgb_common.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <utility>

struct ParamArray {
    std::vector<double> v;
    int per = CGB_PER_GLOBAL;
};

struct LutTable {
    std::vector<double> H, T;
    int extrap = CGB_EXTRAP_LINE;
};

struct cgb_handle {
    cgb_config cfg{};
    std::string err;
    int N = 0, nl = 0, nvar = 1;
    long long M = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    // host-side description
    std::vector<double> edges;
    std::vector<int> cell_layer;
    std::map<std::string, ParamArray> params;
    std::vector<int> fc_kind, fc_solver;
    std::map<int, LutTable> tables;
    std::vector<int> table_map;
    struct HostForcing { int kind = 0; double value = 0, period = 1, amp = 0, phase = 0, level = 0; std::vector<double> t, y; } forc[2];
    bool dirty = true;        // device copy of the description is stale
    bool have_state = false;
    double t_front = 0.0;     // latest time any advance was asked to reach (host-side bookkeeping for the save path)
    bool fast_ok = false;     // FreeWater fast kernel applicable (set by finalize)
    bool sfcc_fast_ok = false; // SFCC single-class fast kernel applicable (set by finalize -> cgb_sfcc_fast_prepare)
    cgb::SfccFast sf{};
    int sf_threads = 384;
    double* nt_cache = nullptr;       // per-column curve tables of k_heat_newton_tab (cgb_newton_tab_prepare); freed with the description
    bool nt_cache_valid = false;
    // device memory
    std::vector<void*> allocs;
    size_t n_state_allocs = 0;   // allocations made by cgb_create (the rest belong to the description)
    cgb::Dev dev{};
    cgb::SaltDev salt{};
    double* d_scratch = nullptr;      // diag scratch, (N+1)*M doubles x kScratch
    static constexpr int kScratch = 4;
    double* d_lite = nullptr;         // LiteImplicit scratch: 5*N*M doubles
    double* d_stage[2] = {nullptr, nullptr};
    size_t stage_elems = 0;
    double* d_selfull = nullptr;      // full saved fields of one chunk of cgb_solve_saveat_select (stay on the device)
    size_t selfull_elems = 0;
    cudaEvent_t stage_ready[2] = {nullptr, nullptr}, stage_free[2] = {nullptr, nullptr};
    long long launches = 0;
    bool profiling = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    long long lite_iters_total = 0;
};

#define CGB_FAIL(h, code, ...)                                   \
    do {                                                         \
        char _b[512];                                            \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                   \
        (h)->err = _b;                                           \
        return (code);                                           \
    } while (0)

#define CUDA_TRY(h, expr)                                                                           \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) CGB_FAIL(h, CGB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));  \
    } while (0)

// Clears (and, with CGB_DEBUG=1 in the environment, reports) an error left behind by an earlier unchecked call,
// so that a launch check never blames the wrong call.
static inline void clear_stale_error(const char* where)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        static const bool dbg = getenv("CGB_DEBUG") != nullptr;
        if (dbg) fprintf(stderr, "[cgb200] stale CUDA error at %s: %s\n", where, cudaGetErrorString(e));
    }
}

template <typename T>
static int dev_alloc(cgb_handle* h, T** p, size_t n)
{
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) CGB_FAIL(h, CGB_ERR_ALLOC, "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
    h->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return CGB_OK;
}

// Opts the kernel into large dynamic shared memory when needed (per-cell "Heterogeneous" parameter blocks) and returns the
// number of resident CTAs per SM for this launch configuration.
static inline cudaError_t cgb_prepare_kernel(cgb_handle* h, const void* kern, size_t smem, int* blocks_per_sm)
{
    (void)h;
    if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, kern, 128, smem);
    if (e == cudaSuccess && *blocks_per_sm < 1) *blocks_per_sm = 1;
    return e;
}

template <typename T>
static int dev_upload(cgb_handle* h, T** p, const std::vector<T>& v)
{
    int rc = dev_alloc(h, p, v.size());
    if (rc) return rc;
    if (!v.empty()) CUDA_TRY(h, cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    return CGB_OK;
}

