// cgb_init.cuh -- on-device initial conditions: the reference re-runs initialcondition! on the host
// for every remade ensemble member (problem.jl:117-119); at 1e5-1e6 members that serial step dominates once stepping
// is fast.  Two initialisers:
//   * InterpInitializer(:T, profile) + initialize_soil_heat!   (initializers.jl:42-67, Soils/soil_heat.jl:38-63)
//   * ThermalSteadyStateInit(T0, Qgeo, maxiters, thresh)       (Heat/heat_init.jl:58-85), layer by layer
#pragma once
#include "cgb_handle.h"
#include "cgb_euler.cuh"

namespace cgb {

// θw, C from T (initialize_soil_heat!): FreeWater θw = T > 0 ? θwi : 0; SFCC θw = sfcc(T, sat; θsat = por);
// C = heatcapacity(soil, state, i) (the REGULAR mixing rule, not sfccheatcap); H = T*C + L*θw (heat_methods.jl:93).
__device__ __forceinline__ double enthalpy_from_T(const Dev& P, const double* lc, double T)
{
    int kind = reinterpret_cast<const int*>(lc + LC_KIND)[0] & 0xff;
    double thw;
    if (kind == 0) thw = (T > 0.0) ? lc[LC_THWI] : 0.0;
    else { double d; thw = sfcc_theta(lc, T, d); }
    double C = fma(P.ch_w - P.ch_i, thw, lc[LC_CB]);
    return T * C + P.L * thw;
}

// One thread per (cell, column): T = interp(profile)(z_cell) -> H.   vk is [nk] (per_column = 0) or [nk*M].
__global__ void __launch_bounds__(256) k_init_interp(const __grid_constant__ Dev P, const double* __restrict__ zc, const double* __restrict__ zk,
                                                     const double* __restrict__ vk, int nk, int per_column)
{
    const long long M = P.M;
    long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int cell = blockIdx.y;
    if (col >= M) return;
    double z = zc[cell];
    double T;
    auto V = [&](int k) { return per_column ? vk[(size_t)k * M + col] : vk[k]; };
    if (nk == 1) T = V(0);
    else if (z <= zk[0]) T = V(0);                      // Flat extrapolation (initializers.jl:49)
    else if (z >= zk[nk - 1]) T = V(nk - 1);
    else {
        int lo = 0, hi = nk - 1;
        while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (zk[mid] <= z) lo = mid; else hi = mid; }
        double w = (z - zk[lo]) / (zk[lo + 1] - zk[lo]);
        T = (1.0 - w) * V(lo) + w * V(lo + 1);
    }
    if (P.tform) { P.u[(size_t)cell * M + col] = T; return; }      // HeatBalance(:T): the prognostic state is T itself
    double lc[LC_STRIDE];
    layer_consts(P, col, P.cell_layer[cell], lc);
    P.u[(size_t)cell * M + col] = enthalpy_from_T(P, lc, T);
}

// One thread per column; T and kc live in a global scratch ([N*M] each, coalesced across the threads of a warp).
__global__ void __launch_bounds__(128) k_init_steadystate(const __grid_constant__ Dev P, const double* __restrict__ zc, const double* __restrict__ T0,
                                                          int per_column, double Qgeo, int maxiters, double thresh, double* __restrict__ Ts,
                                                          double* __restrict__ kcs)
{
    const long long M = P.M;
    const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (col >= M) return;
    const int N = P.N;
    const double dC = P.ch_w - P.ch_i, dK = P.sk_w - P.sk_i;
    double lc[LC_STRIDE];
    for (int i = 0; i < N; ++i) { Ts[(size_t)i * M + col] = 0.0; kcs[(size_t)i * M + col] = 0.0; }
    double Tup = per_column ? T0[col] : T0[0];
    int start = 0;
    while (start < N) {
        const int l = P.cell_layer[start];
        int end = start;
        while (end < N && P.cell_layer[end] == l) ++end;
        layer_consts(P, col, l, lc);
        // apply T_new on the layer: H from T (initialcondition!), then T, kc back from H (computediagnostic!)
        auto sweep = [&](bool measure) -> double {
            double dT = 0.0;
            for (int i = start; i < end; ++i) {
                size_t o = (size_t)i * M + col;
                double Tn = Tup + zc[i] * Qgeo / kcs[o];
                if (measure) {
                    double d = fabs(Tn - Ts[o]);
                    dT = (isnan(d) || isnan(dT)) ? nan("") : fmax(dT, d);      // Julia maximum() propagates NaN
                }
                double H = enthalpy_from_T(P, lc, Tn);
                double T = Tn, thw, C, dHdT, dth, kc;
                freezethaw_body<0, true>(P, lc, H, dC, dK, T, thw, C, dHdT, dth, kc);
                Ts[o] = T; kcs[o] = kc;
                P.u[o] = H;
            }
            return dT;
        };
        int it = 1;
        double dT = INFINITY;
        while (it < maxiters && dT > thresh) { dT = sweep(true); ++it; }       // heat_init.jl:73-81
        sweep(false);                                                          // heat_init.jl:82-86
        Tup = Ts[(size_t)(end - 1) * M + col];                                 // heat_init.jl:65
        start = end;
    }
    // stratigraphy.jl:131-141: the per-layer initialcondition! runs once more on the final temperatures
    int cur = -1;
    for (int i = 0; i < N; ++i) {
        int l = P.cell_layer[i];
        if (l != cur) { layer_consts(P, col, l, lc); cur = l; }
        size_t o = (size_t)i * M + col;
        P.u[o] = enthalpy_from_T(P, lc, Ts[o]);
    }
}

} // namespace cgb
