/*
 * cg_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, double-precision, single-column restatement of the CryoGrid.jl v0.23.4 hot path
 * (enthalpy heat conduction, FreeWater / SFCC freeze curves, CGEuler, LiteImplicitEuler,
 * coupled heat+salt).  Each function cites the reference file:line it follows (paths relative
 * to /root/reference/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the shipped engine (libcgb200.so) never does.
 *
 * PARITY PINNING
 *   - Stencils, Thomas solver, grids, boundary fluxes, n-factors, forcing interpolation, the
 *     explicit Fourier test and the LiteImplicit periodic/Stefan tests are pinned against the
 *     reference's own golden vectors (tests/test_oracle_golden.py).
 *   - "parity unpinned": everything behind the FreezeCurves.jl (compat 0.9, un-vendored,
 *     Project.toml:18,61) boundary -- freewater(), the SFCC closed forms (DallAmico,
 *     PainterKarra, DallAmicoSalt, VanGenuchten), the SFCCPreSolver table builder and
 *     sfccsolve().  The reference tests hold no numeric value for them and Julia cannot run in
 *     this container, so those functions restate the published FreezeCurves.jl algorithms from
 *     its documentation/papers (Dall'Amico et al. 2011; Painter & Karra 2014; van Genuchten 1980).
 */
#ifndef CG_ORACLE_H
#define CG_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* freeze curve kinds */
#define CG_FC_FREEWATER     0
#define CG_FC_DALLAMICO     1
#define CG_FC_PAINTERKARRA  2
#define CG_FC_DALLAMICOSALT 3
/* SFCC solver kinds */
#define CG_SOLVER_PRESOLVER 0   /* SFCCPreSolver LUT (reference default, Soils/ground.jl:27) */
#define CG_SOLVER_NEWTON    1   /* tightly converged Newton (accuracy mode) */
/* LUT extrapolation */
#define CG_EXTRAP_FLAT 0
#define CG_EXTRAP_LINE 1
/* boundary kinds (traits.jl:8-29) */
#define CG_BC_DIRICHLET 0
#define CG_BC_NEUMANN   1
/* forcing kinds */
#define CG_FORCING_CONSTANT 0   /* ConstantBC, simple_bc.jl:14-22 */
#define CG_FORCING_PERIODIC 1   /* PeriodicBC, simple_bc.jl:24-38 */
#define CG_FORCING_TABLE    2   /* Interpolated1D, IO/forcings/providers.jl:14-33 */
/* step limiters (steplimiters.jl) */
#define CG_LIMITER_MAXDELTA 0
#define CG_LIMITER_CFL      1
/* heat operator */
#define CG_OP_ENTHALPY 0          /* Diffusion1D(:H) */
#define CG_OP_ENTHALPY_IMPLICIT 1 /* EnthalpyImplicit */
#define CG_OP_TEMPERATURE 2       /* Diffusion1D(:T): prognostic T, dT = dH / dHdT (heat_conduction.jl:147-152) */

/* status bits returned by the steppers */
#define CG_STATUS_OK        0
#define CG_STATUS_MAXITERS  1   /* cglite_step.jl:74-77 */
#define CG_STATUS_NAN       2   /* cglite_step.jl:68-70 */
#define CG_STATUS_FORCING_RANGE 4 /* Interpolated1D BoundsError, test/IO/forcing_tests.jl:13-14 */
#define CG_STATUS_BAD_DT    8   /* tile.jl:174 @assert dtmax > 0 */

typedef struct {
    double kh_w, kh_i, kh_a, kh_m, kh_o;   /* thermal_properties.jl:13-15, Soils/para/simple.jl:37-38 */
    double ch_w, ch_i, ch_a, ch_m, ch_o;   /* thermal_properties.jl:16-18, simple.jl:42-43 */
    double L;                              /* rho_w*Lsl, heat_types.jl:65-69 */
} cg_thermal;

typedef struct {
    int32_t kind;       /* CG_FC_* */
    int32_t solver;     /* CG_SOLVER_* */
    double alpha, n, theta_res;        /* van Genuchten */
    double Tm;          /* melting point, degC (FreezeCurves SoilFreezeThawProperties) */
    double beta, omega; /* PainterKarra */
    double g, Lsl, R, rho_w;
    /* presolver table (knots of the H -> T interpolant) */
    int32_t nk;
    int32_t extrap;     /* CG_EXTRAP_* */
    const double* Hk;
    const double* Tk;
} cg_freezecurve;

typedef struct {
    int32_t kind;
    int32_t n;                                  /* table length */
    double value;                               /* constant */
    double period, amplitude, phase, level;     /* periodic */
    const double* t;                            /* table knots (seconds) */
    const double* y;
} cg_forcing;

typedef struct {
    int32_t n;               /* number of cells */
    int32_t n_layers;
    const double* edges;     /* n+1 */
    const int32_t* cell_layer; /* n */
    const double* por;       /* per layer */
    const double* sat;
    const double* org;
    const cg_freezecurve* fc;/* per layer */
    cg_thermal tp;
    int32_t op;              /* CG_OP_* */
    int32_t top_kind;        /* CG_BC_* */
    int32_t use_nfactor;
    int32_t bot_kind;
    cg_forcing top;          /* boundary *value* (temperature or flux) */
    cg_forcing bot;          /* boundary *value*: GeothermalHeatFlux passes -Qgeo (heat_bc.jl:113) */
    double nf, nt;           /* NFactor, heat_bc.jl:60-86 */
    int32_t limiter;         /* CG_LIMITER_* */
    int32_t _pad;
    double maxdelta;         /* MaxDelta.Δmax (5e4 for enthalpy, heat_types.jl:18) */
    double courant;          /* CFL.courant_number */
} cg_column;

/* diagnostic/work arrays for one column (all caller-allocated) */
typedef struct {
    double *T, *thw, *C, *dHdT, *dthwdT, *kc;  /* n   */
    double *k, *jH;                            /* n+1 */
    double *dH;                                /* n   */
    double *an, *as, *ap, *bp;                 /* n   (implicit coefficients) */
    double *zc, *dk, *dxc;                     /* grid: midpoints n, cell sizes n, midpoint distances n-1 */
    double *dT;                                /* n   (temperature form: dT = dH / dHdT) */
} cg_work;

/* ---- Numerics (src/Numerics/grid.jl, math.jl) ---- */
void   cg_grid(const double* edges, int n_edges, double* cells, double* d_edges, double* d_cells);
double cg_flux(double x1, double x2, double dx, double k);
double cg_divergence(double j1, double j2, double dj);
void   cg_flux_vec(double* j, const double* x, const double* dx, const double* k, int n);
void   cg_divergence_vec(double* dx, const double* j, const double* dj, int n);
void   cg_nonlineardiffusion(double* dx, double* j, const double* x, const double* Dx, const double* k, const double* Dk, int n);
double cg_harmonicmean(double x1, double x2, double w1, double w2);
void   cg_tdma_solve(double* x, double* a, double* b, double* c, double* d, int n);

/* ---- freeze curves ---- */
void   cg_freewater(double H, double thwi, double L, double* thw);
double cg_enthalpyinv_freewater(double H, double thwi, double C, double L);
double cg_vg_theta(double psi, double thsat, double thres, double alpha, double n, double* dth_dpsi);
double cg_vg_psi(double th, double thsat, double thres, double alpha, double n);
double cg_sfcc(const cg_freezecurve* fc, double T, double sat, double thsat, double saltconc,
               double* dthw_dT, double* dthw_dc);
double cg_sfcc_heatcap(const cg_thermal* tp, double por, double org, double thw, double thwi, double thsat);
int    cg_presolver_build(const cg_freezecurve* fc, const cg_thermal* tp, double por, double sat, double org,
                          double Tmin, double errtol, double Tmax, double* Hk, double* Tk, int cap);
double cg_lut_eval(const double* Hk, const double* Tk, int nk, int extrap, double H);
double cg_sfcc_newton(const cg_freezecurve* fc, const cg_thermal* tp, double por, double sat, double org,
                      double H, double T0);

/* ---- forcing / BCs ---- */
double cg_forcing_eval(const cg_forcing* f, double t, int* status);
double cg_nfactor(double Tair, double nf, double nt);

/* ---- column RHS ---- */
void   cg_work_grid(const cg_column* col, cg_work* w);
void   cg_freezethaw(const cg_column* col, const double* H, cg_work* w);
void   cg_thermalconductivity(const cg_column* col, cg_work* w);
int    cg_rhs(const cg_column* col, const double* H, double t, cg_work* w);
double cg_timestep(const cg_column* col, const double* H, const cg_work* w);

/* ---- steppers ---- */
int cg_cgeuler_advance(const cg_column* col, double* H, double* t, double* dt, double t_target,
                       double dtmin, double dtmax, int64_t* nsteps, cg_work* w);
int cg_cgeuler_advance_trace(const cg_column* col, double* H, double* t, double* dt, double t_target,
                             double dtmin, double dtmax, int64_t* nsteps, cg_work* w,
                             double* dt_trace, int64_t cap, int64_t* nlimited);
int cg_lite_step(const cg_column* col, double* H, double* t, double dt, int miniters, int maxiters,
                 double tol, int64_t* iters, cg_work* w);
int cg_lite_advance(const cg_column* col, double* H, double* t, double* dt, double t_target,
                    int miniters, int maxiters, double tol, int64_t* nsteps, int64_t* iters, cg_work* w);

/* ---- initial conditions ---- */
void cg_interp_profile(const double* zk, const double* vk, int nk, const double* z, int n, double* out);
void cg_init_soil_heat(const cg_column* col, const double* T, double* H, cg_work* w);
void cg_steadystate_init(const cg_column* col, double T0, double Qgeo, int maxiters, double thresh,
                         double* H, cg_work* w);

/* ---- heat + salt (src/Physics/Salt) ---- */
typedef struct {
    int32_t n;
    int32_t surface_state;   /* SaltGradient.surfaceState (0 = inundated) */
    const double* edges;     /* n+1 */
    const double* por;       /* per cell (SedimentCompactionInitializer) */
    double sat, org;
    cg_freezecurve fc;       /* DallAmicoSalt */
    cg_thermal tp;
    double tau, ds0;         /* SaltProperties, salt_types.jl:1-4 */
    double T_ub;             /* ConstantTemperature(0.0) */
    double benthic_salt;     /* SaltGradient.benthicSalt */
    double bot_value;        /* -Qgeo */
    double courant, dc_max;  /* CFL(0.5, MaxDelta(5e-3)), salt_types.jl:13 */
    double heat_courant;     /* heat CFL(0.5, MaxDelta(Inf)), heat_types.jl:17 */
    int32_t use_top_forcing; /* 1: the upper boundary temperature is top(t) (TemperatureBC with an Input / PeriodicBC), else T_ub */
    int32_t _pad;
    cg_forcing top;
} cg_salt_column;

typedef struct {
    double *thw, *dthwdT, *dthwdc, *C, *dHdT, *H, *kc, *ds_mid;  /* n */
    double *k, *ds, *jH, *jc;                                    /* n+1 */
    double *dH, *dc_F, *dT, *dc;                                 /* n */
    double *zc, *dk, *dxc;
} cg_salt_work;

void   cg_compaction(const double* edges, int n, double por0, double por_min, double* por);
void   cg_salt_rhs(const cg_salt_column* col, const double* T, const double* c, cg_salt_work* w);
int    cg_salt_rhs_t(const cg_salt_column* col, const double* T, const double* c, double t, cg_salt_work* w);
double cg_salt_timestep(const cg_salt_column* col, const cg_salt_work* w);
int    cg_salt_advance(const cg_salt_column* col, double* T, double* c, double* t, double* dt, double t_target,
                       double dtmin, double dtmax, int64_t* nsteps, cg_salt_work* w);

int    cg_salt_advance_trace(const cg_salt_column* col, double* T, double* c, double* t, double* dt, double t_target,
                             double dtmin, double dtmax, int64_t* nsteps, cg_salt_work* w, double* dt_trace, int64_t cap);

/* ---- batch drivers (OpenMP over columns) for the CPU baseline ---- */
int  cg_cgeuler_advance_batch(const cg_column* cols, int64_t M, double* H /* [M][n] */, double* t, double* dt,
                              double t_target, double dtmin, double dtmax, int64_t* nsteps, int nthreads);
int  cg_lite_advance_batch(const cg_column* cols, int64_t M, double* H, double* t, double* dt, double t_target,
                           int miniters, int maxiters, double tol, int64_t* nsteps, int64_t* iters, int nthreads);
int  cg_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
