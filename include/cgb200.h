/*
 * cgb200.h -- C ABI of libcgb200.so: a B200-native (sm_100a) batched soil-column engine for the
 * CryoGrid.jl hot path (enthalpy heat conduction, FreeWater/SFCC freeze curves, CGEuler and
 * LiteImplicitEuler stepping, coupled heat+salt) over M independent 1-D columns.
 *
 * CryoGrid.jl (v0.23.4) has no FFI for this path -- it is pure Julia.  The entry points below are what a
 * Julia `ccall` binding for that path binds; each cites the reference interface it replaces
 * (file:line relative to the CryoGrid.jl source tree).  See INTEGRATION.md for the Julia glue.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; all reals are double, all indices int32_t/int64_t.
 *   - Every function returns 0 (CGB_OK) or a negative CGB_ERR_* code; cgb_last_error() gives text.
 *     No exception crosses the ABI.  Per-column numerical failures are status bits (cgb_get_status).
 *   - The caller owns every pointer it passes; the library copies in/out before returning and keeps
 *     no host pointer.  The handle owns all device memory.  A handle is not thread-safe.
 *   - State layout: structure-of-arrays, column-fastest: u[(v*N + cell)*M + col] for prognostic
 *     variable v (heat: v=0 is H [J/m^3]; heat+salt: v=0 is T [degC], v=1 is c [mol/m^3]).
 *   - Per-column/per-layer parameter arrays are column-fastest as well: p[layer*M + col].
 *   - There is NO CPU fallback: cgb_create fails with CGB_ERR_NO_DEVICE when no sm_100 GPU is present.
 */
#ifndef CGB200_H
#define CGB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGB_ABI_VERSION 1

/* return codes */
#define CGB_OK               0
#define CGB_ERR_INVALID     -1   /* bad argument / unknown name / call order */
#define CGB_ERR_NO_DEVICE   -2   /* no CUDA device (this library has no CPU path) */
#define CGB_ERR_CUDA        -3   /* CUDA runtime error, see cgb_last_error */
#define CGB_ERR_UNSUPPORTED -4   /* configuration outside what the kernels implement */
#define CGB_ERR_ALLOC       -5

/* physics (what u holds) */
#define CGB_PHYSICS_HEAT      0  /* HeatBalance(:H)  -- Physics/Heat/heat_types.jl:28-33 */
#define CGB_PHYSICS_HEAT_SALT 1  /* SalineGround: HeatBalance(:T)+SaltMassBalance -- Physics/Salt/salt_types.jl:16-23 */
#define CGB_PHYSICS_HEAT_T    2  /* HeatBalance(:T), heat only: u is T [degC], du = dT = dH / dHdT -- Physics/Heat/heat_conduction.jl:147-152,
                                  * Soils/soil_heat.jl:100-113 (SFCC layers only: FreeWater has no temperature form in the reference);
                                  * CGEuler with the CFL limiter (heat_types.jl:17, heat_conduction.jl:162-182) */
/* stepper */
#define CGB_STEPPER_EULER         0  /* CGEuler            -- Solvers/basic_solvers.jl:9-79 */
#define CGB_STEPPER_LITE_IMPLICIT 1  /* LiteImplicitEuler  -- Solvers/LiteImplicit/cglite_types.jl:1-6 */
/* step limiter (explicit stepper) */
#define CGB_LIMITER_MAXDELTA 0       /* Physics/steplimiters.jl:9-31 ; heat default heat_types.jl:18 */
#define CGB_LIMITER_CFL      1       /* Physics/steplimiters.jl:37-40 ; heat_conduction.jl:162-182 */
/* boundary kinds -- traits.jl:8-29 */
#define CGB_BC_DIRICHLET 0
#define CGB_BC_NEUMANN   1
/* which boundary */
#define CGB_TOP    0
#define CGB_BOTTOM 1
/* freeze curves (per layer) */
#define CGB_FC_FREEWATER     0       /* FreeWater: Heat/heat_conduction.jl:18-57 */
#define CGB_FC_DALLAMICO     1       /* FreezeCurves.DallAmico */
#define CGB_FC_PAINTERKARRA  2       /* FreezeCurves.PainterKarra */
#define CGB_FC_DALLAMICOSALT 3       /* FreezeCurves.DallAmicoSalt (heat+salt only) */
/* SFCC inverse-enthalpy solver (per layer) */
#define CGB_SOLVER_LUT    0          /* SFCCPreSolver table lookup (reference default, Soils/ground.jl:27) */
#define CGB_SOLVER_NEWTON 1          /* on-device Newton to 1e-13 K (accuracy mode, per-column SFCC parameters) */
/* LUT extrapolation outside the knots */
#define CGB_EXTRAP_FLAT 0
#define CGB_EXTRAP_LINE 1
/* granularity of a parameter array passed to cgb_set_param */
#define CGB_PER_GLOBAL       0       /* 1 value */
#define CGB_PER_LAYER        1       /* n_layers values */
#define CGB_PER_COLUMN       2       /* M values */
#define CGB_PER_COLUMN_LAYER 3       /* n_layers*M values, p[layer*M + col] */
#define CGB_PER_CELL         4       /* N values (heat+salt porosity profile shared by all columns) */
#define CGB_PER_CELL_COLUMN  5       /* N*M values, p[cell*M + col] (heat+salt porosity profile of every column) */

/* per-column status bits (cgb_get_status) */
#define CGB_STATUS_MAXITERS      1   /* LiteImplicit hit maxiters: cglite_step.jl:74-77 (ReturnCode.MaxIters) */
#define CGB_STATUS_NAN           2   /* non-finite residual: cglite_step.jl:68-70 (error) */
#define CGB_STATUS_FORCING_RANGE 4   /* forcing queried outside its knots: IO/forcings/providers.jl:30-33 (BoundsError) */
#define CGB_STATUS_BAD_DT        8   /* timestep <= 0: Tiles/tile.jl:174 (@assert) */

typedef struct cgb_handle cgb_handle;

/* One POD config for the whole batch -- replaces the constructor kwargs of Tile/HeatBalance/CGEuler/
 * LiteImplicitEuler (basic_solvers.jl:22, integrator.jl:54-62, cglite_types.jl:1-6, heat_types.jl:17-18). */
typedef struct {
    int32_t abi_version;     /* CGB_ABI_VERSION */
    int32_t device;          /* CUDA device ordinal */
    int32_t n_cells;         /* N */
    int32_t n_layers;        /* stratigraphy layers (>=1) */
    int64_t n_columns;       /* M (columns owned by this handle / rank) */
    int32_t physics;         /* CGB_PHYSICS_* */
    int32_t stepper;         /* CGB_STEPPER_* */
    int32_t limiter;         /* CGB_LIMITER_* */
    int32_t top_kind;        /* CGB_BC_* */
    int32_t bot_kind;        /* CGB_BC_* */
    int32_t use_nfactor;     /* TemperatureBC{NFactor}: heat_bc.jl:60-86 */
    double  maxdelta;        /* MaxDelta.Δmax (5e4 J/m^3 for enthalpy) */
    double  courant;         /* CFL.courant_number */
    double  dt0;             /* initial dt: 60 s (CGEuler), 86400 s (LiteImplicit) */
    double  dtmin;           /* opts.dtmin (1 s) */
    double  dtmax;           /* opts.dtmax (3600 s; LiteImplicit: = dt0) */
    int32_t lite_miniters;   /* 2 */
    int32_t lite_maxiters;   /* 1000 */
    double  lite_tol;        /* 1e-3 K */
} cgb_config;

typedef struct {
    int64_t n_columns;
    int64_t n_cells;
    int64_t column_steps;    /* sum over columns of accepted steps since create/reset */
    int64_t cell_steps;      /* column_steps * N  (BASELINE metric numerator) */
    int64_t lite_iterations; /* sum of Picard iterations (LiteImplicit) */
    int64_t kernel_launches; /* CUDA kernels launched by this handle */
    int64_t min_steps, max_steps; /* per-column step count extremes */
    double  min_dt, max_dt;  /* current per-column dt extremes */
    int32_t status_or;       /* OR of all per-column status words */
    int32_t _pad;
} cgb_stats;

/* ---- lifetime ------------------------------------------------------------------------------------- */
int cgb_abi_version(void);
/* Replaces Tile(...)+CryoGridProblem(...)+__init (Tiles/tile.jl:66-98, problem.jl:47-103, basic_solvers.jl:22-52). */
int cgb_create(const cgb_config* cfg, cgb_handle** out);
void cgb_destroy(cgb_handle* h);
const char* cgb_last_error(const cgb_handle* h);   /* h may be NULL: error of the last failed cgb_create */

/* ---- static model description --------------------------------------------------------------------- */
/* Grid(edges): Numerics/grid.jl:17-39 -- one grid shared by all columns. edges has N+1 entries. */
int cgb_set_grid(cgb_handle* h, const double* edges);
/* Stratigraphy -> per-cell layer index (Tiles/stratigraphy.jl; layers are contiguous runs). N entries. */
int cgb_set_layers(cgb_handle* h, const int32_t* cell_layer);
/* Param values (ModelParameters.Param / SimpleSoil fields, Soils/para/simple.jl:8-45; ThermalProperties,
 * Heat/thermal_properties.jl:10-19; NFactor, heat_bc.jl:60-63; SaltProperties, Salt/salt_types.jl:1-4).
 * names: por sat org | vg_alpha vg_n theta_res Tm beta omega | kh_w kh_i kh_a kh_m kh_o ch_w ch_i ch_a ch_m ch_o L |
 *        nf nt top_offset bot_value | tau ds0 benthic_salt salt_dc_max salt_para_por */
int cgb_set_param(cgb_handle* h, const char* name, const double* values, int64_t n, int32_t per);
/* freezecurve(sub) and fcsolver(soil) for one layer: Heat/heat_methods.jl:63, Soils/ground.jl:18-30 */
int cgb_set_freezecurve(cgb_handle* h, int32_t layer, int32_t fc_kind, int32_t solver);
/* Caller-built SFCCPreSolver knots (H -> T), FreezeCurves.initialize! call site Soils/soil_heat.jl:26-36.
 * table_id in [0, 65536). */
int cgb_set_sfcc_table(cgb_handle* h, int32_t table_id, const double* Hk, const double* Tk, int64_t nk, int32_t extrap);
/* FreezeCurves.initialize!(solver, sfcc, hc; sat, θsat) ON THE DEVICE (call site Soils/soil_heat.jl:26-36; default solver
 * SFCCPreSolver(Cache1D; Tmin = -60, errtol = 1e-4), Soils/ground.jl:27): builds the presolver table of every DISTINCT soil class
 * among the (layer, column) pairs that use an SFCC with CGB_SOLVER_LUT -- one thread per class -- and installs the tables and
 * the table map (replacing earlier cgb_set_sfcc_table / cgb_set_sfcc_table_map calls).  Call after all cgb_set_param /
 * cgb_set_freezecurve calls.  At most 65536 classes (fully per-column parameters: use CGB_SOLVER_NEWTON). */
int cgb_build_sfcc_tables(cgb_handle* h, double Tmin, double errtol, double Tmax, int32_t extrap, int32_t* n_tables_out);
/* Read a table back (nk = its length; Hk, Tk of capacity cap may be NULL to query the length only). */
int cgb_get_sfcc_table(cgb_handle* h, int32_t table_id, double* Hk, double* Tk, int64_t cap, int64_t* nk);
/* which table each (layer, column) uses; n_layers*M entries, map[layer*M+col]; NULL => table_id = layer */
int cgb_set_sfcc_table_map(cgb_handle* h, const int32_t* map);
/* Boundary *value* as a function of time (Input(:Tair) via Interpolated1D: IO/forcings/providers.jl:14-33;
 * ConstantBC / PeriodicBC: Physics/simple_bc.jl:14-38).  For CGB_BOTTOM + Neumann the value is what
 * boundaryvalue() returns, i.e. -Qgeo for GeothermalHeatFlux (heat_bc.jl:113); per-column bottom values
 * go through cgb_set_param("bot_value").  The per-column "top_offset" is added to the top value. */
int cgb_set_forcing_table(cgb_handle* h, int32_t which, const double* t, const double* y, int64_t n);
int cgb_set_forcing_periodic(cgb_handle* h, int32_t which, double period, double amplitude, double phase, double level);
int cgb_set_forcing_constant(cgb_handle* h, int32_t which, double value);

/* ---- state ---------------------------------------------------------------------------------------- */
/* u0 and t0 (prob.u0, prob.tspan[1]); resets every column's t=t0, dt=dt0, counters and status. */
int cgb_set_state(cgb_handle* h, const double* u, double t0);
/* initialcondition!(tile, tspan) evaluated ON THE DEVICE (Tiles/tile.jl:187-200), so that remade ensemble members do not
 * go through a serial host loop (problem.jl:117-119).  Both reset t = t0, dt = dt0, counters and status like cgb_set_state.
 *  - InterpInitializer(:T, profile) + initialize_soil_heat! (initializers.jl:42-67, Soils/soil_heat.jl:38-63):
 *    zk[nk] depths, vk[nk] (per = CGB_PER_GLOBAL) or vk[k*M + col] (per = CGB_PER_COLUMN); linear in depth, flat outside.
 *  - ThermalSteadyStateInit(T0, Qgeo, maxiters, thresh) (Heat/heat_init.jl:58-85); T0 has 1 (GLOBAL) or M (COLUMN) values. */
int cgb_init_interp(cgb_handle* h, const double* zk, const double* vk, int64_t nk, int32_t per, double t0);
int cgb_init_steadystate(cgb_handle* h, const double* T0, int32_t per, double Qgeo, int32_t maxiters, double thresh, double t0);
/* u, and optionally per-column t and dt (either may be NULL). This is also the checkpoint. */
int cgb_get_state(cgb_handle* h, double* u, double* t, double* dt);
/* Checkpoint restore: the counterpart of cgb_get_state.  u as in cgb_set_state, t[M] and dt[M] the per-column time and NEXT
 * step size a checkpoint returned (integrator.t / integrator.dt, Solvers/integrator.jl:40-52); steps[M] may be NULL (counters
 * restart at 0).  Status words are cleared.  A run continued from a checkpoint takes exactly the steps the uninterrupted run
 * takes (tests/test_restart_parity.py). */
int cgb_restore_state(cgb_handle* h, const double* u, const double* t, const double* dt, const int64_t* steps);
/* per-column status words / step counts (M entries each, either may be NULL) */
int cgb_get_status(cgb_handle* h, int32_t* status, int64_t* steps);
int cgb_get_stats(cgb_handle* h, cgb_stats* out);

/* ---- the hot path --------------------------------------------------------------------------------- */
/* ONE right-hand-side evaluation at the current state: tile(du,u,p,t) -- Tiles/tile_base.jl:21,
 * Tiles/tile.jl:136-158.  du_out: same layout as u (host pointer). For LiteImplicit du is 0 by
 * construction (heat_implicit.jl:106); query DT_ap etc. with cgb_get_diag. */
int cgb_rhs(cgb_handle* h, double t, double* du_out);
/* Integrate every column with its own adaptive dt until t_target is hit exactly:
 * step!/perform_step!/saveat! -- Solvers/integrator.jl:84-90,111-132; basic_solvers.jl:54-79;
 * LiteImplicit/cglite_step.jl:1-99.  Asynchronous w.r.t. the host; any cgb_get_* synchronises. */
int cgb_advance_to(cgb_handle* h, double t_target);
int cgb_synchronize(cgb_handle* h);
/* Diagnostic variable recomputed from the current state at time t (getvar/getvars: Tiles/tile.jl:251-262,
 * Numerics/statevars.jl:81-106).  var: T thw C dHdT dthwdT kc (N*M) | k jH (N+1)*M | dH dT Henth | DT_an DT_as DT_ap DT_bp
 * (heat+salt adds: H dT dc ds jc dthwdc).  out is a host pointer. */
int cgb_get_diag(cgb_handle* h, const char* var, double t, double* out);
/* solve(prob; saveat): advance through tsave[0..nsave) and write the saved variables
 * out[((s*nvars + v)*N + cell)*M + col] (savefunc/saveat!: problem.jl:67-68, integrator.jl:111-132).
 * vars: comma-separated subset of "H,T,thw,T_now,thw_now".  out is a HOST pointer (pinned or pageable); copies are
 * overlapped with stepping on a second stream.
 * REFERENCE-VERBATIM SEMANTICS of the diagnostic variables T and thw: CryoGrid.jl's savefunc returns the *cached* diagnostic
 * arrays of the last RHS evaluation (getvars -> retrieve, Numerics/statevars.jl:60-106), and CGEuler evaluates the tile
 * BEFORE it applies u += dt*du (basic_solvers.jl:61-66); LiteImplicit's last evaluation is the start of its last Picard
 * iteration (cglite_step.jl:45-62).  So the reference's saved T, θw belong to the state one step (one iteration) before the
 * saved H.  "T"/"thw" reproduce exactly that (the stepping kernel writes them on the last step of each interval, no extra
 * launch); "T_now"/"thw_now" are recomputed from the saved state itself; "H" is the prognostic state at the save time.
 * When only "T"/"thw" are requested and tsave is strictly increasing, up to 8 save times share one kernel launch (a column
 * stays in registers across them; CGEuler, all freeze curves, and LiteImplicitEuler); otherwise one launch per save time.
 * With the on-device Newton solver the two schedules agree to the solver tolerance (1e-9 K), not bitwise: a running column
 * carries its Newton seed, a fresh launch restarts it from H/c_w (soil_heat.jl:127). */
int cgb_solve_saveat(cgb_handle* h, const double* tsave, int64_t nsave, const char* vars, double* out);

/* Device-side output path: like cgb_solve_saveat, but only the cells cell_index[0..nsel) are kept --
 * out[((s*nvars + v)*nsel + k)*M + col] -- which is what users read from CryoGridOutput (out.T[Z(Near(zs))],
 * examples/cglite_parameter_ensembles.jl:96) -- and, if thawdepth_out != NULL, the sub-grid thaw depth of every column at
 * every save, thawdepth_out[s*M + col] (Diagnostics.thawdepth, Diagnostics/permafrost.jl:38-59; its yearly maximum is the
 * active layer thickness, :66-72).  nsel may be 0 to return thaw depths only.  Host pointers. */
int cgb_solve_saveat_select(cgb_handle* h, const double* tsave, int64_t nsave, const char* vars, const int32_t* cell_index,
                            int32_t nsel, double* out, double* thawdepth_out);

/* Permafrost diagnostics reduced ON THE DEVICE over the saves tsave[0..nsave) of this call -- one "year" of the reference's annual
 * collapse (Diagnostics/permafrost.jl); only M-sized results leave the GPU.  All outputs are host pointers and may be NULL:
 *   alt[M]       active_layer_thickness  = max over the saves of Diagnostics.thawdepth        (permafrost.jl:38-72)
 *   pf_table[M]  permafrosttable         = first cell depth whose window-max T < Tmelt        (:80-86)
 *   pf_base[M]   permafrostbase          (verbatim index arithmetic)                           (:95-101)
 *   zaa[M]       zero_annual_amplitude   = depth where max - min of T drops below threshold    (:7-30; Inf if never)
 *   magt[nsave*M] mean_annual_ground_temperature = depth-mean of T between magt_upper and magt_lower at every save (:108-113)
 * T is the reference's saved T (cache of the last RHS evaluation of each interval, see cgb_solve_saveat). */
int cgb_solve_saveat_permafrost(cgb_handle* h, const double* tsave, int64_t nsave, double Tmelt, double zaa_threshold, double magt_upper,
                                double magt_lower, double* alt, double* pf_table, double* pf_base, double* zaa, double* magt);

/* ---- ensemble statistics / multi-GPU seams -------------------------------------------------------- */
/* Local (this handle's columns) sums for ensemble moments of a diagnostic at the current state:
 * sum[cell] = Σ_col x, sumsq[cell] = Σ_col x^2 (N entries each, host pointers).  Ranks all-reduce these
 * (NCCL via torch.distributed) -- the reference's analogue is the EnsembleProblem reduction
 * (Solvers/DiffEq/ensemble.jl:33-45). */
int cgb_ensemble_partial(cgb_handle* h, const char* var, double t, double* sum, double* sumsq);
/* Raw device pointers (for zero-copy use by the host framework: NCCL gathers, torch tensors).
 * Valid until cgb_destroy. */
int cgb_device_state(cgb_handle* h, void** u_dev, void** t_dev, void** dt_dev);
/* Device-side variant of cgb_ensemble_partial: writes 2*N doubles (sum then sumsq) to a DEVICE buffer. */
int cgb_ensemble_partial_device(cgb_handle* h, const char* var, double t, void* out_dev);

/* ---- measurement ----------------------------------------------------------------------------------- */
/* CUDA-event timing of the stepping kernels on the handle's own stream (torch.cuda.Event only sees torch's
 * current stream).  While enabled, every cgb_advance_to / cgb_solve_saveat stepping launch is bracketed by an
 * event pair; cgb_profile_read synchronises, returns the number of launches and their summed duration since
 * the last read, and resets. */
int cgb_profile_enable(cgb_handle* h, int32_t on);
int cgb_profile_read(cgb_handle* h, int64_t* n_launches, double* total_ms);
/* Self-test: evaluates on the device the reciprocal sequences the kernels use in place of IEEE division
 * (raw MUFU.RCP64H seed, seed + 2 Newton steps, seed + 1 cubic step) for n host inputs. */
int cgb_selftest_rcp(const double* x, int64_t n, double* out_raw, double* out_newton2, double* out_cubic);
/* Self-test: the kernels' exp / log replacements (van Genuchten powers, FreezeCurves.jl closed forms):
 * out_exp[i] = exp_fast(x[i]) (argument clamped to [-708, 709]), out_log[i] = log_pos(|x[i]|) (normal range). */
int cgb_selftest_explog(const double* x, int64_t n, double* out_exp, double* out_log);
/* Same for the table-driven variants the STEPPING kernels use (64-entry shared-memory tables, exp_tab / log_tab). */
int cgb_selftest_explog_tab(const double* x, int64_t n, double* out_exp, double* out_log);

/* Self-test of the SFCC single-class fast path (many columns of ONE soil class through the presolver LUT, MaxDelta limiter:
 * csrc/cgb_euler_sfcc_fast.cuh): thw_out[i] = θw(T[i]) evaluated from the shared-memory curve table exactly as the stepping
 * kernel does (degree-7 piecewise polynomials on log-spaced bins of T* - T, replacing the two pow() of the closed form that
 * Soils/soil_heat.jl:131-139 evaluates after the table lookup).  CGB_ERR_UNSUPPORTED when the handle does not qualify. */
int cgb_selftest_sfcc_curve(cgb_handle* h, const double* T, int64_t n, double* thw_out);

#ifdef __cplusplus
}
#endif
#endif /* CGB200_H */
