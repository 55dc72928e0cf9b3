/*
 * cg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code). See cg_oracle.h.
 *
 * Literal, unoptimised restatement of the CryoGrid.jl v0.23.4 hot path for ONE column in
 * double precision, in the reference's own operation order.  Citations are file:line relative
 * to /root/reference/.  Functions behind the FreezeCurves.jl boundary are tagged
 * [FreezeCurves 0.9, parity unpinned].
 */
#define _GNU_SOURCE
#include "cg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------
 * Numerics
 * ---------------------------------------------------------------------------------------- */

/* Grid constructor, src/Numerics/grid.jl:24-33: cells = (e[i]+e[i+1])/2, Δedges, Δcells. */
void cg_grid(const double* edges, int n_edges, double* cells, double* d_edges, double* d_cells)
{
    int nc = n_edges - 1;
    for (int i = 0; i < nc; ++i) cells[i] = (edges[i] + edges[i + 1]) / 2.0;
    for (int i = 0; i < nc; ++i) d_edges[i] = edges[i + 1] - edges[i];
    for (int i = 0; i < nc - 1; ++i) d_cells[i] = cells[i + 1] - cells[i];
}

/* src/Numerics/math.jl:41  flux(x1,x2,Δx,k) = -k*(x2-x1)/Δx  (positive downward) */
double cg_flux(double x1, double x2, double dx, double k) { return -k * (x2 - x1) / dx; }

/* src/Numerics/math.jl:68  divergence(j1,j2,Δj) = (j1-j2)/Δj */
double cg_divergence(double j1, double j2, double dj) { return (j1 - j2) / dj; }

/* src/Numerics/math.jl:57-64  flux!(j,x,Δx,k): interior faces j[2:end-1] += flux(x[i-1],x[i],Δx[i-1],k[i]).
 * n = length(x); j,k have n+1 entries; dx has n-1. */
void cg_flux_vec(double* j, const double* x, const double* dx, const double* k, int n)
{
    for (int i = 1; i < n; ++i) j[i] += cg_flux(x[i - 1], x[i], dx[i - 1], k[i]);
}

/* src/Numerics/math.jl:82-87  divergence!(dx,j,Δj): dx[i] += (j[i]-j[i+1])/Δj[i] */
void cg_divergence_vec(double* dx, const double* j, const double* dj, int n)
{
    for (int i = 0; i < n; ++i) dx[i] += cg_divergence(j[i], j[i + 1], dj[i]);
}

/* src/Numerics/math.jl:91-123  nonlineardiffusion! (single pass with loop-carried j) */
void cg_nonlineardiffusion(double* dx, double* j, const double* x, const double* Dx, const double* k,
                           const double* Dk, int n)
{
    /* Julia: for i in 2:length(j)-1 (1-based), length(j) = n+1 */
    for (int i = 1; i < n; ++i) {
        double j2 = cg_flux(x[i - 1], x[i], Dx[i - 1], k[i]);
        double div = cg_divergence(j[i - 1], j2, Dk[i - 1]);
        j[i] += j2;
        dx[i - 1] += div;
    }
    dx[n - 1] += cg_divergence(j[n - 1], j[n], Dk[n - 1]);
}

/* src/Numerics/math.jl:141  harmonicmean(x1,x2,w1,w2) = (w1+w2)/(w1*x1^-1 + w2*x2^-1) */
double cg_harmonicmean(double x1, double x2, double w1, double w2)
{
    return (w1 + w2) / (w1 * (1.0 / x1) + w2 * (1.0 / x2));
}

/* src/Numerics/math.jl:169-197  Thomas algorithm; a,c have n-1 entries; c,d are overwritten. */
void cg_tdma_solve(double* x, double* a, double* b, double* c, double* d, int n)
{
    (void)a;
    c[0] = c[0] / b[0];
    d[0] = d[0] / b[0];
    for (int i = 1; i < n - 1; ++i) {
        double temp = b[i] - a[i - 1] * c[i - 1];
        c[i] = c[i] / temp;
        d[i] = (d[i] - a[i - 1] * d[i - 1]) / temp;
    }
    d[n - 1] = (d[n - 1] - a[n - 2] * d[n - 2]) / (b[n - 1] - a[n - 2] * c[n - 2]);
    x[n - 1] = d[n - 1];
    for (int i = n - 2; i >= 0; --i) x[i] = d[i] - c[i] * x[i + 1];
}

/* ------------------------------------------------------------------------------------------
 * Soil composition and thermal properties
 * ---------------------------------------------------------------------------------------- */

/* Soils/para/simple.jl:19-24 soilcomponent */
static inline double soil_thwi(double por, double sat) { return por * sat; }
static inline double soil_thm(double por, double org) { return (1.0 - por) * (1.0 - org); }
static inline double soil_tho(double por, double org) { return (1.0 - por) * org; }

/* Soils/soil_heat.jl:76-85 soil_volumetricfractions + Heat/heat_methods.jl:52-57 heatcapacity:
 * C = linearmean(cs, θs) = sum(map(*, cs, θs)) (Numerics/math.jl:162), summed left to right. */
static inline double soil_heatcapacity(const cg_thermal* tp, double thw, double thwi, double thm, double tho)
{
    double tha = 1.0 - thwi - thm - tho;
    double thi = thwi - thw;
    return tp->ch_w * thw + tp->ch_i * thi + tp->ch_a * tha + tp->ch_m * thm + tp->ch_o * tho;
}

/* Heat/heat_methods.jl:40-45 + thermal_properties.jl:37: kc = (Σ sqrt(k_i) θ_i)^2 */
static inline double soil_thermalconductivity(const cg_thermal* tp, double thw, double thwi, double thm, double tho)
{
    double tha = 1.0 - thwi - thm - tho;
    double thi = thwi - thw;
    double s = sqrt(tp->kh_w) * thw + sqrt(tp->kh_i) * thi + sqrt(tp->kh_a) * tha + sqrt(tp->kh_m) * thm +
               sqrt(tp->kh_o) * tho;
    return s * s;
}

/* ------------------------------------------------------------------------------------------
 * Free water freeze curve
 * ---------------------------------------------------------------------------------------- */

/* [FreezeCurves 0.9, parity unpinned] freewater(H, θtot, L), call site Heat/heat_conduction.jl:23:
 * θtot floored at 1e-8; θw = (I_c*(H/Lθ) + I_t)*θtot with I_t = H > Lθ, I_c = 0 < H <= Lθ. */
void cg_freewater(double H, double thwi, double L, double* thw)
{
    double thtot = thwi > 1e-8 ? thwi : 1e-8;
    double Lth = L * thtot;
    int I_t = (H > Lth);
    int I_c = (H > 0.0 && H <= Lth);
    /* Julia Bool*Float is a "strong zero" (false*Inf == 0.0), hence the selects instead of products */
    *thw = ((I_c ? (H / Lth) : 0.0) + (I_t ? 1.0 : 0.0)) * thtot;
}

/* Heat/heat_conduction.jl:39-57 enthalpyinv(::FreeWater, H, θwi, C, L) */
double cg_enthalpyinv_freewater(double H, double thwi, double C, double L)
{
    double Lth = L * thwi;
    double T_f = H / C;
    double T_t = (H - Lth) / C;
    return (H < 0.0) ? T_f : ((H >= Lth) ? T_t : 0.0);
}

/* ------------------------------------------------------------------------------------------
 * SFCC closed forms  [FreezeCurves 0.9, parity unpinned]
 * ---------------------------------------------------------------------------------------- */

/* van Genuchten (1980): θ(ψ) = θres + (θsat-θres)(1+|αψ|^n)^-(1-1/n), ψ<=0; θsat otherwise.
 * Also returns dθ/dψ (what ForwardDiff produces for the same expression). */
double cg_vg_theta(double psi, double thsat, double thres, double alpha, double n, double* dth_dpsi)
{
    double m = 1.0 - 1.0 / n;
    if (psi <= 0.0) {
        double x = fabs(-alpha * psi);
        double xn = pow(x, n);
        double base = 1.0 + xn;
        double th = thres + (thsat - thres) * pow(base, -m);
        if (dth_dpsi) {
            /* d/dψ |αψ|^n = n|αψ|^(n-1) * α * sign(ψ)(-1)... for ψ<0: d|αψ|/dψ = -α */
            double dxn = (x > 0.0) ? n * pow(x, n - 1.0) * (-alpha) : 0.0;
            *dth_dpsi = (thsat - thres) * (-m) * pow(base, -m - 1.0) * dxn;
        }
        return th;
    }
    if (dth_dpsi) *dth_dpsi = 0.0;
    return thsat;
}

/* inverse: ψ(θ) = -1/α ((Se^(-1/m) - 1))^(1/n) for θ<θsat, else 0 */
double cg_vg_psi(double th, double thsat, double thres, double alpha, double n)
{
    double m = 1.0 - 1.0 / n;
    if (th < thsat) {
        double Se = (th - thres) / (thsat - thres);
        return -1.0 / alpha * pow(pow(Se, -1.0 / m) - 1.0, 1.0 / n);
    }
    return 0.0;
}

/* SFCC θw(T, sat; θsat[, saltconc]) with ∂θw/∂T (and ∂θw/∂c for DallAmicoSalt).
 * Call sites: Soils/soil_heat.jl:57,106,131-132; Salt/salt_diffusion.jl:19.
 *  DallAmico (2011):    ψ0 = ψ(θtot); T* = Tm + g Tm/Lsl ψ0;  ψ(T) = ψ0 + Lsl/(g T*) (T-T*) [T<=T*]
 *  PainterKarra (2014): T* = Tm + ω g Tm/Lsl ψ0;            ψ(T) = ψ0 + β Lsl/(g T*) (T-T*) [T<=T*]
 *  DallAmicoSalt:       Tm <- Tm + Tm/Lf (-R c Tm) (freezing point depression), then DallAmico.
 * Temperatures are normalised to Kelvin (T + 273.15) as FreezeCurves.normalize_temperature does. */
double cg_sfcc(const cg_freezecurve* fc, double T, double sat, double thsat, double saltconc,
               double* dthw_dT, double* dthw_dc)
{
    double thtot = sat * thsat;
    double thsat_e = thtot > thsat ? thtot : thsat;
    double psi0 = cg_vg_psi(thtot, thsat_e, fc->theta_res, fc->alpha, fc->n);
    double TmK = fc->Tm + 273.15;
    double dTm_dc = 0.0;
    double beta = 1.0, omega = 1.0;
    if (fc->kind == CG_FC_PAINTERKARRA) { beta = fc->beta; omega = fc->omega; }
    if (fc->kind == CG_FC_DALLAMICOSALT) {
        double Lf = fc->Lsl * fc->rho_w;
        dTm_dc = TmK / Lf * (-fc->R * TmK);
        TmK = TmK + TmK / Lf * (-fc->R * saltconc * TmK);
    }
    double gT = omega * fc->g / fc->Lsl * psi0;        /* T* = Tm (1 + ω g ψ0/Lsl) */
    double Tstar = TmK + gT * TmK;
    double dTstar_dc = dTm_dc * (1.0 + gT);
    double TK = T + 273.15;
    double psi, dpsi_dT, dpsi_dc;
    if (Tstar - TK >= 0.0) { /* heaviside(T* - T) = 1 */
        double cfac = beta * fc->Lsl / (fc->g * Tstar);
        psi = psi0 + cfac * (TK - Tstar);
        dpsi_dT = cfac;
        /* psi = psi0 + beta Lsl/g (TK/Tstar - 1)  ->  dpsi/dc = beta Lsl/g (-TK/Tstar^2) dTstar/dc */
        dpsi_dc = beta * fc->Lsl / fc->g * (-TK / (Tstar * Tstar)) * dTstar_dc;
    } else {
        psi = psi0; dpsi_dT = 0.0; dpsi_dc = 0.0;
    }
    double dth_dpsi;
    double thw = cg_vg_theta(psi, thsat_e, fc->theta_res, fc->alpha, fc->n, &dth_dpsi);
    if (dthw_dT) *dthw_dT = dth_dpsi * dpsi_dT;
    if (dthw_dc) *dthw_dc = dth_dpsi * dpsi_dc;
    return thw;
}

/* Soils/soil_heat.jl:7-17 sfccheatcap: NOTE the reference quirk -- mineral()/organic() already
 * carry (1-por), and are multiplied by (1-θsat) again. Reproduced verbatim. */
double cg_sfcc_heatcap(const cg_thermal* tp, double por, double org, double thw, double thwi, double thsat)
{
    double thi = thwi - thw;
    double tha = thsat - thwi;
    double thm = soil_thm(por, org) * (1.0 - thsat);
    double tho = soil_tho(por, org) * (1.0 - thsat);
    return tp->ch_w * thw + tp->ch_i * thi + tp->ch_a * tha + tp->ch_m * thm + tp->ch_o * tho;
}

/* H(T) for the SFCC inverse-enthalpy objective: H = T*C(θw(T)) + L*θw(T) (heat_methods.jl:93) and
 * dH/dT = C + ∂θw∂T (L + T (c_w - c_i)) (heat_methods.jl:105). */
static double sfcc_enthalpy(const cg_freezecurve* fc, const cg_thermal* tp, double por, double sat, double org,
                            double T, double* dHdT, double* thw_out)
{
    double dth;
    double thw = cg_sfcc(fc, T, sat, por, 0.0, &dth, NULL);
    double C = cg_sfcc_heatcap(tp, por, org, thw, sat * por, por);
    if (dHdT) *dHdT = C + dth * (tp->L + T * (tp->ch_w - tp->ch_i));
    if (thw_out) *thw_out = thw;
    return T * C + tp->L * thw;
}

/* Tightly converged safeguarded Newton for H(T) = H (monotone increasing in T).
 * Accuracy-mode counterpart of SFCCNewtonSolver (Soils/soil_heat.jl:131-132 call site); the
 * reference's default tolerances are loose, so only tightly converged results are compared. */
double cg_sfcc_newton(const cg_freezecurve* fc, const cg_thermal* tp, double por, double sat, double org,
                      double H, double T0)
{
    double lo = -273.0, hi = 1.0e4;
    double T = T0;
    if (!(T > lo && T < hi)) T = 0.0;
    for (int it = 0; it < 200; ++it) {
        double dHdT;
        double r = sfcc_enthalpy(fc, tp, por, sat, org, T, &dHdT, NULL) - H;
        if (r > 0.0) hi = T; else lo = T;
        double Tn = T - r / dHdT;
        /* Newton step if strictly inside the bracket, or on its end when converged (strict-only bisects
         * there; inclusive-only lets Newton 2-cycle between the bracket ends) */
        int small = fabs(Tn - T) <= 1e-15 * fmax(1.0, fabs(Tn));
        int safe = (Tn > lo && Tn < hi) || (small && Tn >= lo && Tn <= hi);
        if (!safe) Tn = 0.5 * (lo + hi);
        if (fabs(Tn - T) <= 1e-15 * fmax(1.0, fabs(Tn))) { T = Tn; break; } /* converged, or bracket collapsed */
        T = Tn;
    }
    return T;
}

/* Gridded(Linear) interpolant evaluation (Interpolations.jl semantics as used by
 * FreezeCurves' presolver cache and IO/forcings/providers.jl:22): find i with x[i] <= x < x[i+1],
 * w = (x-x[i])/(x[i+1]-x[i]); y = (1-w) y[i] + w y[i+1]. */
static double lerp_table(const double* xk, const double* yk, int nk, double x)
{
    int lo = 0, hi = nk - 1; /* searchsortedlast clamped to [0, nk-2] */
    while (hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if (xk[mid] <= x) lo = mid; else hi = mid;
    }
    double w = (x - xk[lo]) / (xk[lo + 1] - xk[lo]);
    return (1.0 - w) * yk[lo] + w * yk[lo + 1];
}

double cg_lut_eval(const double* Hk, const double* Tk, int nk, int extrap, double H)
{
    if (extrap == CG_EXTRAP_FLAT) {
        if (H <= Hk[0]) return Tk[0];
        if (H >= Hk[nk - 1]) return Tk[nk - 1];
    }
    return lerp_table(Hk, Tk, nk, H);
}

/* [FreezeCurves 0.9, parity unpinned] SFCCPreSolver(Cache1D; Tmin=-60, errtol=1e-4) initialize!
 * (call site Soils/soil_heat.jl:26-36): march H from H(Tmin) to H(Tmax); at each knot take the
 * first-order predictor T + ΔH dT/dH, start from ΔH = Lθtot/10 and halve until the predictor
 * matches the nonlinear solve within errtol (the halving happens after the test, so the accepted
 * step is half the tested one); store (H,T) knots.  Returns the number of knots (<= cap) or -1. */
int cg_presolver_build(const cg_freezecurve* fc, const cg_thermal* tp, double por, double sat, double org,
                       double Tmin, double errtol, double Tmax, double* Hk, double* Tk, int cap)
{
    double thtot = sat * por;
    double dHdT;
    double H = sfcc_enthalpy(fc, tp, por, sat, org, Tmin, &dHdT, NULL);
    double Hmax = sfcc_enthalpy(fc, tp, por, sat, org, Tmax, NULL, NULL);
    double T = Tmin;
    double dTdH = 1.0 / dHdT;
    int nk = 0;
    Hk[nk] = H; Tk[nk] = T; ++nk;
    while (H < Hmax) {
        double eps = INFINITY;
        double dH = tp->L * (thtot > 1e-8 ? thtot : 1e-8) / 10.0;
        int guard = 0;
        while (fabs(eps) > errtol && guard++ < 200) {
            double Test = T + dH * dTdH;
            double Tsol = cg_sfcc_newton(fc, tp, por, sat, org, H + dH, Test);
            eps = fabs(Tsol - Test);
            dH *= 0.5;
        }
        double Hnew = H + dH;
        double Tnew = cg_sfcc_newton(fc, tp, por, sat, org, Hnew, T + dH * dTdH);
        sfcc_enthalpy(fc, tp, por, sat, org, Tnew, &dHdT, NULL);
        dTdH = 1.0 / dHdT;
        H = Hnew; T = Tnew;
        if (nk >= cap) return -1;
        Hk[nk] = H; Tk[nk] = T; ++nk;
    }
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * Forcing / boundary values
 * ---------------------------------------------------------------------------------------- */

/* ConstantBC (simple_bc.jl:19), PeriodicBC (simple_bc.jl:37): A sin(2π(t/P) + φ) + level,
 * Interpolated1D (IO/forcings/providers.jl:30-33): gridded linear, throws outside the knots. */
double cg_forcing_eval(const cg_forcing* f, double t, int* status)
{
    switch (f->kind) {
    case CG_FORCING_CONSTANT: return f->value;
    case CG_FORCING_PERIODIC: return f->amplitude * sin(2.0 * M_PI * (t / f->period) + f->phase) + f->level;
    default:
        if (t < f->t[0] || t > f->t[f->n - 1]) {
            if (status) *status |= CG_STATUS_FORCING_RANGE;
            return NAN;
        }
        return lerp_table(f->t, f->y, f->n, t);
    }
}

/* Heat/heat_bc.jl:78  nfactor(Tair,nfw,nfs) = (Tair<=0)*nfw + (Tair>0)*nfs */
double cg_nfactor(double Tair, double nf, double nt)
{
    return (Tair <= 0.0 ? 1.0 : 0.0) * nf + (Tair > 0.0 ? 1.0 : 0.0) * nt;
}

/* ------------------------------------------------------------------------------------------
 * Column RHS
 * ---------------------------------------------------------------------------------------- */

#ifdef CG_ORACLE_FAST
/* timing copy: the grid arrays of a work block are computed once per (grid, work block), not once per RHS evaluation.  The cache is
 * keyed by address, so work_alloc() -- whose fresh, zeroed block may land on the address of the one just freed -- resets it. */
static __thread const double* cg_fast_last_edges = NULL;
static __thread const double* cg_fast_last_zc = NULL;
#endif

void cg_work_grid(const cg_column* col, cg_work* w)
{
#ifdef CG_ORACLE_FAST
    if (cg_fast_last_edges == col->edges && cg_fast_last_zc == w->zc) return;
    cg_fast_last_edges = col->edges; cg_fast_last_zc = w->zc;
#endif
    cg_grid(col->edges, col->n + 1, w->zc, w->dk, w->dxc);
}

/* FreeWater: Heat/heat_conduction.jl:18-30.  SFCC: Soils/soil_heat.jl:116-140. */
void cg_freezethaw(const cg_column* col, const double* H, cg_work* w)
{
    const cg_thermal* tp = &col->tp;
    for (int i = 0; i < col->n; ++i) {
        int l = col->cell_layer[i];
        double por = col->por[l], sat = col->sat[l], org = col->org[l];
        const cg_freezecurve* fc = &col->fc[l];
        double thwi = soil_thwi(por, sat); /* Hydrology.watercontent, Soils/para/simple.jl:87 */
        if (col->op == CG_OP_TEMPERATURE) {
            /* Soils/soil_heat.jl:100-113 freezethaw!(sfcc, soil, heat::HeatBalance{<:TemperatureBased}): the state IS T.
             * QUIRK reproduced verbatim: the curve is called as sfcc(T; θsat = porosity), i.e. with the DEFAULT saturation 1.0
             * of the FreezeCurves functors, whatever the soil's saturation; C is the regular heatcapacity() (not sfccheatcap).
             * FreeWater has no temperature-form method in the reference (MethodError): NaN here. */
            double T = H[i];
            if (fc->kind == CG_FC_FREEWATER) { w->T[i] = T; w->thw[i] = w->C[i] = w->dHdT[i] = w->dthwdT[i] = NAN; continue; }
            double dth;
            double thw = cg_sfcc(fc, T, 1.0, por, 0.0, &dth, NULL);
            double C = soil_heatcapacity(tp, thw, thwi, soil_thm(por, org), soil_tho(por, org));
            w->T[i] = T; w->thw[i] = thw; w->C[i] = C; w->dthwdT[i] = dth;
            w->dHdT[i] = C + dth * (tp->L + T * (tp->ch_w - tp->ch_i)); /* heat_methods.jl:105 */
            continue;
        }
        if (fc->kind == CG_FC_FREEWATER) {
            double thw;
            cg_freewater(H[i], thwi, tp->L, &thw);
            w->thw[i] = thw;
            double C = w->C[i] = soil_heatcapacity(tp, thw, thwi, soil_thm(por, org), soil_tho(por, org));
            double T = w->T[i] = cg_enthalpyinv_freewater(H[i], thwi, C, tp->L);
            w->dHdT[i] = (T == 0.0) ? 1e8 : C; /* heat_conduction.jl:27 */
            w->dthwdT[i] = 0.0;
        } else {
            double sat_e = thwi / por; /* soil_heat.jl:126 */
            double T;
            if (fc->solver == CG_SOLVER_PRESOLVER) {
                T = cg_lut_eval(fc->Hk, fc->Tk, fc->nk, fc->extrap, H[i]);
            } else {
                double T0 = (i > 0) ? w->T[i - 1] : H[i] / tp->ch_w; /* soil_heat.jl:127 */
                T = cg_sfcc_newton(fc, tp, por, sat_e, org, H[i], T0);
            }
            double dth;
            double thw = cg_sfcc(fc, T, sat_e, por, 0.0, &dth, NULL);
            double C = cg_sfcc_heatcap(tp, por, org, thw, thwi, por);
            w->T[i] = T; w->thw[i] = thw; w->C[i] = C; w->dthwdT[i] = dth;
            w->dHdT[i] = C + dth * (tp->L + T * (tp->ch_w - tp->ch_i)); /* heat_methods.jl:105 */
        }
    }
}

/* Heat/thermal_properties.jl:53-66 (+ layer interfaces Heat/heat_conduction.jl:1-9,128-139, which
 * apply the same harmonic mean to the shared edge). */
void cg_thermalconductivity(const cg_column* col, cg_work* w)
{
    const cg_thermal* tp = &col->tp;
    int n = col->n;
#ifdef CG_ORACLE_FAST
    const double sw = sqrt(tp->kh_w), si = sqrt(tp->kh_i), sa = sqrt(tp->kh_a), sm = sqrt(tp->kh_m), so = sqrt(tp->kh_o);
#endif
    for (int i = 0; i < n; ++i) {
        int l = col->cell_layer[i];
        double por = col->por[l], sat = col->sat[l], org = col->org[l];
#ifdef CG_ORACLE_FAST
        {
            double thwi = soil_thwi(por, sat), thm = soil_thm(por, org), tho = soil_tho(por, org);
            double s = sw * w->thw[i] + si * (thwi - w->thw[i]) + sa * (1.0 - thwi - thm - tho) + sm * thm + so * tho;
            w->kc[i] = s * s;
        }
#else
        w->kc[i] = soil_thermalconductivity(tp, w->thw[i], soil_thwi(por, sat), soil_thm(por, org), soil_tho(por, org));
#endif
        if (i > 0) w->k[i] = cg_harmonicmean(w->kc[i - 1], w->kc[i], w->dk[i - 1], w->dk[i]);
    }
    w->k[0] = w->kc[0];
    w->k[n] = w->kc[n - 1];
}

static double top_value(const cg_column* col, double t, int* status)
{
    double v = cg_forcing_eval(&col->top, t, status);
    if (col->top_kind == CG_BC_DIRICHLET && col->use_nfactor)
        v = cg_nfactor(v, col->nf, col->nt) * v; /* heat_bc.jl:80-86 */
    return v;
}

/* Heat/heat_implicit.jl:7-9 */
static double apbc(int kind, double k, double dk) { return kind == CG_BC_DIRICHLET ? k / (dk * dk / 2.0) : 0.0; }

/* One evaluation of tile(du,u,p,t): Tiles/tile.jl:136-158.
 * Explicit enthalpy form: resetfluxes! -> computediagnostic! -> interact! -> computeprognostic!.
 * Implicit form (Heat/heat_implicit.jl): fills an, as, ap, bp instead of dH. */
int cg_rhs(const cg_column* col, const double* H, double t, cg_work* w)
{
    int n = col->n, status = 0;
    cg_work_grid(col, w);
    /* resetfluxes!: heat_conduction.jl:193-198 / heat_implicit.jl:108-117 */
    for (int i = 0; i < n; ++i) { w->dH[i] = 0.0; w->an[i] = w->as[i] = w->ap[i] = w->bp[i] = 0.0; }
    for (int i = 0; i <= n; ++i) w->jH[i] = 0.0;
    double top = top_value(col, t, &status);
    double bot = cg_forcing_eval(&col->bot, t, &status);
    /* computediagnostic!: heat_conduction.jl:120-126 */
    cg_freezethaw(col, H, w);
    cg_thermalconductivity(col, w);
    if (col->op == CG_OP_ENTHALPY || col->op == CG_OP_TEMPERATURE) {
        /* interact! Top: heat_bc.jl:9-17,28-31 ; Neumann: methods.jl:156 */
        if (col->top_kind == CG_BC_DIRICHLET) w->jH[0] += cg_flux(top, w->T[0], w->dk[0] / 2.0, w->k[0]);
        else w->jH[0] += top;
        /* interact! Bottom: heat_bc.jl:18-27,32-35 */
        if (col->bot_kind == CG_BC_DIRICHLET) w->jH[n] += cg_flux(w->T[n - 1], bot, w->dk[n - 1] / 2.0, w->k[n]);
        else w->jH[n] += bot;
        /* computeprognostic!: heat_conduction.jl:141-146 -> heatflux! :65-72 */
        cg_flux_vec(w->jH, w->T, w->dxc, w->k, n);
        cg_divergence_vec(w->dH, w->jH, w->dk, n);
        /* temperatureflux!: heat_conduction.jl:79-84 */
        if (col->op == CG_OP_TEMPERATURE) for (int i = 0; i < n; ++i) w->dT[i] = w->dH[i] / w->dHdT[i];
    } else {
        /* prefactors!: heat_implicit.jl:10-24 (layer interfaces :87-103 give the same terms) */
        for (int i = 0; i < n; ++i) {
            if (i > 0) w->an[i] = w->k[i] / w->dxc[i - 1] / w->dk[i];
            if (i < n - 1) w->as[i] = w->k[i + 1] / w->dxc[i] / w->dk[i];
            w->ap[i] = w->an[i] + w->as[i];
        }
        /* Top interact!: heat_implicit.jl:59-64,72-79 */
        {
            double dk = w->dk[0], k = w->k[0];
            double jtop = (col->top_kind == CG_BC_DIRICHLET) ? 2.0 * top * k / dk : top;
            w->bp[0] += jtop / dk;
            w->ap[0] += apbc(col->top_kind, k, dk);
        }
        /* Bottom interact!: heat_implicit.jl:65-70,80-86 -- QUIRKS reproduced verbatim: k = ssub.k[1]
         * (first edge of the bottom LAYER) inside apbc, and the Neumann value enters with the sign of
         * boundaryvalue (= -Qgeo for GeothermalHeatFlux). */
        {
            double dk = w->dk[n - 1];
            double jbot = (col->bot_kind == CG_BC_DIRICHLET) ? 2.0 * bot * w->k[n] / dk : bot;
            int first = n - 1;
            while (first > 0 && col->cell_layer[first - 1] == col->cell_layer[n - 1]) --first;
            double k1 = w->k[first];
            w->bp[n - 1] += jbot / dk;
            w->ap[n - 1] += apbc(col->bot_kind, k1, dk);
        }
    }
    return status;
}

/* steplimiters.jl:15-18 MaxDelta(du,u,t) */
static double maxdelta_limit(double dmax, double du)
{
    double dt = fabs(dmax / du);
    return isfinite(dt) ? dt : INFINITY;
}

/* Heat/heat_conduction.jl:162-191 (+ Tiles/stratigraphy.jl:230-236 min over layers). */
double cg_timestep(const cg_column* col, const double* H, const cg_work* w)
{
    (void)H;
    double dtmax = INFINITY;
    if (col->limiter == CG_LIMITER_MAXDELTA) {
        for (int i = 0; i < col->n; ++i) dtmax = fmin(dtmax, maxdelta_limit(col->maxdelta, w->dH[i]));
        dtmax = (isfinite(dtmax) && dtmax > 0.0) ? dtmax : INFINITY;
    } else {
        for (int i = 0; i < col->n; ++i) {
            double v = w->dHdT[i] / w->kc[i];
            double dx = w->dk[i];
            double dt = col->courant * v * (dx * dx);
            /* derivative(::TemperatureBased) = dT, derivative(::EnthalpyBased) = dH: heat_conduction.jl:163-164 */
            dt = fmin(dt, maxdelta_limit(col->maxdelta, col->op == CG_OP_TEMPERATURE ? w->dT[i] : w->dH[i]));
            dtmax = fmin(dtmax, dt);
        }
        dtmax = isfinite(dtmax) ? dtmax : INFINITY;
    }
    return dtmax;
}

/* ------------------------------------------------------------------------------------------
 * CGEuler: Solvers/basic_solvers.jl:54-79 + Solvers/integrator.jl:84-90,111-132
 * Advance one column from (*t,*dt) until t >= t_target; dt is clipped so the target is hit
 * exactly (saveat!), but -- as in the reference -- only after a step, never before the first.
 * ---------------------------------------------------------------------------------------- */
/* Trace variant (test infrastructure for the restart-parity tests): dt_trace[k] = the dt step k was taken with
 * (k < cap; may be NULL), *nlimited = number of steps whose NEXT dt was decided by the limiter itself, i.e.
 * dtmin < lim < dtmax and not clipped by the save time. */
int cg_cgeuler_advance_trace(const cg_column* col, double* H, double* t, double* dt, double t_target,
                             double dtmin, double dtmax, int64_t* nsteps, cg_work* w,
                             double* dt_trace, int64_t cap, int64_t* nlimited)
{
    int status = 0;
    int n = col->n;
    int64_t k = 0;
    while (*t < t_target) {
        status |= cg_rhs(col, H, *t, w);                  /* tile(du,u,p,t0) */
        const double* du = (col->op == CG_OP_TEMPERATURE) ? w->dT : w->dH;
        for (int i = 0; i < n; ++i) H[i] += (*dt) * du[i];    /* u += dt*du */
        if (dt_trace && k < cap) dt_trace[k] = *dt;
        ++k;
        *t = *t + *dt;
        if (nsteps) ++*nsteps;
        double lim = cg_timestep(col, H, w);              /* timestep from the old du */
        if (!(lim > 0.0)) status |= CG_STATUS_BAD_DT;     /* tile.jl:174 */
        *dt = fmax(fmin(lim, dtmax), dtmin);              /* basic_solvers.jl:77 */
        int limited = (lim > dtmin && lim < dtmax);
        if (*t < t_target) {
            if (t_target - *t < *dt) limited = 0;
            *dt = fmin(*dt, t_target - *t);               /* saveat!: integrator.jl:126-131 */
        }
        *dt = fmin(dtmax, *dt);                           /* integrator.jl:89 */
        if (nlimited && limited) ++*nlimited;
        if (status & (CG_STATUS_FORCING_RANGE)) break;
    }
    return status;
}

int cg_cgeuler_advance(const cg_column* col, double* H, double* t, double* dt, double t_target,
                       double dtmin, double dtmax, int64_t* nsteps, cg_work* w)
{
    return cg_cgeuler_advance_trace(col, H, t, dt, t_target, dtmin, dtmax, nsteps, w, NULL, 0, NULL);
}

/* ------------------------------------------------------------------------------------------
 * LiteImplicitEuler: Solvers/LiteImplicit/cglite_step.jl
 * ---------------------------------------------------------------------------------------- */
int cg_lite_step(const cg_column* col, double* H, double* t, double dt, int miniters, int maxiters,
                 double tol, int64_t* iters, cg_work* w)
{
    int n = col->n, status = 0;
    double tn = *t + dt; /* cglite_step.jl:9 forcing at the NEW time */
    double* H0 = (double*)malloc(sizeof(double) * n * 7);
    double *Tnew = H0 + n, *A = Tnew + n, *B = A + n, *Cc = B + n, *D = Cc + n, *Sp = D + n;
    memcpy(H0, H, sizeof(double) * n); /* cache.uprev */
    /* explicit_step!: dH == 0 for the implicit operator (heat_implicit.jl:106) -> no-op */
    int iter = 1;
    double emax = INFINITY;
    while ((emax > tol && iter <= maxiters) || iter < miniters) { /* cglite_step.jl:43 */
        status |= cg_rhs(col, H, tn, w);
        /* cglite_linsolve!: cglite_step.jl:81-99 */
        for (int i = 0; i < n; ++i) {
            Sp[i] = -w->dHdT[i] / dt;
            double Sc = (H0[i] - H[i]) / dt - Sp[i] * w->T[i];
            B[i] = w->ap[i] - Sp[i];
            D[i] = Sc + w->bp[i];
        }
        for (int i = 0; i < n - 1; ++i) { A[i] = -w->an[i + 1]; Cc[i] = -w->as[i]; }
        cg_tdma_solve(Tnew, A, B, Cc, D, n);
        emax = -INFINITY;
        for (int i = 0; i < n; ++i) { /* cglite_step.jl:58-72 */
            double e = Tnew[i] - w->T[i];
            H[i] += w->dHdT[i] * e;
            emax = fmax(emax, fabs(e));
            if (!isfinite(e)) status |= CG_STATUS_NAN;
        }
        if (iters) ++*iters;
        ++iter;
        if (status & CG_STATUS_NAN) break;
    }
    if (iter > maxiters && emax > tol) status |= CG_STATUS_MAXITERS; /* cglite_step.jl:74-77 */
    *t = tn;
    free(H0);
    return status;
}

int cg_lite_advance(const cg_column* col, double* H, double* t, double* dt, double t_target,
                    int miniters, int maxiters, double tol, int64_t* nsteps, int64_t* iters, cg_work* w)
{
    int status = 0;
    double dtmax = *dt; /* cglite_types.jl:29 dtmax=dt */
    while (*t < t_target) {
        status |= cg_lite_step(col, H, t, *dt, miniters, maxiters, tol, iters, w);
        if (nsteps) ++*nsteps;
        if (*t < t_target) *dt = fmin(*dt, t_target - *t); /* saveat!: integrator.jl:126-131 */
        *dt = fmin(dtmax, *dt);
        if (status & (CG_STATUS_NAN | CG_STATUS_FORCING_RANGE)) break;
    }
    return status;
}

/* ------------------------------------------------------------------------------------------
 * Initial conditions
 * ---------------------------------------------------------------------------------------- */

/* initializers.jl:42-67 InterpInitializer: gridded-linear in depth, Flat extrapolation. */
void cg_interp_profile(const double* zk, const double* vk, int nk, const double* z, int n, double* out)
{
    for (int i = 0; i < n; ++i) {
        if (nk == 1) { out[i] = vk[0]; continue; }
        double zi = z[i];
        if (zi <= zk[0]) out[i] = vk[0];
        else if (zi >= zk[nk - 1]) out[i] = vk[nk - 1];
        else out[i] = lerp_table(zk, vk, nk, zi);
    }
}

/* Soils/soil_heat.jl:38-63 initialize_soil_heat!: θw, C from T, H = T*C + L*θw. */
void cg_init_soil_heat(const cg_column* col, const double* T, double* H, cg_work* w)
{
    const cg_thermal* tp = &col->tp;
    for (int i = 0; i < col->n; ++i) {
        int l = col->cell_layer[i];
        double por = col->por[l], sat = col->sat[l], org = col->org[l];
        double thwi = soil_thwi(por, sat);
        const cg_freezecurve* fc = &col->fc[l];
        double thw;
        if (fc->kind == CG_FC_FREEWATER) thw = (T[i] > 0.0) ? thwi : 0.0;
        else thw = cg_sfcc(fc, T[i], sat, por, 0.0, NULL, NULL);
        double C = soil_heatcapacity(tp, thw, thwi, soil_thm(por, org), soil_tho(por, org));
        w->T[i] = T[i]; w->thw[i] = thw; w->C[i] = C;
        H[i] = T[i] * C + tp->L * thw; /* heat_methods.jl:93 enthalpy */
    }
}

/* Heat/heat_init.jl:58-85 ThermalSteadyStateInit, applied layer by layer (top -> bottom) with the
 * upper layer's last temperature as T0 of the next (heat_init.jl:64-67); z is the ABSOLUTE depth of
 * the cell midpoints (cells(state.grid)) in every layer, as in the reference. */
void cg_steadystate_init(const cg_column* col, double T0, double Qgeo, int maxiters, double thresh,
                         double* H, cg_work* w)
{
    int n = col->n;
    cg_work_grid(col, w);
    double* Tn = (double*)malloc(sizeof(double) * n);
    for (int i = 0; i < n; ++i) { w->T[i] = 0.0; w->kc[i] = 0.0; H[i] = 0.0; }
    int start = 0;
    double Tup = T0;
    while (start < n) {
        int end = start;
        while (end < n && col->cell_layer[end] == col->cell_layer[start]) ++end;
        /* restrict the column to this layer */
        cg_column sub = *col;
        cg_work sw = *w;
        sub.n = end - start;
        sub.edges = col->edges + start;
        sub.cell_layer = col->cell_layer + start;
        sw.T = w->T + start; sw.thw = w->thw + start; sw.C = w->C + start; sw.dHdT = w->dHdT + start;
        sw.dthwdT = w->dthwdT + start; sw.kc = w->kc + start; sw.k = w->k + start; sw.jH = w->jH + start;
        sw.dH = w->dH + start; sw.zc = w->zc + start; sw.dk = w->dk + start; sw.dxc = w->dxc + start;
        double* Hs = H + start;
        int it = 1;
        double dT = INFINITY;
        while (it < maxiters && dT > thresh) {
            dT = 0.0;
            for (int i = 0; i < sub.n; ++i) {
                Tn[i] = Tup + sw.zc[i] * Qgeo / sw.kc[i];
                double d = fabs(Tn[i] - sw.T[i]);
                /* Julia maximum() propagates NaN */
                if (isnan(d) || isnan(dT)) dT = NAN; else dT = fmax(dT, d);
            }
            cg_init_soil_heat(&sub, Tn, Hs, &sw);
            cg_freezethaw(&sub, Hs, &sw);
            cg_thermalconductivity(&sub, &sw);
            ++it;
        }
        for (int i = 0; i < sub.n; ++i) Tn[i] = Tup + sw.zc[i] * Qgeo / sw.kc[i];
        cg_init_soil_heat(&sub, Tn, Hs, &sw);
        cg_freezethaw(&sub, Hs, &sw);
        cg_thermalconductivity(&sub, &sw);
        Tup = sw.T[sub.n - 1];
        start = end;
    }
    /* stratigraphy.jl:131-141: the per-layer initialcondition! runs once more on the final T */
    memcpy(Tn, w->T, sizeof(double) * n);
    cg_init_soil_heat(col, Tn, H, w);
    free(Tn);
}

/* ------------------------------------------------------------------------------------------
 * Coupled heat + salt: src/Physics/Salt  (single SalineGround layer, HeatBalance(:T))
 * ---------------------------------------------------------------------------------------- */

/* Salt/salt_init.jl:9-17 compaction! */
void cg_compaction(const double* edges, int n, double por0, double por_min, double* por)
{
    double zmin = (edges[0] + edges[1]) / 2.0;
    double bd0 = 1.0 / ((por0 + 0.6845) / 1.8);
    for (int i = 0; i < n; ++i) {
        double z = (edges[i] + edges[i + 1]) / 2.0 - zmin;
        double bd = bd0 + 0.0037 * pow(z, 0.766);
        por[i] = fmax(1.80 * pow(bd, -1.0) - 0.6845, por_min);
    }
}

/* Salt/salt_diffusion.jl:3-30 (freezethaw!), :51-78 (computediagnostic!), salt_bc.jl:18-31 and
 * heat_bc.jl:9-35 (interact!), salt_diffusion.jl:124-165 (computeprognostic!). */
void cg_salt_rhs(const cg_salt_column* col, const double* T, const double* c, cg_salt_work* w) { cg_salt_rhs_t(col, T, c, 0.0, w); }

int cg_salt_rhs_t(const cg_salt_column* col, const double* T, const double* c, double t, cg_salt_work* w)
{
    int n = col->n, status = 0;
    const double T_ub = col->use_top_forcing ? cg_forcing_eval(&col->top, t, &status) : col->T_ub;
    const cg_thermal* tp = &col->tp;
    cg_grid(col->edges, n + 1, w->zc, w->dk, w->dxc);
    /* SimpleSoil defaults of the parameterisation (Soils/para/simple.jl:9): mineral()/organic() use
     * para.por = 0.5, NOT the compacted state.por. */
    const double para_por = 0.5;
    double thm = soil_thm(para_por, col->org), tho = soil_tho(para_por, col->org);
    for (int i = 0; i < n; ++i) {
        double por = col->por[i];
        double dT, dc;
        double thw = cg_sfcc(&col->fc, T[i], col->sat, por, c[i], &dT, &dc);
        double thwi = por * col->sat;
        w->thw[i] = thw; w->dthwdT[i] = dT; w->dthwdc[i] = dc;
        double C = w->C[i] = soil_heatcapacity(tp, thw, thwi, thm, tho);
        w->dHdT[i] = C + tp->L * dT;
        w->H[i] = T[i] * C + tp->L * thw;
        w->ds_mid[i] = col->ds0 * thw / col->tau;
        w->kc[i] = soil_thermalconductivity(tp, thw, thwi, thm, tho);
    }
    w->k[0] = w->kc[0]; w->k[n] = w->kc[n - 1];
    w->ds[0] = w->ds_mid[0]; w->ds[n] = w->ds_mid[n - 1];
    for (int i = 1; i < n; ++i) {
        w->k[i] = cg_harmonicmean(w->kc[i - 1], w->kc[i], w->dk[i - 1], w->dk[i]);
        w->ds[i] = cg_harmonicmean(w->ds_mid[i - 1], w->ds_mid[i], w->dk[i - 1], w->dk[i]);
    }
    for (int i = 0; i < n; ++i) { w->dH[i] = 0.0; w->dc_F[i] = 0.0; }
    for (int i = 0; i <= n; ++i) { w->jH[i] = 0.0; w->jc[i] = 0.0; }
    /* boundaries */
    w->jH[0] += cg_flux(T_ub, T[0], w->dk[0] / 2.0, w->k[0]);
    w->jc[0] = (col->surface_state == 0) ? -w->ds[0] * (c[0] - col->benthic_salt) / (w->dk[0] / 2.0) : 0.0;
    w->jH[n] += col->bot_value;
    cg_nonlineardiffusion(w->dH, w->jH, T, w->dxc, w->k, w->dk, n);
    cg_nonlineardiffusion(w->dc_F, w->jc, c, w->dxc, w->ds, w->dk, n);
    for (int i = 0; i < n; ++i) {
        double A = w->dHdT[i];
        double B = tp->L * w->dthwdc[i];
        double D = w->dH[i];
        double E = w->thw[i] + c[i] * w->dthwdc[i];
        double F = c[i] * w->dthwdT[i];
        double G = w->dc_F[i];
        w->dT[i] = (-B * G + D * E) / (A * E - B * F);
        w->dc[i] = (-F * D + A * G) / (A * E - B * F);
    }
    return status;
}

/* salt_diffusion.jl:107-122 (salt CFL + MaxDelta) and heat_conduction.jl:162-182 (heat CFL). */
double cg_salt_timestep(const cg_salt_column* col, const cg_salt_work* w)
{
    double dts = INFINITY, dth = INFINITY;
    for (int i = 0; i < col->n; ++i) {
        double dx = w->dk[i];
        double v = 1.0 / w->ds[i];
        double dtc = col->courant * v * (dx * dx);
        dts = fmin(fmin(dts, dtc), col->dc_max / fabs(w->dc[i]));
        double vh = w->dHdT[i] / w->kc[i];
        double dt = col->heat_courant * vh * (dx * dx);
        dt = fmin(dt, maxdelta_limit(INFINITY, w->dT[i]));
        dth = fmin(dth, dt);
    }
    dth = isfinite(dth) ? dth : INFINITY;
    return fmin(dts, dth);
}

int cg_salt_advance_trace(const cg_salt_column* col, double* T, double* c, double* t, double* dt, double t_target,
                          double dtmin, double dtmax, int64_t* nsteps, cg_salt_work* w, double* dt_trace, int64_t cap)
{
    int status = 0, n = col->n;
    int64_t k = 0;
    while (*t < t_target) {
        status |= cg_salt_rhs_t(col, T, c, *t, w);
        for (int i = 0; i < n; ++i) { T[i] += (*dt) * w->dT[i]; c[i] += (*dt) * w->dc[i]; }
        if (dt_trace && k < cap) dt_trace[k] = *dt;
        ++k;
        *t = *t + *dt;
        if (nsteps) ++*nsteps;
        double lim = cg_salt_timestep(col, w);
        if (!(lim > 0.0)) status |= CG_STATUS_BAD_DT;
        *dt = fmax(fmin(lim, dtmax), dtmin);
        if (*t < t_target) *dt = fmin(*dt, t_target - *t);
        *dt = fmin(dtmax, *dt);
    }
    return status;
}

int cg_salt_advance(const cg_salt_column* col, double* T, double* c, double* t, double* dt, double t_target,
                    double dtmin, double dtmax, int64_t* nsteps, cg_salt_work* w)
{
    return cg_salt_advance_trace(col, T, c, t, dt, t_target, dtmin, dtmax, nsteps, w, NULL, 0);
}

/* ------------------------------------------------------------------------------------------
 * Batch drivers (OpenMP over columns) -- the "CPU path timed beside it" (SURVEY.md 8d).
 * ---------------------------------------------------------------------------------------- */
static void work_alloc(cg_work* w, int n, double** block)
{
    double* p = (double*)calloc((size_t)(17 * (n + 1)), sizeof(double));
    *block = p;
#ifdef CG_ORACLE_FAST
    cg_fast_last_edges = NULL; cg_fast_last_zc = NULL;
#endif
    w->T = p; p += n; w->thw = p; p += n; w->C = p; p += n; w->dHdT = p; p += n; w->dthwdT = p; p += n;
    w->kc = p; p += n; w->k = p; p += n + 1; w->jH = p; p += n + 1; w->dH = p; p += n;
    w->an = p; p += n; w->as = p; p += n; w->ap = p; p += n; w->bp = p; p += n;
    w->zc = p; p += n; w->dk = p; p += n; w->dxc = p; p += n; w->dT = p;
}

int cg_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int cg_cgeuler_advance_batch(const cg_column* cols, int64_t M, double* H, double* t, double* dt,
                             double t_target, double dtmin, double dtmax, int64_t* nsteps, int nthreads)
{
    int status = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cg_work w; double* block = NULL; int nalloc = 0;
#pragma omp for schedule(dynamic, 1) reduction(| : status)
        for (int64_t c = 0; c < M; ++c) {
            int n = cols[c].n;
            if (n > nalloc) { free(block); work_alloc(&w, n, &block); nalloc = n; }
            status |= cg_cgeuler_advance(&cols[c], H + c * n, &t[c], &dt[c], t_target, dtmin, dtmax,
                                         nsteps ? &nsteps[c] : NULL, &w);
        }
        free(block);
    }
    return status;
}

int cg_lite_advance_batch(const cg_column* cols, int64_t M, double* H, double* t, double* dt, double t_target,
                          int miniters, int maxiters, double tol, int64_t* nsteps, int64_t* iters, int nthreads)
{
    int status = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        cg_work w; double* block = NULL; int nalloc = 0;
#pragma omp for schedule(dynamic, 1) reduction(| : status)
        for (int64_t c = 0; c < M; ++c) {
            int n = cols[c].n;
            if (n > nalloc) { free(block); work_alloc(&w, n, &block); nalloc = n; }
            status |= cg_lite_advance(&cols[c], H + c * n, &t[c], &dt[c], t_target, miniters, maxiters, tol,
                                      nsteps ? &nsteps[c] : NULL, iters ? &iters[c] : NULL, &w);
        }
        free(block);
    }
    return status;
}
