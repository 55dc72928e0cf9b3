# CryoGridB200.jl -- Julia glue a CryoGrid.jl maintainer would add to bind libcgb200.so (include/cgb200.h).
#
# UNVERIFIED: Julia is not installed in the build image, so this file has never been executed.  It is written
# against the same header the Python ctypes harness (cryogrid.jl_b200/_lib.py) is tested with.
#
# Seams used (CryoGrid.jl v0.23.4):
#   * algorithm dispatch  DiffEqBase.__solve(prob, alg::CryoGridODEAlgorithm)   src/Solvers/integrator.jl:92-107
#   * ensemble algorithm  solve(::EnsembleProblem, alg, ::EnsembleAlgorithm)    src/Solvers/DiffEq/ensemble.jl:33-45
module CryoGridB200

using CryoGrid
using CryoGrid: Tile, CryoGridProblem, CryoGridODEAlgorithm
import DiffEqBase, SciMLBase

const libcgb = get(ENV, "CGB200_LIB", "libcgb200.so")

# mirrors cgb_config (include/cgb200.h) field for field
struct CgbConfig
    abi_version::Int32; device::Int32; n_cells::Int32; n_layers::Int32
    n_columns::Int64
    physics::Int32; stepper::Int32; limiter::Int32; top_kind::Int32; bot_kind::Int32; use_nfactor::Int32
    maxdelta::Float64; courant::Float64; dt0::Float64; dtmin::Float64; dtmax::Float64
    lite_miniters::Int32; lite_maxiters::Int32; lite_tol::Float64
end

check(h, rc) = rc == 0 || error("libcgb200: " * unsafe_string(ccall((:cgb_last_error, libcgb), Cstring, (Ptr{Cvoid},), h)))

"Explicit Euler on the GPU for M columns at once (replaces CGEuler, src/Solvers/basic_solvers.jl:9-79)."
struct CGEulerB200 <: CryoGridODEAlgorithm end
"CryoGridLite implicit Euler on the GPU (replaces LiteImplicitEuler, src/Solvers/LiteImplicit/cglite_types.jl:1-6)."
Base.@kwdef struct LiteImplicitEulerB200 <: CryoGridODEAlgorithm
    miniters::Int = 2; maxiters::Int = 1000; tolerance::Float64 = 1e-3
end
"EnsembleAlgorithm: all trajectories of an EnsembleProblem become the column axis of one GPU batch."
struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    device::Int
end

"""
    solve_batch(probs::Vector{<:CryoGridProblem}, alg; dt, dtmax=3600.0, dtmin=1.0, saveat, savevars="T")

`probs` are the `remake`d members of an ensemble (same Tile structure and grid, different parameters/u0).
Flattens them to the structure-of-arrays batch, runs the whole integration in one `cgb_solve_saveat` call and
returns `out[col, cell, var, save]` (column-fastest, as the library writes it).
"""
function solve_batch(probs, alg; dt=60.0, dtmax=3600.0, dtmin=1.0, saveat, savevars="T", device=0)
    tile1 = Tile(first(probs).f)
    edges = collect(Float64, CryoGrid.Numerics.edges(tile1.grid))
    N, M = length(edges) - 1, length(probs)
    layers = flatten_layers(tile1)                      # per-cell layer index (0-based), see below
    nl = maximum(layers) + 1
    lite = alg isa LiteImplicitEulerB200
    cfg = CgbConfig(1, device, N, nl, M, 0, lite ? 1 : 0, 0, 0, 1, 1, 5e4, 0.5, dt, dtmin, lite ? dt : dtmax,
                    lite ? alg.miniters : 2, lite ? alg.maxiters : 1000, lite ? alg.tolerance : 1e-3)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cgb_create, libcgb), Cint, (Ref{CgbConfig}, Ref{Ptr{Cvoid}}), cfg, href)
    rc == 0 || error("cgb_create: " * unsafe_string(ccall((:cgb_last_error, libcgb), Cstring, (Ptr{Cvoid},), C_NULL)))
    h = href[]
    try
        check(h, ccall((:cgb_set_grid, libcgb), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, edges))
        check(h, ccall((:cgb_set_layers, libcgb), Cint, (Ptr{Cvoid}, Ptr{Int32}), h, Int32.(layers)))
        # per-(layer, column) soil parameters, p[layer*M + col] (PER_COLUMN_LAYER = 3)
        for name in ("por", "sat", "org")
            vals = [getfield(soilpara(Tile(p.f), l), Symbol(name)) for p in probs, l in 1:nl]   # M x nl, column-fastest
            check(h, ccall((:cgb_set_param, libcgb), Cint, (Ptr{Cvoid}, Cstring, Ptr{Float64}, Int64, Int32),
                           h, name, vec(Float64.(vals)), M * nl, 3))
        end
        # forcing table shared by all members (Interpolated1D, src/IO/forcings/providers.jl:14-33)
        ts, Tair = forcing_knots(tile1, :Tair)
        check(h, ccall((:cgb_set_forcing_table, libcgb), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int64),
                       h, 0, ts, Tair, length(ts)))
        # SFCC: hand over the Julia-built presolver knots so LUT parity does not depend on our table builder
        for l in 0:nl-1
            fc = freezecurve_of(tile1, l + 1)
            fc isa FreeWater && continue
            check(h, ccall((:cgb_set_freezecurve, libcgb), Cint, (Ptr{Cvoid}, Int32, Int32, Int32), h, l, fckind(fc), 0))
            Hk, Tk = presolver_knots(tile1, l + 1)       # knots of solver.cache.f_H
            check(h, ccall((:cgb_set_sfcc_table, libcgb), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Int32),
                           h, l, Hk, Tk, length(Hk), 1))
        end
        u0 = reduce(hcat, [collect(Float64, p.u0.H) for p in probs])'   # M x N -> memory order (cell, col) column-fastest
        check(h, ccall((:cgb_set_state, libcgb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), h, collect(u0), first(probs).tspan[1]))
        nsave = length(saveat)
        nvars = count(==(','), savevars) + 1
        out = Array{Float64}(undef, M, N, nvars, nsave)
        check(h, ccall((:cgb_solve_saveat, libcgb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Cstring, Ptr{Float64}),
                       h, collect(Float64, saveat), nsave, savevars, out))
        status = Vector{Int32}(undef, M); steps = Vector{Int64}(undef, M)
        check(h, ccall((:cgb_get_status, libcgb), Cint, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int64}), h, status, steps))
        return (; out, status, steps)
    finally
        ccall((:cgb_destroy, libcgb), Cvoid, (Ptr{Cvoid},), h)
    end
end

# SciML ensemble entry point: collect the remade members, run them as one batch, wrap the outputs.
function SciMLBase.__solve(ensprob::SciMLBase.EnsembleProblem, alg::CryoGridODEAlgorithm, ens::EnsembleB200;
                           trajectories, saveat, kwargs...)
    probs = [ensprob.prob_func(deepcopy(ensprob.prob), i, 1) for i in 1:trajectories]
    res = solve_batch(probs, alg; saveat, device=ens.device, kwargs...)
    return res   # a maintainer would wrap res.out[i, :, :, :] into CryoGridOutput per member here
end

# Optional entry points (same handle `h`, before cgb_destroy):
#   on-device initial condition instead of cgb_set_state (initializers.jl:42-67 / Heat/heat_init.jl:58-85):
#     ccall((:cgb_init_interp, libcgb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int32, Float64), h, zk, Tk, nk, 0, t0)
#     ccall((:cgb_init_steadystate, libcgb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Int32, Float64, Float64),
#           h, T0, per, Qgeo, maxiters, thresh, t0)
#   device-side output path: selected cells of every saved variable + Diagnostics.thawdepth per column and save time
#     ccall((:cgb_solve_saveat_select, libcgb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Cstring, Ptr{Int32}, Int32, Ptr{Float64}, Ptr{Float64}),
#           h, tsave, nsave, "T,thw", Int32.(cells .- 1), length(cells), out_sel, out_thaw)
#   ensemble moments without moving the fields (per-cell sum x and sum x^2 over the columns of this handle):
#     ccall((:cgb_ensemble_partial, libcgb), Cint, (Ptr{Cvoid}, Cstring, Float64, Ptr{Float64}, Ptr{Float64}), h, "T", t, s1, s2)
# Saved "T"/"thw" are the reference's saved values (the diagnostic cache of the last RHS evaluation of the interval, one step
# behind the saved "H"); "T_now"/"thw_now" are recomputed from the saved state (INTEGRATION.md).

# --- helpers a maintainer fills in from CryoGrid internals (Stratigraphy / Tile accessors) ---
function flatten_layers end      # per-cell 0-based layer index from Tiles.layeridx (src/Tiles/state.jl:91-131)
function soilpara end            # SimpleSoil of layer l (src/Physics/Soils/para/simple.jl:8-15)
function forcing_knots end       # (ts, values) of an Interpolated1D input
function freezecurve_of end
function fckind end              # DallAmico -> 1, PainterKarra -> 2, DallAmicoSalt -> 3
function presolver_knots end     # knots of the SFCCPreSolver interpolant (FreezeCurves.jl)

end # module
