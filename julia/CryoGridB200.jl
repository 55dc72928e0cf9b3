# CryoGridB200.jl -- Julia glue a CryoGrid.jl maintainer would add to bind libcgb200.so (include/cgb200.h).
#
# UNVERIFIED: Julia is not installed in the build image, so this file has never been executed.  It is written
# against the same header the Python ctypes harness (cryogrid.jl_b200/_lib.py) is tested with, sends everything that harness
# sends (Engine._push_description), and refuses -- with an error, not with defaults -- what it does not wire.
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

Everything `Engine._push_description` of the tested Python harness sends is sent here too: per-(layer, column) soil parameters
and van Genuchten / freeze-curve parameters, n-factors, the bottom boundary value, the forcing table, the freeze curve and
solver of every layer and -- for the presolver -- every member's own Julia-built knots with a (layer, column) -> table map.
Configurations the engine does not implement raise an error instead of running with defaults.
"""
function solve_batch(probs, alg; dt=60.0, dtmax=3600.0, dtmin=1.0, saveat, savevars="T", device=0)
    t0 = first(probs).tspan[1]
    tiles = [resolved_tile(p) for p in probs]           # parameters of every member replaced by their values
    tile1 = tiles[1]
    raw1 = Tile(first(probs).f)                         # Inputs still symbolic: needed to find the forcing
    edges = collect(Float64, ustrip.(CryoGrid.Numerics.edges(tile1.grid)))
    N, M = length(edges) - 1, length(probs)
    layers = flatten_layers(tile1)                      # per-cell 0-based layer index
    nl = Int(maximum(layers)) + 1
    lite = alg isa LiteImplicitEulerB200
    top, bot = top_process(tile1), bottom_process(tile1)
    top isa TemperatureBC || error("CryoGridB200: upper boundary $(typeof(top)) is not supported (TemperatureBC[+NFactor] only)")
    bot isa GeothermalHeatFlux || error("CryoGridB200: lower boundary $(typeof(bot)) is not supported (GeothermalHeatFlux only)")
    use_nf = top.effect isa NFactor
    (top.effect === nothing || use_nf) || error("CryoGridB200: boundary effect $(typeof(top.effect)) is not supported")
    heat1 = sublayer(tile1, 1).heat
    lim = heat1.dtlim
    lim isa CryoGrid.MaxDelta || lite || error("CryoGridB200: step limiter $(typeof(lim)) is not wired in this glue (MaxDelta only)")
    maxdelta = lim isa CryoGrid.MaxDelta ? Float64(ustrip(lim.Δmax)) : 5e4
    #                 abi dev     N  nl  M  physics stepper      limiter top=Dirichlet bot=Neumann
    cfg = CgbConfig(1, device, N, nl, M, 0, lite ? 1 : 0, 0, 0, 1, use_nf ? 1 : 0, maxdelta, 0.5, dt, dtmin, lite ? dt : dtmax,
                    lite ? alg.miniters : 2, lite ? alg.maxiters : 1000, lite ? alg.tolerance : 1e-3)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cgb_create, libcgb), Cint, (Ref{CgbConfig}, Ref{Ptr{Cvoid}}), cfg, href)
    rc == 0 || error("cgb_create: " * unsafe_string(ccall((:cgb_last_error, libcgb), Cstring, (Ptr{Cvoid},), C_NULL)))
    h = href[]
    setparam(name, vals, per) = check(h, ccall((:cgb_set_param, libcgb), Cint, (Ptr{Cvoid}, Cstring, Ptr{Float64}, Int64, Int32),
                                               h, name, collect(Float64, vec(vals)), length(vals), per))
    try
        check(h, ccall((:cgb_set_grid, libcgb), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, edges))
        check(h, ccall((:cgb_set_layers, libcgb), Cint, (Ptr{Cvoid}, Ptr{Int32}), h, Int32.(layers)))
        # per-(layer, column) soil parameters, p[layer*M + col] (PER_COLUMN_LAYER = 3): an M x nl matrix in Julia's column-major order
        for name in (:por, :sat, :org)
            setparam(string(name), [num(getfield(soilpara(tl, l), name)) for tl in tiles, l in 1:nl], 3)
        end
        # thermal properties (Heat/thermal_properties.jl:10-19, Soils/para/simple.jl:32-45): one value for the whole batch
        hp, sp = heat1.prop, soilpara(tile1, 1).heat
        for (name, v) in (("kh_w", hp.kh_w), ("kh_i", hp.kh_i), ("kh_a", hp.kh_a), ("kh_m", sp.kh_m), ("kh_o", sp.kh_o),
                          ("ch_w", hp.ch_w), ("ch_i", hp.ch_i), ("ch_a", hp.ch_a), ("ch_m", sp.ch_m), ("ch_o", sp.ch_o), ("L", hp.L))
            setparam(name, [num(v)], 0)
        end
        # n-factors (heat_bc.jl:60-86) and the geothermal flux: boundaryvalue = -Qgeo (heat_bc.jl:113), per column (PER_COLUMN = 2)
        if use_nf
            setparam("nf", [num(top_process(tl).effect.nf) for tl in tiles], 2)
            setparam("nt", [num(top_process(tl).effect.nt) for tl in tiles], 2)
        end
        setparam("bot_value", [-num(bottom_process(tl).Qgeo) for tl in tiles], 2)
        # upper boundary temperature: an Input served by Interpolated1D (IO/forcings/providers.jl:14-33), or a constant
        Tsrc = top_process(raw1).T
        if Tsrc isa CryoGrid.Input
            ts, vals = forcing_knots(raw1, Tsrc)
            check(h, ccall((:cgb_set_forcing_table, libcgb), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int64), h, 0, ts, vals, length(ts)))
        else
            check(h, ccall((:cgb_set_forcing_constant, libcgb), Cint, (Ptr{Cvoid}, Int32, Float64), h, 0, num(Tsrc)))
        end
        # freeze curves: kind + van Genuchten / curve parameters per (layer, column); presolver knots of EVERY member and layer with
        # a table map, so that LUT parity does not depend on our table builder (FreezeCurves.initialize!, Soils/soil_heat.jl:26-36)
        any_sfcc = false
        tmap = zeros(Int32, M, nl)
        ntab = 0
        for l in 1:nl
            fc = freezecurve_of(tile1, l)
            fc isa FreeWater && continue
            any_sfcc = true
            solver = sublayer(tile1, l).fcsolver
            solver isa SFCCPreSolver || error("CryoGridB200: SFCC solver $(typeof(solver)) of layer $l is not wired in this glue (SFCCPreSolver only)")
            check(h, ccall((:cgb_set_freezecurve, libcgb), Cint, (Ptr{Cvoid}, Int32, Int32, Int32), h, l - 1, fckind(fc), 0))
            for (c, tl) in enumerate(tiles)
                Hk, Tk = presolver_knots(tl, l)
                check(h, ccall((:cgb_set_sfcc_table, libcgb), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Int64, Int32),
                               h, ntab, Hk, Tk, length(Hk), presolver_extrapolation(tl, l)))
                tmap[c, l] = ntab
                ntab += 1
            end
        end
        if any_sfcc
            for (name, get) in (("vg_alpha", fc -> fc.swrc.α), ("vg_n", fc -> fc.swrc.n), ("theta_res", fc -> fc.swrc.vol.θres),
                                ("Tm", fc -> fc.freezethaw.Tₘ), ("beta", fc -> fc isa PainterKarra ? fc.β : 1.0), ("omega", fc -> fc isa PainterKarra ? fc.ω : 1.0))
                setparam(name, [sfcc_value(freezecurve_of(tl, l), get) for tl in tiles, l in 1:nl], 3)
            end
            check(h, ccall((:cgb_set_sfcc_table_map, libcgb), Cint, (Ptr{Cvoid}, Ptr{Int32}), h, vec(tmap)))
        end
        u0 = reduce(hcat, [collect(Float64, p.u0.H) for p in probs])'   # M x N -> memory order (cell, col) column-fastest
        check(h, ccall((:cgb_set_state, libcgb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Float64), h, collect(u0), t0))
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

# --- helpers over CryoGrid internals (v0.23.4) -------------------------------------------------------------------------
using CryoGrid: TemperatureBC, NFactor, GeothermalHeatFlux, FreeWater, SFCCPreSolver, DallAmico, PainterKarra, DallAmicoSalt
using CryoGrid.Utils: pstrip
using CryoGrid.Numerics: edges
import Unitful: ustrip

"Plain Float64 of a parameter / unitful quantity / number."
num(x) = Float64(ustrip(pstrip(x)))

"The member's Tile with every `Param` replaced by its value from `prob.p` (Tiles.materialize, src/Tiles/tile.jl:341-365)."
resolved_tile(prob) = CryoGrid.Tiles.materialize(Tile(prob.f), prob.p, prob.tspan[1])

"Subsurface layer `l` (1-based) of the stratigraphy: `layers(strat)` = (top, sub..., bottom) (src/Tiles/stratigraphy.jl:46-52, :87)."
sublayer(tile, l) = tile.strat[l + 1]
nsublayers(tile) = length(CryoGrid.Tiles.layers(tile.strat)) - 2
top_process(tile) = tile.strat[1].proc                       # Top(proc), src/types.jl:99-103
bottom_process(tile) = tile.strat[length(CryoGrid.Tiles.layers(tile.strat))].proc
soilpara(tile, l) = sublayer(tile, l).para                   # SimpleSoil, src/Physics/Soils/para/simple.jl:8-15
freezecurve_of(tile, l) = soilpara(tile, l).freezecurve

"""
Per-cell 0-based layer index: layer `l` owns the cells from the grid edge at (or just above) its upper boundary down to the next
layer's -- the rule of `subgridinds` (src/Numerics/grid.jl:49-58) / `Tiles.layeridx` (src/Tiles/state.jl:91-131), and the one the
tested Python mirror uses (cryogrid.jl_b200/model.py, Tile.__init__).
"""
function flatten_layers(tile)
    e = collect(Float64, ustrip.(edges(tile.grid)))
    N = length(e) - 1
    bounds = map(b -> Float64(ustrip(pstrip(b))), CryoGrid.Tiles.boundaries(tile.strat))   # (top, sub_1, ..., sub_n, bottom)
    idx = zeros(Int32, N)
    for l in 1:nsublayers(tile)
        s = l == 1 ? 1 : clamp(searchsortedlast(e, bounds[l + 1]), 1, N)
        idx[s:end] .= l - 1
    end
    return idx
end

"(ts, values) of the `Interpolated1D` interpolant behind `input` (src/IO/forcings/providers.jl:14-33): ts in seconds (convert_t)."
function forcing_knots(tile, input::CryoGrid.Input{name}) where {name}
    itp = getproperty(tile.inputs, name)                    # Interpolations.GriddedInterpolation
    return collect(Float64, itp.knots[1]), collect(Float64, itp.coefs)
end

fckind(::DallAmico) = Int32(1)
fckind(::PainterKarra) = Int32(2)
fckind(::DallAmicoSalt) = Int32(3)
sfcc_value(::FreeWater, get) = 0.0
sfcc_value(fc, get) = num(get(fc))

"""
Knots (H, T) of the member's presolver interpolant for layer `l` (FreezeCurves.SFCCPreSolverCache1D, built by
`FreezeCurves.initialize!` during `initialcondition!`, call site src/Physics/Soils/soil_heat.jl:26-36).
"""
function presolver_knots(tile, l)
    f = sublayer(tile, l).fcsolver.cache.f
    itp = hasproperty(f, :itp) ? f.itp : f                   # strip the Extrapolation wrapper
    return collect(Float64, itp.knots[1]), collect(Float64, itp.coefs)
end
"CGB_EXTRAP_FLAT (0) or CGB_EXTRAP_LINE (1), from the interpolant's extrapolation scheme."
function presolver_extrapolation(tile, l)
    f = sublayer(tile, l).fcsolver.cache.f
    return hasproperty(f, :et) && f.et isa CryoGrid.Numerics.Interpolations.Flat ? Int32(0) : Int32(1)
end

end # module
