# gen_golden.jl -- reference-generated golden fixtures for the FreezeCurves.jl boundary of the hot path.
#
# The CPU oracle of this repository (oracle/cg_oracle.c) restates FreezeCurves.jl 0.9 from the published equations because the
# package is not vendored in the CryoGrid.jl tree and Julia is not installed in the build image ("parity unpinned", DESIGN.md 4).
# This script closes that gap on any machine that has Julia:
#
#     julia --project=<CryoGrid.jl checkout, v0.23.4> scripts/gen_golden.jl [outdir = tests/golden]
#
# It writes tests/golden/freezecurves.json, presolver.json, rhs_heat_sfcc.json, rhs_heat_salt.json; commit them and
# tests/test_golden_fixtures.py checks the oracle (CPU) AND the CUDA engine (through the C ABI) against them.  Until the files
# exist that test module skips with an explicit "parity unpinned" message.
#
# Call sites reproduced (CryoGrid.jl v0.23.4): src/Physics/Soils/soil_heat.jl:26-36 (FreezeCurves.initialize!), :57, :106
# (sfcc(T, sat; θsat) under ForwardDiff), :116-140 (SFCCInverseEnthalpyObjective + sfccsolve), src/Physics/Heat/heat_conduction.jl:23
# (FreezeCurves.freewater), src/Physics/Salt/salt_diffusion.jl:19 (sfcc(T, sat, Val{:all}(); θsat, saltconc)),
# examples/heat_sfcc_constantbc.jl, examples/heat_sfcc_salt_constantbc.jl.
using CryoGrid
using CryoGrid: FreezeCurves
using CryoGrid.Numerics: ForwardDiff
using Dates
import JSON3

outdir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden")
mkpath(outdir)
writejson(name, obj) = open(io -> JSON3.write(io, obj), joinpath(outdir, name), "w")
meta() = Dict("CryoGrid" => string(pkgversion(CryoGrid)), "FreezeCurves" => string(pkgversion(FreezeCurves)),
              "julia" => string(VERSION), "generated" => string(now()))

# --------------------------------------------------------------------------------------------------------------------------
# 1. closed forms: θw(T) and ∂θw/∂T on a T x parameter lattice; DallAmicoSalt adds ∂θw/∂c; freewater(H, θtot, L)
# --------------------------------------------------------------------------------------------------------------------------
Ts = vcat(collect(-30.0:2.5:-5.0), collect(-4.0:0.5:-1.0), [-0.5, -0.2, -0.1, -0.05, -0.02, -0.01, -0.005, -0.001, 0.0, 0.001, 0.5, 1.0, 5.0])
swrcs = [(1.0, 2.0), (2.0, 1.6), (4.06, 2.03), (0.5, 1.3), (0.1, 1.8)]
curves = Any[]
for (α, n) in swrcs, θsat in (0.3, 0.5, 0.7), sat in (1.0, 0.8, 0.5)
    swrc = VanGenuchten(α=α, n=n)
    variants = [
        ("dallamico", Dict("beta" => 1.0, "omega" => 1.0), DallAmico(swrc=swrc)),
        ("painterkarra", Dict("beta" => 1.0, "omega" => 1.0), PainterKarra(swrc=swrc)),
        # β != 1 with the DEFAULT ω (= 1/β): decides the "slope uses T* vs ω·Tm" question of DESIGN.md 4
        ("painterkarra", Dict("beta" => 2.0, "omega" => 0.5), PainterKarra(β=2.0, swrc=swrc)),
        ("painterkarra", Dict("beta" => 0.5, "omega" => 1.0), PainterKarra(β=0.5, ω=1.0, swrc=swrc)),
    ]
    for (kind, extra, f) in variants
        θw = Float64[]; dθ = Float64[]
        for T in Ts
            v, d = let g = x -> f(x, sat; θsat=θsat)
                g(T), ForwardDiff.derivative(g, T)
            end
            push!(θw, v); push!(dθ, d)
        end
        push!(curves, merge(Dict("kind" => kind, "alpha" => α, "n" => n, "theta_sat" => θsat, "sat" => sat, "theta_res" => 0.0,
                                 "Tm" => 0.0, "T" => Ts, "theta_w" => θw, "dtheta_w_dT" => dθ), extra))
    end
end
salt = Any[]
for (α, n) in ((4.06, 2.03), (1.0, 2.0)), θsat in (0.4, 0.25, 0.03), c in (0.0, 100.0, 300.0, 900.0)
    f = DallAmicoSalt(swrc=VanGenuchten(α=α, n=n))
    θw = Float64[]; dT = Float64[]; dc = Float64[]
    for T in Ts
        h = x -> f(x[1], 1.0, Val{:all}(); θsat=θsat, saltconc=x[2]).θw
        push!(θw, h([T, c]))
        g = ForwardDiff.gradient(h, [T, c])
        push!(dT, g[1]); push!(dc, g[2])
    end
    push!(salt, Dict("alpha" => α, "n" => n, "theta_sat" => θsat, "sat" => 1.0, "saltconc" => c, "T" => Ts,
                     "theta_w" => θw, "dtheta_w_dT" => dT, "dtheta_w_dc" => dc))
end
L = 3.34e8
Hs = collect(-5.0e7:1.0e7:2.5e8)
fw = Any[]
for θtot in (0.5, 0.05, 1e-3, 0.0)
    res = [FreezeCurves.freewater(H, θtot, L) for H in Hs]
    # freewater returns θw (older versions) or a tuple/NamedTuple whose first entry is θw
    first_of(r) = r isa Number ? r : first(r)
    push!(fw, Dict("theta_tot" => θtot, "L" => L, "H" => Hs, "theta_w" => [Float64(first_of(r)) for r in res]))
end
writejson("freezecurves.json", Dict("meta" => meta(), "sfcc" => curves, "dallamicosalt" => salt, "freewater" => fw))

# --------------------------------------------------------------------------------------------------------------------------
# 2. presolver (SFCCPreSolverCache1D, the default of Soils/ground.jl:27): T(H), θw, C, ∂θw∂T through the reference's own call
#    sequence on a dense H lattice (the interpolant is piecewise linear: its knots can be read off this lattice), plus the raw
#    knots when the cache exposes them.
# --------------------------------------------------------------------------------------------------------------------------
function presolver_case(name, sfcc, por, satw, org)
    soil = Ground(SimpleSoil(por=por, sat=satw, org=org, freezecurve=sfcc), heat=HeatBalance(:H))
    strat = @Stratigraphy(
        0.0u"m" => Top(ConstantBC(HeatBalance, CryoGrid.Neumann, 0.0u"W/m^2")),
        0.0u"m" => soil,
        10.0u"m" => Bottom(ConstantBC(HeatBalance, CryoGrid.Neumann, 0.0u"W/m^2"))
    )
    grid = Grid((0.0:1.0:10.0)u"m")
    tile = Tile(strat, grid, initializer(:T, -10.0))
    u0, du0 = initialcondition!(tile, (0.0, 86400.0))     # runs initialize_sfccsolver! -> FreezeCurves.initialize!
    layer = tile.strat[2]                                 # layers(strat) = (top, sub..., bottom): the Ground layer
    state = CryoGrid.Tiles.getstate(tile, u0, du0, 0.0)
    lstate = getproperty(state, CryoGrid.Tiles.layernames(tile.strat)[2])
    solver = Soils.fcsolver(layer)
    hc = Soils.sfccheatcap(layer, lstate, 1)
    θsat = Soils.porosity(layer, lstate, 1)
    θwi = Hydrology.watercontent(layer, lstate, 1)
    sat = θwi / θsat
    Lh = layer.heat.prop.L
    Hq = collect(range(-1.6e8, Lh * θwi + 4.0e7, length=6001))
    T = Float64[]; θw = Float64[]; C = Float64[]; dθ = Float64[]
    for H in Hq
        obj = FreezeCurves.SFCCInverseEnthalpyObjective(sfcc, (θsat=θsat,), hc, Lh, H, sat)
        res = FreezeCurves.sfccsolve(obj, solver, H / 4.2e6, Val{true}())
        push!(T, res.T); push!(θw, res.θw); push!(C, res.C); push!(dθ, res.∂θw∂T)
    end
    d = Dict{String,Any}("name" => name, "por" => por, "sat" => satw, "org" => org, "H" => Hq, "T" => T, "theta_w" => θw, "C" => C,
                         "dtheta_w_dT" => dθ, "Tmin" => try Float64(solver.Tmin) catch; nothing end,
                         "errtol" => try Float64(solver.errtol) catch; nothing end)
    try   # raw knots of the H -> T interpolant, if the cache exposes an Interpolations.jl object as `f`
        itp = solver.cache.f
        g = hasproperty(itp, :itp) ? itp.itp : itp      # strip an Extrapolation wrapper
        d["knots_H"] = collect(Float64, g.knots[1]); d["knots_T"] = collect(Float64, g.coefs)
    catch err
        d["knots_error"] = string(err)
    end
    return d
end
pres = Any[
    presolver_case("cfg1b PainterKarra a=1 n=2 por=.5", PainterKarra(swrc=VanGenuchten(α=1.0, n=2.0)), 0.5, 1.0, 0.0),
    presolver_case("cfg4' DallAmico a=2 n=1.6 por=.5", DallAmico(swrc=VanGenuchten(α=2.0, n=1.6)), 0.5, 1.0, 0.0),
    presolver_case("DallAmico a=0.5 n=1.3 por=.35 sat=.8 org=.3", DallAmico(swrc=VanGenuchten(α=0.5, n=1.3)), 0.35, 0.8, 0.3),
]
writejson("presolver.json", Dict("meta" => meta(), "cases" => pres))

# --------------------------------------------------------------------------------------------------------------------------
# 3. one RHS evaluation + a 10-day solve of examples/heat_sfcc_constantbc.jl WITH the SFCC actually passed to SimpleSoil
#    (the example as written keeps the default FreeWater, SURVEY 8c oddity 1) -- BASELINE cfg1b
# --------------------------------------------------------------------------------------------------------------------------
let
    grid = CryoGrid.DefaultGrid_10cm
    tempprofile = TemperatureProfile(0.0u"m" => -20.0u"°C", 1000.0u"m" => 10.0u"°C")
    sfcc = PainterKarra(swrc=VanGenuchten(α=1.0, n=2.0))
    soilprofile = SoilProfile(0.0u"m" => SimpleSoil(freezecurve=sfcc))
    tile = CryoGrid.SoilHeatTile(Heat.Diffusion1D(:H), ConstantBC(HeatBalance, CryoGrid.Neumann, 10.0u"W/m^2"),
                                 ConstantBC(HeatBalance, CryoGrid.Neumann, 10.0u"W/m^2"), soilprofile, inputs(),
                                 initializer(:T, tempprofile); grid)
    tspan = (DateTime(2010, 1, 1), DateTime(2010, 1, 11))
    u0, du0 = initialcondition!(tile, tspan)
    vars = Dict{String,Any}()
    for v in (:T, :θw, :C, :kc, :k, :∂H∂T, :∂θw∂T, :jH)
        vars[string(v)] = collect(Float64, getvar(v, tile, u0; interp=false))
    end
    prob = CryoGridProblem(tile, u0, tspan, saveat=24 * 3600.0, savevars=(:T, :θw))
    sol = solve(prob, CGEuler(), dt=60.0, dtmax=900.0)
    out = CryoGridOutput(sol)
    writejson("rhs_heat_sfcc.json", Dict("meta" => meta(), "edges" => collect(Float64, ustrip.(edges(tile.grid))),
        "u0_H" => collect(Float64, u0), "du0_dH" => collect(Float64, du0), "vars" => vars,
        "solve" => Dict("dt0" => 60.0, "dtmax" => 900.0, "saveat" => 86400.0, "nsteps" => try sol.stats.nsteps catch; nothing end,
                        "T" => [collect(Float64, ustrip.(out.T[Ti(i)])) for i in 1:size(out.T, 2)],
                        "theta_w" => [collect(Float64, ustrip.(out.θw[Ti(i)])) for i in 1:size(out.θw, 2)],
                        "H_end" => collect(Float64, sol.u[end]))))
end

# --------------------------------------------------------------------------------------------------------------------------
# 4. one RHS evaluation + a 3-day solve of examples/heat_sfcc_salt_constantbc.jl -- BASELINE cfg5
# --------------------------------------------------------------------------------------------------------------------------
let
    sfcc = DallAmicoSalt(swrc=VanGenuchten(α=4.06, n=2.03))
    upperbc = SaltHeatBC(SaltGradient(benthicSalt=900.0, surfaceState=0), ConstantTemperature(0.0))
    strat = @Stratigraphy(
        0.0u"m" => Top(upperbc),
        0.0u"m" => SalineGround(SimpleSoil(freezecurve=sfcc), heat=HeatBalance(:T), salt=SaltMassBalance()),
        1000.0u"m" => Bottom(GeothermalHeatFlux(0.053u"W/m^2"))
    )
    tile = Tile(strat, CryoGrid.DefaultGrid_10cm, initializer(:T, -2.0), initializer(:c, 0.0), SedimentCompactionInitializer(porosityZero=0.4))
    tspan = (DateTime(1990, 1, 1), DateTime(1990, 1, 4))
    u0, du0 = initialcondition!(tile, tspan)
    vars = Dict{String,Any}()
    for v in (:T, :c, :θw, :C, :kc, :k, :∂H∂T, :∂θw∂T, :∂θw∂c, :dₛ, :dₛ_mid, :jH, :jc, :por)
        try
            vars[string(v)] = collect(Float64, getvar(v, tile, u0; interp=false))
        catch err
            vars[string(v) * "_error"] = string(err)
        end
    end
    # a second evaluation at a salty, partly frozen state (the initial state has c = 0 everywhere)
    N = length(cells(tile.grid))
    u1 = copy(u0); du1 = zero(u0)
    Tq = [-3.0 + 2.5 * sin(0.37 * i) for i in 1:N]; cq = [450.0 + 400.0 * cos(0.23 * i) for i in 1:N]
    u1.T .= Tq; u1.c .= cq
    prob =CryoGridProblem(tile, u0, tspan, saveat=24 * 3600.0, savevars=(:T, :θw))
    tile(du1, u1, prob.p, prob.tspan[1])
    vars1 = Dict{String,Any}("T" => Tq, "c" => cq, "dT" => collect(Float64, du1.T), "dc" => collect(Float64, du1.c))
    for v in (:θw, :∂θw∂T, :∂θw∂c, :∂H∂T, :k, :dₛ)
        try vars1[string(v)] = collect(Float64, getvar(v, tile, u1; interp=false)) catch; end
    end
    integrator = init(prob, CGEuler(), dt=60.0)
    for i in integrator end
    out = CryoGridOutput(integrator.sol)
    writejson("rhs_heat_salt.json", Dict("meta" => meta(), "edges" => collect(Float64, ustrip.(edges(tile.grid))),
        "u0" => Dict("T" => collect(Float64, u0.T), "c" => collect(Float64, u0.c)), "du0" => Dict("T" => collect(Float64, du0.T), "c" => collect(Float64, du0.c)),
        "vars" => vars, "state1" => vars1,
        "solve" => Dict("dt0" => 60.0, "saveat" => 86400.0, "u_end" => Dict("T" => collect(Float64, integrator.sol.u[end].T), "c" => collect(Float64, integrator.sol.u[end].c)))))
end
println("golden fixtures written to ", outdir)
