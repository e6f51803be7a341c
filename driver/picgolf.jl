# picgolf.jl -- Julia-side binding of libpicgolf.so (include/picgolf.h).
#
# A reference script keeps its parameter / initialisation lines and replaces the body of its
# `for t` loop by `step!`.  Example, src/GaussianFixedPointQuiet.jl with the loop on the GPU:
#
#     include("driver/picgolf.jl"); using .PicGolf
#     N=64;P=32N;dt=1/6N;T=2^13;W=32π^2/3;w=W/P*N
#     x=(bitreverse.(0:P-1).+2.0^63)/2.0^64; v=collect(1:P.>P/2).*2 .-1.
#     sim = PicGolf.Sim(scheme=PicGolf.GAUSS_FIXEDPOINT, N=N, P=P, dt=dt, T=T, W=W, w=w,
#                       rtol=4eps(), atol=0.0, half_width=7)
#     PicGolf.set_particles!(sim, x, v)          # or PicGolf.init_quiet!(sim)
#     PicGolf.step!(sim, T)                      # replaces lines 8-15 (minus the plotting)
#     D, sweeps = PicGolf.diagnostics(sim)       # D[t,1:4] exactly as line 11/13 forms it
#     x, v = PicGolf.particles(sim); r, E = PicGolf.fields(sim)
#     # ... the plotting lines 16-21 run unchanged on D
#
# NOTE: Julia is not installed in the build container, so this file is shipped untested; the same
# ABI is exercised through Python ctypes (particleincellcodegolf.jl_b200/__init__.py, tests/).
module PicGolf

const LIB = get(ENV, "PICGOLF_LIB", joinpath(@__DIR__, "..", "particleincellcodegolf.jl_b200", "lib", "libpicgolf.so"))

const NGP_LEAPFROG, GAUSS_LEAPFROG, GAUSS_FIXEDPOINT, CIC_BORIS_2D3V, GAUSS_SIMPSON13 = Int32(1), Int32(2), Int32(3), Int32(4), Int32(5)
const AREA_SIMPSON13, GAUSS_BORIS_1D2V, GAUSS_BORIS_1D2V2S = Int32(6), Int32(7), Int32(8)
const DEPOSIT_AUTO, DEPOSIT_ATOMIC, DEPOSIT_SORTED, DEPOSIT_POLY = Int32(0), Int32(1), Int32(2), Int32(3)

# Mirror of `picgolf_config` (include/picgolf.h) -- field order and types must match.
Base.@kwdef mutable struct Config
    struct_size::Int32 = 0
    scheme::Int32 = GAUSS_FIXEDPOINT
    N::Int64 = 128
    NY::Int64 = 0
    P::Int64 = 4096
    T::Int64 = 1024
    dt::Float64 = 1 / 768
    W::Float64 = 400.0
    w::Float64 = 12.5
    rtol::Float64 = 1e-8
    atol::Float64 = 0.0
    B0::Float64 = 0.0
    half_width::Int32 = 6
    max_sweeps::Int32 = 10
    diag_every::Int32 = 1
    deposit_mode::Int32 = 0
    deterministic::Int32 = 0
    sort_every::Int32 = 0
    device::Int32 = -1
    rank::Int32 = 0
    nranks::Int32 = 1
    reserved_::Int32 = 0
    local_first::Int64 = -1
    local_count::Int64 = -1
    mass_ratio::Float64 = 0.0
    field_history::Int32 = 0     # 2D3V: keep Exs/Eys/phis (Electrostatic2D3V.jl:171-173) on the device
    reserved2_::Int32 = 0
end

struct PicGolfError <: Exception
    code::Int
    msg::String
end

lasterror() = unsafe_string(ccall((:picgolf_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : throw(PicGolfError(rc, lasterror()))

mutable struct Sim
    h::Ptr{Cvoid}
    cfg::Config
    count::Int64
    function Sim(; kw...)
        cfg = Config(; kw...)
        cfg.struct_size = Int32(sizeof(Config))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:picgolf_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h))
        first, count = Ref{Int64}(0), Ref{Int64}(0)
        check(ccall((:picgolf_local_range, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), h[], first, count))
        s = new(h[], cfg, count[])
        finalizer(s -> ccall((:picgolf_destroy, LIB), Cint, (Ptr{Cvoid},), s.h), s)
        return s
    end
end

function set_particles!(s::Sim, x::Vector{Float64}, v::Vector{Float64})
    GC.@preserve x v check(ccall((:picgolf_set_particles, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64),
                                 s.h, x, v, length(x)))
end

function set_particles!(s::Sim, x, y, vx, vy, vz)   # src/Electrostatic2D3V.jl:45-55
    GC.@preserve x y vx vy vz check(ccall((:picgolf_set_particles_2d3v, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
        s.h, x, y, vx, vy, vz, length(x)))
end

function set_particles!(s::Sim, x::Vector{Float64}, vx::Vector{Float64}, vy::Vector{Float64})   # src/NGP1D2V.jl:27-32, NGP1D2V2S.jl:17-20
    GC.@preserve x vx vy check(ccall((:picgolf_set_particles_1d2v, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64), s.h, x, vx, vy, length(x)))
end

init_quiet!(s::Sim) = check(ccall((:picgolf_init_quiet, LIB), Cint, (Ptr{Cvoid},), s.h))
step!(s::Sim, n::Integer=1) = check(ccall((:picgolf_step, LIB), Cint, (Ptr{Cvoid}, Int64), s.h, n))
synchronize(s::Sim) = check(ccall((:picgolf_synchronize, LIB), Cint, (Ptr{Cvoid},), s.h))

function particles(s::Sim)
    x, v = Vector{Float64}(undef, s.count), Vector{Float64}(undef, s.count)
    check(ccall((:picgolf_get_particles, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64), s.h, x, v, s.count))
    return x, v
end

# x, vx, vy of a 1D2V Sim (src/NGP1D2V.jl:30-32; the two-species scheme holds 2P particles, species 1 first)
function particles_1d2v(s::Sim)
    x, vx, vy = (Vector{Float64}(undef, s.count) for _ in 1:3)
    check(ccall((:picgolf_get_particles_1d2v, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64), s.h, x, vx, vy, s.count))
    return x, vx, vy
end

function particles_2d3v(s::Sim)
    a = [Vector{Float64}(undef, s.count) for _ in 1:5]
    check(ccall((:picgolf_get_particles_2d3v, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64),
                s.h, a[1], a[2], a[3], a[4], a[5], s.count))
    return a
end

# Es[N, TO] of src/NGP1D2V.jl:57,64 (time-averaged field of every window of T/TO steps)
function field_history(s::Sim)
    n = Ref{Int64}(0)
    check(ccall((:picgolf_get_field_history, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Int64}), s.h, C_NULL, 0, n))
    Es = zeros(s.cfg.N, max(n[], 1))
    check(ccall((:picgolf_get_field_history, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ref{Int64}), s.h, Es, size(Es, 2), n))
    return Es[:, 1:n[]]
end

function fields(s::Sim)
    r, E = Vector{Float64}(undef, s.cfg.N), Vector{Float64}(undef, s.cfg.N)
    check(ccall((:picgolf_get_fields, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), s.h, r, E))
    return r, E
end

function fields2d(s::Sim)
    n = s.cfg.N * s.cfg.NY
    r, Ex, Ey = (Vector{Float64}(undef, n) for _ in 1:3)
    check(ccall((:picgolf_get_fields_2d, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), s.h, r, Ex, Ey))
    return (reshape(a, s.cfg.N, s.cfg.NY) for a in (r, Ex, Ey))   # column-major, same as Julia
end

# Exs / Eys / phis of src/Electrostatic2D3V.jl:171-173 (which = 0 / 1 / 2) for a Sim created with field_history=1
function snapshots(s::Sim, which::Integer)
    n = Ref{Int64}(0)
    check(ccall((:picgolf_get_snapshots_2d, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Ref{Int64}), s.h, which, C_NULL, 0, n))
    F = zeros(s.cfg.N, s.cfg.NY, max(n[], 1))
    check(ccall((:picgolf_get_snapshots_2d, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Ref{Int64}), s.h, which, F, size(F, 3), n))
    return F[:, :, 1:n[]]
end

function diagnostics(s::Sim)
    rows = Ref{Int64}(0)
    check(ccall((:picgolf_get_diagnostics, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int32}, Ref{Int64}),
                s.h, C_NULL, 0, C_NULL, rows))
    # 5 columns for the 2D3V K[ti,1:5] and for the 1D2V D[ti,1:5] (NGP1D2V.jl:59-61), 4 for the 1D1V scripts
    ncol = s.cfg.scheme in (CIC_BORIS_2D3V, GAUSS_BORIS_1D2V, GAUSS_BORIS_1D2V2S) ? 5 : 4
    D = zeros(max(rows[], 1), ncol)                 # column-major T x ncol: the ABI's layout IS Julia's
    sw = zeros(Int32, max(rows[], 1))
    check(ccall((:picgolf_get_diagnostics, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Int32}, Ref{Int64}),
                s.h, D, size(D, 1), sw, rows))
    return D[1:rows[], :], sw[1:rows[]]
end

# (sorts so far, flushes / window misses, sorts that were fused into the particle passes): the library re-sorts on its own (the
# reference's sortparticles!, Electrostatic2D3V.jl:57-62, has no counterpart a driver must call); this is for curiosity and tuning
function sort_stats(s::Sim)
    a, b, c = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:picgolf_sort_stats, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), s.h, a, b))
    check(ccall((:picgolf_fused_sorts, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), s.h, c))
    return a[], b[], c[]
end

# ------------------------------------------------------------------------------------------------
# PIC2D3V.jl electrostatic path (include/picgolf_es.h): Species / shapes / ElectrostaticField / ElectrostaticDiagnostics.
#
#     using .PicGolf
#     es = PicGolf.ESSim(plasma, field, diagnostics)   # PIC2D3V objects: reads charge, mass, weight, shape, xyv, B0, dt, ...
#     PicGolf.loop!(es, NT)                            # replaces `for t in 0:NT-1; loop!(...); diagnose!(...); end` (src/2D3V.jl:123-126)
#     PicGolf.fetch!(es, plasma, diagnostics)          # xyv, kineticenergy, fieldenergy, momenta, Exs/Eys/phis back into the PIC2D3V objects
# ------------------------------------------------------------------------------------------------
const ES_NGP, ES_AREA, ES_BSPLINE0 = Int32(0), Int32(1), Int32(10)

# Mirror of `picgolf_es_config` (include/picgolf_es.h) -- field order and types must match (256 bytes).
Base.@kwdef mutable struct ESConfig
    struct_size::Int32 = 0
    nspecies::Int32 = 1
    NX::Int64 = 128
    NY::Int64 = 128
    Lx::Float64 = 1.0
    Ly::Float64 = 1.0
    dt::Float64 = 0.0
    B0x::Float64 = 0.0
    B0y::Float64 = 0.0
    B0z::Float64 = 0.0
    NT::Int64 = 1
    ntskip::Int32 = 1
    ngskip::Int32 = 1
    field_accumulate::Int32 = 1      # update! as written (PIC2D3V.jl:294-297)
    field_history::Int32 = 1
    device::Int32 = -1
    rank::Int32 = 0
    nranks::Int32 = 1
    sort_every::Int32 = 0
    species_P::NTuple{4, Int64} = (0, 0, 0, 0)
    species_shape::NTuple{4, Int32} = (0, 0, 0, 0)
    species_charge::NTuple{4, Float64} = (0.0, 0.0, 0.0, 0.0)
    species_mass::NTuple{4, Float64} = (1.0, 1.0, 1.0, 1.0)
    species_weight::NTuple{4, Float64} = (0.0, 0.0, 0.0, 0.0)
end

# shape code of a PIC2D3V.AbstractShape, by type name (no dependency on the PIC2D3V module here)
function shapecode(shape)
    n = string(nameof(typeof(shape)))
    n == "NGPWeighting" && return ES_NGP
    n == "AreaWeighting" && return ES_AREA
    n == "BSplineWeighting" && return ES_BSPLINE0 + Int32(Int(typeof(shape).parameters[1]))
    error("no libpicgolf kernel for shape $(typeof(shape))")
end

pad4(v, z) = ntuple(i -> i <= length(v) ? v[i] : z, 4)

mutable struct ESSim
    h::Ptr{Cvoid}
    cfg::ESConfig
    function ESSim(plasma, field, diagnostics; accumulate=true, history=true, device=-1)
        g = field.gridparams
        NT = length(diagnostics.kineticenergy) * diagnostics.ntskip
        cfg = ESConfig(nspecies=length(plasma), NX=g.NX, NY=g.NY, Lx=g.Lx, Ly=g.Ly, dt=field.boris.dt_2 * 2,
            B0x=field.B0[1], B0y=field.B0[2], B0z=field.B0[3], NT=NT, ntskip=diagnostics.ntskip, ngskip=diagnostics.ngskip,
            field_accumulate=accumulate, field_history=history, device=device,
            species_P=pad4([Int64(size(s.xyv, 2)) for s in plasma], Int64(0)),
            species_shape=pad4([shapecode(s.shape) for s in plasma], Int32(0)),
            species_charge=pad4([s.charge for s in plasma], 0.0), species_mass=pad4([s.mass for s in plasma], 1.0),
            species_weight=pad4([s.weight for s in plasma], 0.0))
        cfg.struct_size = Int32(sizeof(ESConfig))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:picgolf_es_create, LIB), Cint, (Ref{ESConfig}, Ref{Ptr{Cvoid}}), cfg, h))
        es = new(h[], cfg)
        finalizer(e -> ccall((:picgolf_es_destroy, LIB), Cint, (Ptr{Cvoid},), e.h), es)
        for (i, s) in enumerate(plasma)   # Species.xyv is 5 x P column-major: passed as is
            xyv = s.xyv
            GC.@preserve xyv check(ccall((:picgolf_es_set_species_xyv, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64),
                                         es.h, i - 1, xyv, size(xyv, 2)))
        end
        return es
    end
end

loop!(es::ESSim, n::Integer=1) = check(ccall((:picgolf_es_step, LIB), Cint, (Ptr{Cvoid}, Int64), es.h, n))

function fetch!(es::ESSim, plasma, diagnostics)
    for (i, s) in enumerate(plasma)
        xyv = s.xyv
        GC.@preserve xyv check(ccall((:picgolf_es_get_species_xyv, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64),
                                     es.h, i - 1, xyv, size(xyv, 2)))
    end
    ND = length(diagnostics.kineticenergy)
    rows = Ref{Int64}(0)
    pm, cm = zeros(3, ND), zeros(3, ND)
    check(ccall((:picgolf_es_get_diagnostics, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ref{Int64}),
                es.h, diagnostics.kineticenergy, diagnostics.fieldenergy, pm, cm, ND, rows))
    for ti in 1:rows[]
        diagnostics.particlemomentum[ti] .= pm[:, ti]
        diagnostics.characteristicmomentum[ti] .= cm[:, ti]
    end
    diagnostics.ti[] = rows[]
    if es.cfg.field_history != 0
        for (w, F) in enumerate((diagnostics.Exs, diagnostics.Eys, diagnostics.ϕs))   # NX/ngskip x NY/ngskip x ND, column-major
            n = Ref{Int64}(0)
            check(ccall((:picgolf_es_get_field_history, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Int64, Ref{Int64}),
                        es.h, w - 1, F, size(F, 3), n))
        end
    end
    return nothing
end

# |fft| omega-k map of a stored history on the GPU (PIC2D3V.jl:1483-1493, Electrostatic2D3V.jl:219-233):
# which 0/1/2 = Exs/Eys/phis, axis 0/1 = kx/ky, mode 1 = abs.(fft(F))[:, 1, :], mode 0 = sum over lines of abs.(fft(F[:, i, :])).
function spectrum(es::ESSim, which::Integer, axis::Integer, mode::Integer)
    n = (axis == 0 ? es.cfg.NX : es.cfg.NY) ÷ es.cfg.ngskip
    Z = zeros(n, es.cfg.NT ÷ es.cfg.ntskip)
    check(ccall((:picgolf_es_spectrum, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Float64}), es.h, which, axis, mode, Z))
    return Z
end

# the same for a history the driver collected itself, e.g. Exs of src/Electrostatic2D3V.jl:171
function wk_spectrum(F::Array{Float64, 3}, axis::Integer, mode::Integer)
    n = size(F, axis + 1)
    Z = zeros(n, size(F, 3))
    check(ccall((:picgolf_stage_wk_spectrum, LIB), Cint, (Ptr{Float64}, Int64, Int64, Int64, Cint, Cint, Ptr{Float64}),
                F, size(F, 1), size(F, 2), size(F, 3), axis, mode, Z))
    return Z
end

end # module
