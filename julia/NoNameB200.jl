# NoNameB200.jl -- Julia (>= 1.6) host for libnnpm.so: the binding a NoName maintainer adds.
#
# Keeps the reference's surface -- run_noname() returning gp, gd, p; the conf keys
# ParticleDepositInterpolation / ParticleBackInterpolation = none|cic, ParticleDimensions,
# TopgridDimensions (alias dims), Cosmology (aliases cosm_Ωm / cosm_h); Initializer / Evolve /
# ProblemDefinition naming functions (src/NoName.jl:82-83); restart through restart() -- and moves the
# arithmetic of src/ParticleMesh.jl:289-736, 800-817 behind `ccall`s into the C ABI of include/nnpm.h.
# The time loops, dt control and output cadence stay in Julia, line for line with
# src/ParticleMesh.jl:738-796 and :820-894; particles and mesh stay in HBM between steps and are copied
# back only where the reference writes output (analyzePM) and at the end.
#
# NOT RUNNABLE in the build image (no julia binary); noname_b200/*.py is the call-for-call twin that the
# tests drive, and tests/test_julia_binding.py checks every ccall signature and both struct layouts of this
# file against include/nnpm.h.  Host arrays cross the ABI unchanged: Julia's column-major Array{Float64}
# (rank, Npart) / (N1,N2,N3) is exactly the layout nnpm.h documents.
#
# Deliberately NOT here (out of the hot path, SURVEY.md section 2): Hydro/SPH/FOF, Winston plots, the `Profile` dump.
# Optional packages, used when installed: HDF5.jl -- analyzePM then writes NNNNN.h5 with the groups and datasets of
# IOtools.grid_output / particle_output (src/IOtools.jl:73-116), else a Serialization file under the same names;
# MPI.jl -- conf key NGPUs = P (one Julia process per GPU, `mpiexec -n P julia ...`; the reference has no parallelism).
module NoNameB200

using Printf, Serialization

const HAVE_HDF5 = Base.find_package("HDF5") !== nothing
const HAVE_MPI = Base.find_package("MPI") !== nothing
HAVE_HDF5 && @eval import HDF5
HAVE_MPI && @eval import MPI

export run_noname, noname, evolvePM, evolvePMCosmology, InitializeParticleSimulation, restart, DevicePM,
       deposit, compute_potential, compute_acceleration, interpVecToPoints, drift, kick, make_periodic,
       powerSpectrum

const libnnpm = get(ENV, "NNPM_LIB", joinpath(@__DIR__, "..", "noname_b200", "libnnpm.so"))

# ---- nnpm_config / nnpm_stats (include/nnpm.h) -------------------------------------------------
struct NNPMConfig
    struct_size::Int32
    rank::Int32
    ngrid::NTuple{3,Int64}
    npart_total::Int64
    part_capacity::Int64
    deposit::Int32
    backinterp::Int32
    deposit_mode::Int32
    fixed_bits::Int32
    sort_every::Int32
    bugcompat_interp_z::Int32
    device::Int32
    nranks::Int32
    rank_id::Int32
    timing::Int32
    stream::Ptr{Cvoid}
end

struct NNPMStats
    max_abs_v::Float64
    max_v::Float64
    max_pa::Float64
    n_halo_violations::Int64
    npart_local::Int64
    n_migrated::Int64
    n_untiled::Int64
    ms::NTuple{10,Float32}
    kernel_launches::Int32
    pad::Int32
end

const INTERP = Dict("none" => Int32(0), "cic" => Int32(1))
const FIELD_RHO, FIELD_PHI, FIELD_ACC, FIELD_PA = Int32(0), Int32(1), Int32(2), Int32(3)

# ---- GridTypes (src/GridTypes.jl:4-16) ---------------------------------------------------------
struct GridPatch
    LE; RE; id; level; dim
end
struct GridData
    d::Dict
end

mutable struct DevicePM
    h::Ptr{Cvoid}
    rank::Int
    dims::NTuple{3,Int}
    nranks::Int
    rank_id::Int
end

function check(dev, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:nnpm_last_error, libnnpm), Cstring, (Ptr{Cvoid},), dev === nothing ? C_NULL : dev.h))
    error("nnpm error $rc: $msg")      # the reference throws on bad input too (BoundsError/InexactError)
end

function DevicePM(dims, npart; rank = count(d -> d > 1, dims), deposit = "cic", backinterp = "cic",
                  deposit_mode = 0, sort_every = 8, device = 0, nranks = 1, rank_id = 0, part_capacity = 0,
                  bugcompat_interp_z = false, timing = false)
    ccall((:nnpm_version, libnnpm), Cint, ()) == 2 || error("libnnpm.so: ABI version mismatch")
    cfg = NNPMConfig(Int32(sizeof(NNPMConfig)), Int32(rank), (Int64(dims[1]), Int64(dims[2]), Int64(dims[3])),
                     Int64(npart), Int64(part_capacity), INTERP[deposit], INTERP[backinterp], Int32(deposit_mode), 0,
                     Int32(sort_every), Int32(bugcompat_interp_z), Int32(device), Int32(nranks), Int32(rank_id), Int32(timing), C_NULL)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:nnpm_create, libnnpm), Cint, (Ptr{Ptr{Cvoid}}, Ref{NNPMConfig}), h, cfg)
    check(nothing, rc)
    dev = DevicePM(h[], rank, (dims[1], dims[2], dims[3]), nranks, rank_id)
    finalizer(d -> ccall((:nnpm_destroy, libnnpm), Cint, (Ptr{Cvoid},), d.h), dev)
    return dev
end

# ---- thin wrappers, one per entry point --------------------------------------------------------
upload!(dev, x::Array{Float64,2}, v::Array{Float64,2}, m::Float64; first_id = 0) =
    check(dev, ccall((:nnpm_upload_particles, libnnpm), Cint,
                     (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int64, Int64, Float64), dev.h, x, v, size(x, 2), first_id, m))
download!(dev, x::Array{Float64,2}, v::Array{Float64,2}) =
    check(dev, ccall((:nnpm_download_particles, libnnpm), Cint,
                     (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), dev.h, x, v, C_NULL))
download_ids!(dev, x::Array{Float64,2}, v::Array{Float64,2}, ids::Vector{Int64}) =
    check(dev, ccall((:nnpm_download_particles, libnnpm), Cint,
                     (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}), dev.h, x, v, ids))
local_count(dev) = ccall((:nnpm_local_count, libnnpm), Int64, (Ptr{Cvoid},), dev.h)
device_bytes(dev) = ccall((:nnpm_device_bytes, libnnpm), Int64, (Ptr{Cvoid},), dev.h)
redistribute!(dev) = check(dev, ccall((:nnpm_redistribute, libnnpm), Cint, (Ptr{Cvoid},), dev.h))
sort!(dev) = check(dev, ccall((:nnpm_sort, libnnpm), Cint, (Ptr{Cvoid},), dev.h))
get_field!(dev, which, out::Array{Float64}) =
    check(dev, ccall((:nnpm_get_field, libnnpm), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), dev.h, which, out))
set_field!(dev, which, a::Array{Float64}) =
    check(dev, ccall((:nnpm_set_field, libnnpm), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}), dev.h, which, a))
set_option!(dev, name::String, value) =
    check(dev, ccall((:nnpm_set_option, libnnpm), Cint, (Ptr{Cvoid}, Cstring, Float64), dev.h, name, Float64(value)))
init_synthetic!(dev, kind, pdims, seed, nmodes, kmax, disp_rms, vfac, m) =
    check(dev, ccall((:nnpm_init_synthetic, libnnpm), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Int64}, UInt64, Int32, Int32, Float64, Float64, Float64),
                     dev.h, kind, Int64[pdims...], seed, nmodes, kmax, disp_rms, vfac, m))
init_grf!(dev, pdims, seed, ps_index, disp_rms, vfac, m) =
    check(dev, ccall((:nnpm_init_grf, libnnpm), Cint, (Ptr{Cvoid}, Ptr{Int64}, UInt64, Float64, Float64, Float64, Float64),
                     dev.h, Int64[pdims...], seed, ps_index, disp_rms, vfac, m))
init_grf_peaks!(dev, pdims, seed, ps_index, kpeak::Vector{Float64}, lnwidth::Vector{Float64}, disp_rms, vfac, m) =
    check(dev, ccall((:nnpm_init_grf_peaks, libnnpm), Cint,
                     (Ptr{Cvoid}, Ptr{Int64}, UInt64, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64),
                     dev.h, Int64[pdims...], seed, ps_index, length(kpeak), kpeak, lnwidth, disp_rms, vfac, m))
init_pancake!(dev, pdims, xamp, vamp, m) =
    check(dev, ccall((:nnpm_init_pancake, libnnpm), Cint, (Ptr{Cvoid}, Ptr{Int64}, Float64, Float64, Float64),
                     dev.h, Int64[pdims...], xamp, vamp, m))
dev_make_periodic!(dev) = check(dev, ccall((:nnpm_make_periodic, libnnpm), Cint, (Ptr{Cvoid},), dev.h))
dev_deposit!(dev) = check(dev, ccall((:nnpm_deposit, libnnpm), Cint, (Ptr{Cvoid},), dev.h))
dev_solve!(dev) = check(dev, ccall((:nnpm_solve, libnnpm), Cint, (Ptr{Cvoid},), dev.h))
dev_accel!(dev) = check(dev, ccall((:nnpm_accel, libnnpm), Cint, (Ptr{Cvoid},), dev.h))
dev_interp!(dev) = check(dev, ccall((:nnpm_interp, libnnpm), Cint, (Ptr{Cvoid},), dev.h))
dev_drift!(dev, dt) = check(dev, ccall((:nnpm_drift, libnnpm), Cint, (Ptr{Cvoid}, Float64), dev.h, dt))
dev_kick!(dev, dt) = check(dev, ccall((:nnpm_kick, libnnpm), Cint, (Ptr{Cvoid}, Float64), dev.h, dt))
dev_kick_comoving!(dev, dt, a, dadt) =
    check(dev, ccall((:nnpm_kick_comoving, libnnpm), Cint, (Ptr{Cvoid}, Float64, Float64, Float64), dev.h, dt, a, dadt))

function step_static!(dev, dt)
    st = Ref{NNPMStats}()
    check(dev, ccall((:nnpm_step_static, libnnpm), Cint, (Ptr{Cvoid}, Float64, Ref{NNPMStats}), dev.h, dt, st))
    return reduce_stats!(dev, st)
end
function step_comoving!(dev, dt, a_d1, a_k, adot_k, a_d2)
    st = Ref{NNPMStats}()
    check(dev, ccall((:nnpm_step_comoving, libnnpm), Cint,
                     (Ptr{Cvoid}, Float64, Float64, Float64, Float64, Float64, Ref{NNPMStats}),
                     dev.h, dt, a_d1, a_k, adot_k, a_d2, st))
    return reduce_stats!(dev, st)
end
function reduce_stats!(dev, st::Ref{NNPMStats})       # global maxima when the mesh is slab-decomposed
    dev.nranks > 1 && check(dev, ccall((:nnpm_reduce_stats, libnnpm), Cint, (Ptr{Cvoid}, Ref{NNPMStats}), dev.h, st))
    return st[]
end
function power_spectrum(dev)
    lenk = round(Int, maximum(dev.dims) / 2)
    k = zeros(lenk); p = zeros(lenk)
    check(dev, ccall((:nnpm_power_spectrum, libnnpm), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32), dev.h, k, p, lenk))
    return k, p
end

# Multi-GPU (new; Readme.md:53 "No kind of parallelism so far"): one Julia process per GPU.  `bcast` is the host's
# broadcast of 128 bytes from rank 0 (e.g. b -> MPI.Bcast!(b, 0, comm)).
function comm_init!(dev, bcast)
    id = zeros(UInt8, 128)
    dev.rank_id == 0 && ccall((:nnpm_comm_unique_id, libnnpm), Cint, (Ptr{UInt8},), id) != 0 && error("nnpm_comm_unique_id failed")
    bcast(id)
    check(dev, ccall((:nnpm_comm_init, libnnpm), Cint, (Ptr{Cvoid}, Ptr{UInt8}), dev.h, id))
end
# pinned host staging for large outputs
function host_alloc(bytes)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    ccall((:nnpm_host_alloc, libnnpm), Cint, (Ptr{Ptr{Cvoid}}, UInt64), p, bytes) == 0 || error("nnpm_host_alloc failed")
    return p[]
end
host_free(p) = ccall((:nnpm_host_free, libnnpm), Cint, (Ptr{Cvoid},), p)

# ---- operator-granular drop-ins (same names / argument order as src/ParticleMesh.jl) ----------
# They take the reference's host arrays, run the device operator and write the first argument.
size3(a) = (size(a, 1), size(a, 2), size(a, 3))
function deposit(rho, x, m; interpolation = "none")                       # :289
    dev = DevicePM(size3(rho), size(x, 2); rank = size(x, 1), deposit = interpolation, sort_every = 0)
    upload!(dev, x, zeros(size(x)), Float64(m))
    dev_deposit!(dev)
    get_field!(dev, FIELD_RHO, rho); nothing
end
function compute_potential(phi, rho, rt)                                  # :417 (rt: caller scratch, unused)
    dev = DevicePM(size3(rho), 1; rank = count(d -> d > 1, size3(rho)))
    set_field!(dev, FIELD_RHO, rho)
    dev_solve!(dev)
    get_field!(dev, FIELD_PHI, phi); nothing
end
function compute_acceleration(a, phi)                                     # :631
    dev = DevicePM(size3(phi), 1; rank = count(d -> d > 1, size3(phi)))
    set_field!(dev, FIELD_PHI, phi)
    dev_accel!(dev)
    get_field!(dev, FIELD_ACC, a); nothing
end
function interpVecToPoints(pa, a, x; interpolation = "none")              # :640
    dev = DevicePM(size(a)[2:4], size(x, 2); rank = size(x, 1), backinterp = interpolation, sort_every = 0)
    upload!(dev, x, zeros(size(x)), 1.0)
    set_field!(dev, FIELD_ACC, a)
    dev_interp!(dev)
    get_field!(dev, FIELD_PA, pa); nothing
end
function drift(x, v, dt)                                                  # :723
    dev = DevicePM((4, 4, size(x, 1) == 3 ? 4 : 1), size(x, 2); rank = size(x, 1))
    upload!(dev, x, v, 1.0)
    dev_drift!(dev, dt)
    download!(dev, x, similar(v)); nothing
end
function kick(v, pa, dt, a = nothing, dadt = nothing)                     # :731 / :800
    dev = DevicePM((4, 4, size(v, 1) == 3 ? 4 : 1), size(v, 2); rank = size(v, 1))
    upload!(dev, v, v, 1.0)
    set_field!(dev, FIELD_PA, pa)
    a === nothing ? dev_kick!(dev, dt) : dev_kick_comoving!(dev, dt, a, dadt)
    download!(dev, similar(v), v); nothing
end
function make_periodic(x)                                                 # :409
    dev = DevicePM((4, 4, size(x, 1) == 3 ? 4 : 1), size(x, 2); rank = size(x, 1))
    upload!(dev, x, zeros(size(x)), 1.0)
    dev_make_periodic!(dev)
    download!(dev, x, similar(x)); nothing
end
function powerSpectrum(gp, gd, p, conf)                                   # :65-100
    make_periodic(p["x"])
    dev = DevicePM(gp[1].dim, size(p["x"], 2); rank = size(p["x"], 1), deposit = conf["ParticleDepositInterpolation"], sort_every = 0)
    upload!(dev, p["x"], p["v"], Float64(p["m"]))
    return power_spectrum(dev)
end

# ---- conf files (AppConf-style `key = value`; src/NoName.jl:58-72, src/default.conf) -----------
function parse_value(s::AbstractString)
    s = strip(s)
    s == "true" && return true
    s == "false" && return false
    (startswith(s, "\"") && endswith(s, "\"")) && return String(s[2:end-1])
    v = tryparse(Int, replace(s, "_" => ""));     v !== nothing && return v
    f = tryparse(Float64, replace(s, "_" => "")); f !== nothing && return f
    if (startswith(s, "[") && endswith(s, "]")) || (startswith(s, "(") && endswith(s, ")"))
        return Any[parse_value(q) for q in split_top(s[2:end-1])]
    end
    return String(s)                               # bare word (function name, cic) or constructor call
end
function split_top(s)
    parts = String[]; depth = 0; cur = IOBuffer()
    for ch in s
        ch in "([{" && (depth += 1)
        ch in ")]}" && (depth -= 1)
        if ch == ',' && depth == 0
            push!(parts, String(take!(cur)))
        else
            write(cur, ch)
        end
    end
    t = strip(String(take!(cur)))
    isempty(t) || push!(parts, t)
    return parts
end
function parseconf(path)
    out = Dict{String,Any}()
    for raw in eachline(path)
        line = strip(first(split(raw, '#')))
        (isempty(line) || !occursin('=', line)) && continue
        k, v = split(line, '='; limit = 2)
        out[strip(k)] = parse_value(v)
    end
    return out
end
const DEFAULTS = Dict{String,Any}(            # verbatim values of the reference's src/default.conf
    "dims" => [30, 1, 1], "Nghosts" => [3, 3, 3], "DomainLeftEdge" => [0.0, 0.0, 0.0], "DomainRightEdge" => [1.0, 1.0, 1.0],
    "CourantFactor" => 0.3, "CurrentOutputNumber" => 0, "OutputDirectory" => "/tmp/noname", "OutputSeparateDirectories" => true,
    "LastOutputCycle" => -1, "LastOutputTime" => 0.0, "OutputEveryNCycle" => 0, "OutputDtDataDump" => 0.0, "verbose" => false,
    "RestartFileName" => "/tmp/restart.jld", "ProfilingResultsFile" => "profiling_results.bin", "InitialDt" => 0.0000001,
    "DtFractionOfTime" => 10000.0, "MaximumDt" => 1e30, "StartTime" => 0, "CurrentTime" => 0, "StopTime" => 0, "StartCycle" => 0,
    "CurrentCycle" => 0, "StopCycle" => 1000000, "ParticleDimensions" => [1, 1, 1], "ParticleDepositInterpolation" => "cic",
    "ParticleBackInterpolation" => "cic")
function load_conf(path; overrides = Dict{String,Any}())
    conf = copy(DEFAULTS)
    user = joinpath(get(ENV, "HOME", "."), ".noname", "default.conf")
    isfile(user) && merge!(conf, parseconf(user))                          # :61-66
    file = path === nothing ? Dict{String,Any}() : parseconf(path)
    merge!(conf, file, overrides)
    if !haskey(conf, "TopgridDimensions") && (haskey(file, "dims") || haskey(overrides, "dims"))
        conf["TopgridDimensions"] = conf["dims"]                           # alias named by the north star
    end
    if !haskey(conf, "Cosmology") && (haskey(conf, "cosm_Ωm") || haskey(conf, "cosm_h"))
        conf["Cosmology"] = "cosmology(h=$(get(conf, "cosm_h", 0.69)),OmegaM=$(get(conf, "cosm_Ωm", 0.29)),OmegaR=0.0)"
    end
    return conf
end

# ---- cosmology scalars (src/cosmology.jl), closed forms instead of fzero on a quadrature ------
struct Cosmo
    h::Float64; Ω_m::Float64; Ω_Λ::Float64; Ω_r::Float64
end
function parse_cosmology(expr)
    m = match(r"^\s*cosmology\((.*)\)\s*$", expr)
    m === nothing && error("Cosmology must be cosmology(h=..,OmegaM=..,OmegaR=..): $expr")
    kw = Dict{String,Float64}()
    for q in split(m.captures[1], ','); k, v = split(q, '='); kw[strip(k)] = parse(Float64, strip(v)); end
    Ωm, Ωr = get(kw, "OmegaM", 0.27), get(kw, "OmegaR", 0.0)
    Ωr == 0.0 || error("OmegaR != 0 needs the quadrature of Cosmology.jl; the shipped confs state OmegaR = 0.0")
    return Cosmo(get(kw, "h", 0.69), Ωm, 1.0 - get(kw, "OmegaK", 0.0) - Ωm - Ωr, Ωr)
end
const gyr_in_s = 1.e9 * 31_556_952                                            # :4
const hubble_time_gyr = 9.77813952                                            # Cosmology.jl: hubble_time_gyr0(c) = 9.77813952 / c.h
function age_gyr(c::Cosmo, z)                                                 # flat matter + Λ, closed form
    a = 1.0 / (1.0 + z); tH = hubble_time_gyr / c.h
    c.Ω_Λ == 0.0 && return tH * (2.0 / 3.0) * a^1.5 / sqrt(c.Ω_m)
    return tH * (2.0 / (3.0 * sqrt(c.Ω_Λ))) * asinh(sqrt(c.Ω_Λ / c.Ω_m) * a^1.5)
end
function z_from_t(c::Cosmo, t)                                                # :7-11 (inverse of age_gyr)
    tH = hubble_time_gyr / c.h
    a = c.Ω_Λ == 0.0 ? (t * 1.5 * sqrt(c.Ω_m) / tH)^(2.0 / 3.0) :
        (sqrt(c.Ω_m / c.Ω_Λ) * sinh(1.5 * sqrt(c.Ω_Λ) * t / tH))^(2.0 / 3.0)
    return 1.0 / a - 1.0
end
unit_T(c::Cosmo, zinit) = 2.519445e17 / sqrt(c.Ω_m) / c.h / (1 + zinit)^1.5   # get_units(...).T  :61-72
a_from_t(c, t; zwherea1 = 0.0) = (1.0 + zwherea1) / (1.0 + z_from_t(c, t))    # :13-17
a_from_t_internal(c, t, zinit; zwherea1 = 0.0) = a_from_t(c, t * (unit_T(c, zinit) / gyr_in_s), zwherea1 = zwherea1)  # :19-23
z_from_a(c, a; zwherea1 = 0.0) = (1.0 + zwherea1) / a - 1.0
z_from_t_internal(c, t, zinit) = z_from_a(c, a_from_t_internal(c, t, zinit))  # :25-28
function dadt(c, a; zwherea1 = 0.0)                                           # :30-37
    tv = a / (1.0 + zwherea1)
    O_k = 1 - c.Ω_m - c.Ω_Λ - c.Ω_r
    return sqrt(2.0 / (3.0 * c.Ω_m * a) * (c.Ω_m + O_k * tv + c.Ω_Λ * tv^3))
end
dadt_from_t_internal(c, t, zinit) = dadt(c, a_from_t_internal(c, t, zinit, zwherea1 = zinit); zwherea1 = zinit)   # :39-42

# ---- initialisation (src/ParticleMesh.jl:103-287) ----------------------------------------------
function initialize_particles_uniform(x, pd)                                  # :257-287, i fastest
    rank = size(x, 1); c = 1
    for k in 1:(rank == 3 ? pd[3] : 1), j in 1:pd[2], i in 1:pd[1]
        x[1, c] = i; x[2, c] = j; rank == 3 && (x[3, c] = k); c += 1
    end
    for d in 1:rank; x[d, :] .= (x[d, :] .- 0.5) ./ pd[d]; end
    nothing
end
UniformParticles(gp, gd, p, conf) = initialize_particles_uniform(p["x"], conf["ParticleDimensions"])
function UniformParticleWind(gp, gd, p, conf)
    initialize_particles_uniform(p["x"], conf["ParticleDimensions"])
    p["v"][1, :] .= conf["UniformParticleWindVelocity"]; nothing
end
function pancake_amplitudes(conf)                                             # :178-191
    zc, kZ, zinit, cosmo = conf["ZeldovichCausticRedshift"], 1.0, conf["InitialRedshift"], conf["_cosmo"]
    A = 5.0 * (1.0 + zc) / (2kZ) * 1.0 / (1.0 + zinit)
    ȧ = dadt_from_t_internal(cosmo, conf["StartTime"], zinit)
    a = a_from_t_internal(cosmo, conf["StartTime"], zinit, zwherea1 = zinit)
    return 2.0 / 5 * A * a, 2.0 / 5 * A * ȧ
end
# generated on the device: only the request is recorded; p["x"] / p["v"] are filled at the first output
ZeldovichPancakeParticles(gp, gd, p, conf) = (p["_ic"] = (:pancake, pancake_amplitudes(conf)); nothing)   # :177-198
# InputPowerSpectrum = scaleFreePS(n, norm) | multiplePeaksPS(n, norm, [k...], lnwidth | [w...])   (src/InputPowerSpectra.jl:1-24)
function parse_input_power_spectrum(expr)
    m = match(r"^\s*(scaleFreePS|multiplePeaksPS)\s*\((.*)\)\s*$", string(expr))
    m === nothing && error("InputPowerSpectrum = $expr: scaleFreePS(n, norm) or multiplePeaksPS(n, norm, k, lnwidth) required")
    args = Any[parse_value(q) for q in split_top(m.captures[2])]
    m.captures[1] == "scaleFreePS" && return Float64(args[1]), Float64(args[2]), nothing
    ks = Float64[q for q in args[3]]
    ws = args[4] isa AbstractVector ? Float64[q for q in args[4]] : fill(Float64(args[4]), length(ks))   # :20-24
    length(ws) == length(ks) || error("multiplePeaksPS: one width per peak (or a single number)")
    return Float64(args[1]), Float64(args[2]), (ks, ws)
end
function PowerSpectrumParticles(gp, gd, p, conf)                              # :209-254 (seeded, on the device)
    haskey(conf, "InputPowerSpectrum") || error("Requested to use PowerSpectrumParticles but did not specify InputPowerSpectrum")
    n, _, peaks = parse_input_power_spectrum(conf["InputPowerSpectrum"])
    vfac = 0.0
    if haskey(conf, "_cosmo") && haskey(conf, "InitialRedshift")
        zi = conf["InitialRedshift"]
        vfac = dadt_from_t_internal(conf["_cosmo"], conf["StartTime"], zi) / a_from_t_internal(conf["_cosmo"], conf["StartTime"], zi, zwherea1 = zi)
    end
    seed = UInt64(get(conf, "RandomSeed", 12345))
    rms = Float64(get(conf, "DisplacementRMS", 0.3 / maximum(conf["TopgridDimensions"])))
    p["_ic"] = peaks === nothing ? (:grf, (seed, n, rms, vfac)) : (:grfpeaks, (seed, n, peaks[1], peaks[2], rms, vfac)); nothing
end
const PROBLEMS = Dict("UniformParticles" => UniformParticles, "UniformParticleWind" => UniformParticleWind,
                      "ZeldovichPancakeParticles" => ZeldovichPancakeParticles, "PowerSpectrumParticles" => PowerSpectrumParticles)

function InitializeParticleSimulation(gp, gd, p, conf)                        # :103-159
    tgd = Int.(conf["TopgridDimensions"]); pd = Int.(conf["ParticleDimensions"])
    rank = sum(tgd .> 1); conf["_rank"] = rank
    Npart = prod(pd); dim = (tgd...,)
    gp[1] = GridPatch(conf["DomainLeftEdge"], conf["DomainRightEdge"], 1, 1, dim)
    gd[1] = GridData(Dict("ρD" => ones(dim), "Φ" => ones(dim)))
    device_ic = conf["ProblemDefinition"] in ("ZeldovichPancakeParticles", "PowerSpectrumParticles")
    p["x"] = device_ic ? nothing : zeros(rank, Npart)
    p["v"] = device_ic ? nothing : zeros(rank, Npart)
    p["m"] = 1.0 * prod(dim) / Npart
    if haskey(conf, "Cosmology")
        cosmo = parse_cosmology(conf["Cosmology"]); conf["_cosmo"] = cosmo
        if haskey(conf, "InitialRedshift")
            zstart = conf["InitialRedshift"]; T = unit_T(cosmo, zstart)
            conf["CurrentTime"] = conf["StartTime"] = age_gyr(cosmo, zstart) * (gyr_in_s / T)
            conf["StopTime"] = age_gyr(cosmo, conf["FinalRedshift"]) * (gyr_in_s / T)
            conf["CurrentRedshift"] = zstart
            conf["OmegaCDM"] = cosmo.Ω_m - get(conf, "OmegaBaryon", 0.0)
            p["m"] *= conf["OmegaCDM"]
        end
    end
    PROBLEMS[conf["ProblemDefinition"]](gp, gd, p, conf)
    nothing
end

# ---- output hook (src/ParticleMesh.jl:13-63, src/IOtools.jl:40-59) -----------------------------
function readyForOutput(c)
    output = (c["OutputEveryNCycle"] > 0 && mod(c["CurrentCycle"], c["OutputEveryNCycle"]) == 0) ||
             (c["OutputDtDataDump"] != 0 && abs(c["CurrentTime"] - c["LastOutputTime"]) >= c["OutputDtDataDump"])
    if output
        c["LastOutputTime"] = c["CurrentTime"]; c["LastOutputCycle"] = c["CurrentCycle"]
    end
    return output
end
output_due(c, t, cycle) = !get(c, "OutputDisabled", false) &&
    ((c["OutputEveryNCycle"] > 0 && mod(cycle, c["OutputEveryNCycle"]) == 0) ||
     (c["OutputDtDataDump"] != 0 && abs(t - c["LastOutputTime"]) >= c["OutputDtDataDump"]))

"device -> host copy of what the reference keeps in gd / p; gd[1].d[\"ρD\"] is the density of the last deposit.
With NGPUs > 1 the slabs / particle shares of all ranks are gathered: every rank ends up with the complete arrays."
function sync_host!(dev, gd, p)
    n = local_count(dev)
    if dev.nranks == 1
        if p["x"] === nothing || size(p["x"], 2) != n
            p["x"] = zeros(dev.rank, n); p["v"] = zeros(dev.rank, n)
        end
        download!(dev, p["x"], p["v"])
        get_field!(dev, FIELD_PHI, gd[1].d["Φ"])
        get_field!(dev, FIELD_RHO, gd[1].d["ρD"])       # kept by the step (option keep_rho)
        return nothing
    end
    comm = MPI.COMM_WORLD
    xl = zeros(dev.rank, n); vl = zeros(dev.rank, n); ids = zeros(Int64, n)
    download_ids!(dev, xl, vl, ids)
    counts = MPI.Allgather(Int32(n), comm)               # particles held by every rank (Int32 counts: < 2^31 / rank each)
    ntot = Int(sum(counts))
    idall = zeros(Int64, ntot); xall = zeros(dev.rank, ntot); vall = zeros(dev.rank, ntot)
    MPI.Allgatherv!(ids, MPI.VBuffer(idall, counts), comm)
    MPI.Allgatherv!(xl, MPI.VBuffer(xall, counts .* Int32(dev.rank)), comm)
    MPI.Allgatherv!(vl, MPI.VBuffer(vall, counts .* Int32(dev.rank)), comm)
    if p["x"] === nothing || size(p["x"], 2) != ntot
        p["x"] = zeros(dev.rank, ntot); p["v"] = zeros(dev.rank, ntot)
    end
    p["x"][:, idall .+ 1] = xall                         # the reference's particle order (ids are 0-based)
    p["v"][:, idall .+ 1] = vall
    N1, N2, N3 = dev.dims
    slab = zeros(N1, N2, div(N3, dev.nranks))            # slabs are contiguous in k, the slowest Julia index
    for (which, key) in ((FIELD_PHI, "Φ"), (FIELD_RHO, "ρD"))
        get_field!(dev, which, slab)
        MPI.Allgather!(slab, MPI.UBuffer(gd[1].d[key], length(slab)), comm)
    end
    nothing
end

# IOtools.grid_output + particle_output (src/IOtools.jl:73-116) through HDF5.jl
function hdf5_writer(fname, gp, gd, p)
    HDF5.h5open(fname * ".h5", "w") do f
        for (_, cgp) in gp
            g = HDF5.create_group(f, string("grid", cgp.id))
            for (name, a) in gd[cgp.id].d
                g[name] = dropdims(a; dims = Tuple(i for i in 1:ndims(a) if size(a, i) == 1))   # IOtools.squeeze
                HDF5.attributes(g[name])["vsz_range"] = Float64[gp[1].LE[1], gp[1].LE[2], gp[1].RE[1], gp[1].RE[2]]
            end
        end
        g = HDF5.create_group(f, "particles")
        rank = size(p["x"], 1)
        g["x"] = p["x"][1, :]; g["vx"] = p["v"][1, :]
        if rank > 1; g["y"] = p["x"][2, :]; g["vy"] = p["v"][2, :]; end
        if rank > 2; g["z"] = p["x"][3, :]; g["vz"] = p["v"][3, :]; end
        g["m"] = p["m"]
    end
    nothing
end
serialization_writer(fname, gp, gd, p) = open(io -> serialize(io, Dict("grid1/ρD" => gd[1].d["ρD"], "grid1/Φ" => gd[1].d["Φ"],
                                                                       "particles" => p)), fname * ".jls", "w")
default_writer(fname, gp, gd, p) = HAVE_HDF5 ? hdf5_writer(fname, gp, gd, p) : serialization_writer(fname, gp, gd, p)

function analyzePM(gp, gd, p, conf, dev; force = false, writer = default_writer)
    (!readyForOutput(conf) && !force) && return nothing
    get(conf, "OutputDisabled", false) && return nothing
    sync_host!(dev, gd, p)
    c = conf
    pk = get(c, "OutputPowerSpectrum", false) === true ? power_spectrum(dev) : nothing   # :47-61 (numbers, not a plot)
    if dev.nranks > 1 && dev.rank_id != 0
        c["CurrentOutputNumber"] += 1                    # same counters on every rank; rank 0 writes the files
        return nothing
    end
    isdir(c["OutputDirectory"]) || mkpath(c["OutputDirectory"])
    s = @sprintf("%5.5i", c["CurrentOutputNumber"])
    newD = c["OutputSeparateDirectories"] ? joinpath(c["OutputDirectory"], string(get(c, "OutputDirPrefix", ""), s)) : c["OutputDirectory"]
    isdir(newD) || mkpath(newD)
    open(joinpath(newD, s * "_parameters.info"), "w") do f
        for (k, v) in c; startswith(k, "_") || write(f, string(k, " = ", v, "\n")); end
    end
    c["CurrentOutputNumber"] += 1
    writer(joinpath(newD, s), gp, gd, p)
    if pk !== nothing
        open(joinpath(newD, s * "_PS.txt"), "w") do f
            println(f, "# k P(k)   time = ", c["CurrentTime"])
            for (k, q) in zip(pk[1], pk[2]); println(f, k, " ", q); end
        end
    end
    nothing
end

# ProfilingOn = true (src/NoName.jl:86-93 wraps the run in Julia's profiler): accumulate the per-phase CUDA-event
# milliseconds of every step instead; noname() writes them to ProfilingResultsFile
const MS_NAMES = ("deposit", "solve", "push", "sort", "exchange", "comm", "fft_z", "fft_xy", "fft_x", "fft_y")
function profile_step!(conf, st::NNPMStats)
    get(conf, "ProfilingOn", false) === true || return nothing
    prof = get!(conf, "_profile", Dict{String,Any}("steps" => 0, "kernel_launches" => 0, "ms" => Dict{String,Float64}()))
    prof["steps"] += 1; prof["kernel_launches"] += Int(st.kernel_launches)
    for (i, name) in enumerate(MS_NAMES); prof["ms"][name] = get(prof["ms"], name, 0.0) + Float64(st.ms[i]); end
    nothing
end
function write_profiling_results(fname, prof)
    open(fname, "w") do f
        println(f, "steps = ", prof["steps"]); println(f, "kernel_launches = ", prof["kernel_launches"])
        for (k, v) in prof["ms"]; println(f, "ms_per_step[", k, "] = ", v / max(1, prof["steps"])); end
    end
end

# ---- evolve loops: conf["Evolve"] plugins ------------------------------------------------------
function device_for(gp, p, conf)
    P = Int(get(conf, "NGPUs", 1)); rid = 0; device = Int(get(conf, "Device", 0))
    npart = prod(Int.(conf["ParticleDimensions"]))
    if P > 1
        HAVE_MPI || error("NGPUs = $P needs MPI.jl (one Julia process per GPU: mpiexec -n $P julia ...)")
        MPI.Initialized() || MPI.Init()
        comm = MPI.COMM_WORLD
        MPI.Comm_size(comm) == P || error("NGPUs = $P needs $P processes; this job has $(MPI.Comm_size(comm))")
        rid = MPI.Comm_rank(comm); device = rid % P      # one node: local rank == rank
    end
    cap = (P > 1 && haskey(conf, "PartCapacityFactor")) ? Int(floor(Float64(conf["PartCapacityFactor"]) * npart / P)) + 4096 : 0
    dev = DevicePM(gp[1].dim, npart; rank = conf["_rank"],
                   deposit = conf["ParticleDepositInterpolation"], backinterp = conf["ParticleBackInterpolation"],
                   sort_every = get(conf, "SortEvery", 8), deposit_mode = get(conf, "DepositMode", "atomic") == "fixedpoint" ? 1 : 0,
                   device = device, bugcompat_interp_z = get(conf, "BackInterpBugCompat", false),
                   nranks = P, rank_id = rid, part_capacity = cap, timing = get(conf, "ProfilingOn", false) === true)
    P > 1 && comm_init!(dev, id -> MPI.Bcast!(id, 0, MPI.COMM_WORLD))
    ic = get(p, "_ic", nothing)
    pd = Int.(conf["ParticleDimensions"])
    if ic !== nothing && p["x"] === nothing
        if ic[1] == :grf
            init_grf!(dev, pd, ic[2]..., Float64(p["m"]))
        elseif ic[1] == :grfpeaks
            init_grf_peaks!(dev, pd, ic[2]..., Float64(p["m"]))
        else
            init_pancake!(dev, pd, ic[2]..., Float64(p["m"]))
        end
    elseif P > 1
        # every rank holds the full host arrays (the initialisers ran on each): upload one contiguous share, then route
        lo, hi = div(rid * npart, P), div((rid + 1) * npart, P)
        upload!(dev, p["x"][:, lo+1:hi], p["v"][:, lo+1:hi], Float64(p["m"]); first_id = lo)
        redistribute!(dev)
    else
        upload!(dev, p["x"], p["v"], Float64(p["m"]))
    end
    return dev
end

"evolvePM, src/ParticleMesh.jl:738-796"
function evolvePM(gp, gd, p, conf)
    c = conf
    dev = device_for(gp, p, c)
    dx = 1.0 / maximum(gp[1].dim)                                            # :759
    dt = c["InitialDt"]
    c["CurrentTime"] = c["StartTime"]; c["CurrentCycle"] = c["StartCycle"]   # :762-763
    while c["CurrentTime"] < c["StopTime"] && c["CurrentCycle"] < c["StopCycle"]
        t1, cyc1 = c["CurrentTime"] + dt, c["CurrentCycle"] + 1
        set_option!(dev, "keep_rho", output_due(c, t1, cyc1) || !(t1 < c["StopTime"] && cyc1 < c["StopCycle"]))
        st = step_static!(dev, dt)                                           # :766-778
        profile_step!(c, st)
        c["CurrentTime"] += dt
        dt = c["CourantFactor"] * dx / (st.max_abs_v + 1e-30)                # :782
        dt = min(dt, c["MaximumDt"])
        if c["CurrentTime"] + dt > c["StopTime"]
            dt = (1.0 + 1e-15) * (c["StopTime"] - c["CurrentTime"])          # :785-788
            println("final dt = ", dt)
        end
        c["CurrentCycle"] += 1
        analyzePM(gp, gd, p, c, dev)                                         # copies back only when an output is due
    end
    sync_host!(dev, gd, p)
    nothing
end

"evolvePMCosmology, src/ParticleMesh.jl:820-894; cosmology scalars as in src/cosmology.jl"
function evolvePMCosmology(gp, gd, p, conf)
    c = conf; cosmo = c["_cosmo"]
    dev = device_for(gp, p, c)
    dx = 1.0 / maximum(gp[1].dim)
    dt = c["InitialDt"]; zinit = c["InitialRedshift"]
    c["CurrentTime"] = c["StartTime"]; c["CurrentCycle"] = c["StartCycle"]
    c["LastOutputTime"] = c["CurrentTime"] - c["OutputDtDataDump"]           # :848
    while c["CurrentTime"] < c["StopTime"] && c["CurrentCycle"] < c["StopCycle"]
        t = c["CurrentTime"]
        a1 = a_from_t_internal(cosmo, t + dt / 2 / 2, zinit, zwherea1 = zinit)       # :853
        ak = a_from_t_internal(cosmo, t + dt / 2, zinit, zwherea1 = zinit)           # :863
        adk = dadt_from_t_internal(cosmo, t + dt / 2, zinit)                         # :864
        a2 = a_from_t_internal(cosmo, t + 3.0 * dt / 2 / 2, zinit, zwherea1 = zinit) # :866
        t1, cyc1 = t + dt, c["CurrentCycle"] + 1
        set_option!(dev, "keep_rho", output_due(c, t, c["CurrentCycle"]) || output_due(c, t1, cyc1) ||
                                     !(t1 < c["StopTime"] && cyc1 < c["StopCycle"]))
        st = step_comoving!(dev, dt, a1, ak, adk, a2)                                # :854-868
        profile_step!(c, st)
        analyzePM(gp, gd, p, c, dev)
        c["CurrentTime"] += dt
        c["CurrentRedshift"] = z_from_t_internal(cosmo, c["CurrentTime"], zinit)
        println("max pa:", st.max_pa, " max v:", st.max_v, " dt:", dt, " z=", c["CurrentRedshift"])  # :875-876
        dt = c["CourantFactor"] * dx / (st.max_abs_v + 1e-30)                        # :878
        dt = min(dt, c["DtFractionOfTime"] * c["CurrentTime"], c["MaximumDt"])
        if c["CurrentTime"] + dt > c["StopTime"]
            dt = (1.0 + 1e-6) * (c["StopTime"] - c["CurrentTime"])                   # :882-885
            println("final dt = ", dt)
        end
        c["CurrentCycle"] += 1
        analyzePM(gp, gd, p, c, dev)
    end
    analyzePM(gp, gd, p, c, dev, force = true)                                       # :890
    sync_host!(dev, gd, p)
    nothing
end

# ---- driver (src/NoName.jl:19-100, src/restart.jl, src/IOtools.jl:118-138) ---------------------
write_all(fname, conf, gp, gd, p) = open(io -> serialize(io, Dict("conf" => Dict(k => v for (k, v) in conf if !startswith(k, "_")),
                                                                  "grids" => gp, "gridData" => gd, "particles" => p)), fname, "w")
function restart(gp, gd, p, conf)                                                    # src/restart.jl
    blob = open(deserialize, conf["_restart_file"])
    merge!(conf, blob["conf"]); merge!(gp, blob["grids"]); merge!(gd, blob["gridData"]); merge!(p, blob["particles"])
    isfile(get(conf, "_rconf", "")) && merge!(conf, parseconf(conf["_rconf"]))
    conf["StartTime"] = conf["CurrentTime"]; conf["StartCycle"] = conf["CurrentCycle"]
    haskey(conf, "Cosmology") && (conf["_cosmo"] = parse_cosmology(conf["Cosmology"]))
    conf["_rank"] = size(p["x"], 1)
    nothing
end
const INITIALIZERS = Dict("InitializeParticleSimulation" => InitializeParticleSimulation, "restart" => restart)
const EVOLVERS = Dict("evolvePM" => evolvePM, "evolvePMCosmology" => evolvePMCosmology)

conf = Dict{String,Any}()
function noname(conf)                                                                # NoName.jl:19-39
    gp = Dict(); gd = Dict(); p = Dict{String,Any}()
    INITIALIZERS[conf["Initializer"]](gp, gd, p, conf)
    EVOLVERS[conf["Evolve"]](gp, gd, p, conf)
    write_all(conf["RestartFileName"], conf, gp, gd, p)
    (get(conf, "ProfilingOn", false) === true && haskey(conf, "_profile")) && write_profiling_results(conf["ProfilingResultsFile"], conf["_profile"])
    gp, gd, p
end
"gp, gd, p = run_noname(\"particle_mesh.conf\")   (NoName.jl:51-100; restart = true makes the file a restart dump)"
function run_noname(configuration_file = ARGS[end]; restart = false, rconf = "", overrides = Dict{String,Any}())
    global conf
    if restart
        conf = load_conf(nothing; overrides = overrides)
        conf["Initializer"] = "restart"; conf["_restart_file"] = configuration_file; conf["_rconf"] = rconf
    else
        conf = load_conf(configuration_file; overrides = overrides)
    end
    return noname(conf)
end

end # module
