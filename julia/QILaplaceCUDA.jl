# QILaplaceCUDA.jl -- the reference-side binding of libqilcuda.so.  NON-FUNCTIONAL UNTIL RUN UNDER JULIA: Julia is not
# installed in the build container, so this file has never been executed (round-1 review found four symbol / arity
# mistakes in it, fixed in round 2 by reading the reference; there may be more).  The identical C ABI is exercised by the
# Python ctypes harness.
#
# Drop this file into QILaplace.jl's `src/` and `include` it after `mps.jl`/`mpo.jl`.  It overrides the BODIES of
# the hot-path functions; signatures, return types and exception types stay those of the reference.  ITensor
# `Index` bookkeeping stays here; only flat buffers cross the boundary (C-order [l][s][r] == Array(T, r, s, l)).
module QILaplaceCUDA

using ITensors, Random, Printf
import LinearAlgebra
using ..Mps: SignalMPS, ZTMPS, PairCore, _as_signal_2n, _writeback_signal_2n
using ..Mpo: SingleSiteMPO, PairedSiteMPO
using ..ApplyMPO: _as_single_site_mpo, _paired_from_single      # src/linalg/apply.jl:16-58

const LIB = get(ENV, "QILCUDA_LIB", "libqilcuda.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

struct QilError <: Exception
    code::Cint
    msg::String
end

function _throw(code::Cint)
    msg = unsafe_string(ccall((:qil_last_error, LIB), Cstring, ()))
    code == 1 && throw(ArgumentError(msg))
    code == 2 && throw(DomainError(msg))
    code == 3 && throw(ErrorException(msg))
    code == 4 && throw(AssertionError(msg))
    throw(QilError(code, msg))
end
_check(code::Cint) = code == 0 ? nothing : _throw(code)

function ctx()
    if CTX[] == C_NULL
        out = Ref{Ptr{Cvoid}}(C_NULL)
        _check(ccall((:qil_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), 0, out))
        CTX[] = out[]
    end
    return CTX[]
end

_iscomplex(::Type{<:Complex}) = Cint(1)
_iscomplex(::Type{<:Real}) = Cint(0)
_maxdim(m) = m >= typemax(Int) ÷ 2 ? Int64(0) : Int64(m)

# ---- handle <-> ITensor conversion -------------------------------------------------------------------
# C-order [l][s][r] is Julia's column-major Array(T, r, s, l)
function _upload(ψ::SignalMPS)
    n = length(ψ.data)
    T = promote_type(map(eltype, ψ.data)...)
    bufs = Vector{Array{T}}(undef, n)
    bond = ones(Int64, n + 1)
    for i in 1:n
        l = i == 1 ? nothing : ψ.bonds[i-1]
        r = i == n ? nothing : ψ.bonds[i]
        inds_rsl = filter(!isnothing, (r, ψ.sites[i], l))
        bufs[i] = Array{T}(Array(ψ.data[i], inds_rsl...))
        bond[i+1] = i == n ? 1 : dim(ψ.bonds[i])
    end
    ptrs = [pointer(b) for b in bufs]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve bufs begin
        _check(ccall((:qil_mps_from_host, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Cdouble, Ref{Ptr{Cvoid}}),
                     ctx(), n, _iscomplex(T), bond, ptrs, ψ.amplitude, out))
    end
    return out[]
end

function _download_mps(h::Ptr{Cvoid}, sites::Vector{<:Index}; bondtag="bond-%d")
    n = Ref{Cint}(0); ic = Ref{Cint}(0); amp = Ref{Cdouble}(0)
    _check(ccall((:qil_mps_info, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}, Ref{Cdouble}), h, n, ic, amp))
    N = Int(n[]); T = ic[] == 1 ? ComplexF64 : Float64
    bond = Vector{Int64}(undef, N + 1)
    _check(ccall((:qil_mps_dims, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h, bond))
    bonds = [Index(Int(bond[i+1]); tags=Printf.format(Printf.Format(bondtag), i)) for i in 1:(N-1)]
    data = Vector{ITensor}(undef, N)
    for i in 1:N
        buf = Array{T}(undef, Int(bond[i+1]), 2, Int(bond[i]))          # (r, s, l)
        _check(ccall((:qil_mps_get_core, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), h, i - 1, buf))
        is = Any[]
        i < N && push!(is, bonds[i]); push!(is, sites[i]); i > 1 && push!(is, bonds[i-1])
        data[i] = ITensor(reshape(buf, (dim(x) for x in is)...), is...)
    end
    ccall((:qil_mps_free, LIB), Cint, (Ptr{Cvoid},), h)
    return SignalMPS(data, sites, bonds; amplitude=amp[])
end

# ---- signal_mps (src/signals/SignalConverters.jl:228-233) ------------------------------------------
function signal_mps(x::AbstractVector{<:Number}; method::Symbol=:svd, cutoff::Real=1e-15,
                    maxdim::Int=typemax(Int), k::Int=20, p::Int=10, q::Int=0, random_seed::Int=1234,
                    mindim::Int=1, kwargs...)
    method ∈ (:svd, :rsvd) || throw(ArgumentError("tensor_to_mps: unknown method $method. Use :svd or :rsvd."))
    T = eltype(x) <: Complex ? ComplexF64 : Float64
    xv = Vector{T}(x)
    n = round(Int, log2(length(xv)))
    sites = [Index(2; tags=@sprintf("site-%d", i)) for i in 1:n]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    if method == :svd
        _check(ccall((:qil_encode_svd, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Cdouble, Int64, Ref{Ptr{Cvoid}}),
                     ctx(), _iscomplex(T), xv, length(xv), cutoff, _maxdim(maxdim), out))
    else
        # same side effect as the reference: the global RNG is reseeded (rsvd.jl:74).  The normal stream is MEANT to be the
        # one `random_itensor(eltype, cR, alpha)` would consume (column-major fill, Omega[c, j] = stream[c + C*(j-1)]); that
        # this reproduces the reference's Omega bit for bit is a CLAIM nobody has run (no Julia in the build image)
        Random.seed!(random_seed)
        cols_top = 2^(n - n ÷ 2)
        stream = randn(T, cols_top * min(k + p, cols_top))
        _check(ccall((:qil_encode_rsvd, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Cint, Cint, Cint, Int64, Cdouble, Int64, Int64,
                      Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}),
                     ctx(), _iscomplex(T), xv, length(xv), k, p, q, random_seed, cutoff, _maxdim(maxdim), mindim,
                     stream, length(stream), 0 #= flags: 1 = QIL_RSVD_ADAPTIVE (opt-in) =#, out))
    end
    return _download_mps(out[], sites)
end

# ---- coefficient (src/mps.jl:669-693): batched form used by the tutorials' loops ------------------
function coefficients(ψ::SignalMPS, bits::AbstractMatrix{<:Integer})     # bits is B x n
    size(bits, 2) == length(ψ.data) ||
        throw(ArgumentError("coefficient: expected $(length(ψ.data)) entries, got $(size(bits, 2))"))
    h = _upload(ψ)
    b = Matrix{UInt8}(permutedims(bits))                                 # n x B column-major == [B][n] C-order
    T = any(t -> eltype(t) <: Complex, ψ.data) ? ComplexF64 : Float64
    out = Vector{T}(undef, size(bits, 1))
    rc = ccall((:qil_coefficient_batch, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Int64, Ptr{Cvoid}),
               ctx(), h, b, size(bits, 1), out)
    ccall((:qil_mps_free, LIB), Cint, (Ptr{Cvoid},), h)
    _check(rc)
    return out
end
coefficient(ψ::SignalMPS, config::AbstractVector{<:Integer}) = coefficients(ψ, reshape(collect(config), 1, :))[1]
coefficient(ψ::ZTMPS, config) = coefficient(_as_signal_2n(ψ), config)

# ---- dense coefficient grids (pole scans docs/src/tutorials/zt.jl:152-157, 283-411; mps_to_vector mps.jl:716-743) --
# site_mode[i] in (0, 1) fixes site i, 2 frees it; out_bit[j] = output-index bit of the j-th free site (nothing = big-endian)
function coefficient_grid(ψ::SignalMPS, site_mode::AbstractVector{<:Integer}; out_bit=nothing)
    length(site_mode) == length(ψ.data) ||
        throw(ArgumentError("coefficient_grid: expected $(length(ψ.data)) site modes, got $(length(site_mode))"))
    h = _upload(ψ)
    mode = Vector{UInt8}(site_mode)
    F = count(==(2), mode)
    T = any(t -> eltype(t) <: Complex, ψ.data) ? ComplexF64 : Float64
    out = Vector{T}(undef, 2^F)
    ob = out_bit === nothing ? C_NULL : Vector{Int32}(out_bit)
    rc = GC.@preserve mode ob out ccall((:qil_coefficient_grid, LIB), Cint,
               (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}, Ptr{Int32}, Ptr{Cvoid}), ctx(), h, mode, ob, out)
    ccall((:qil_mps_free, LIB), Cint, (Ptr{Cvoid},), h)
    _check(rc)
    return out
end
mps_to_vector(ψ::SignalMPS; reverse::Bool=false) =
    coefficient_grid(ψ, fill(2, length(ψ.data)); out_bit=reverse ? collect(0:length(ψ.data)-1) : nothing)
mps_to_vector(ψ::ZTMPS; reverse::Bool=false) = mps_to_vector(_as_signal_2n(ψ); reverse=reverse)

# ---- one signal row-sharded over several GPUs (one Julia process per GPU, e.g. MPI.jl + NCCL.jl) ---------------
# struct qil_comm { Cint rank; Cint world; Ptr{Cvoid} user; allreduce_sum_f64; allgather_f64 }: build it with
#   @cfunction((user, buf, n) -> (NCCL.Allreduce!(unsafe_wrap(CuArray, Ptr{Float64}(buf), n), +, comm); Cint(0)), ...)
# and call qil_encode_rsvd_sharded_dev(ctx, Ref(comm), is_complex, d_x_local, N_total, k, p, q, seed, cutoff,
# maxdim, mindim, C_NULL, 0, flags, out) on every rank; all ranks receive the same MPS handle contents.

# ---- MPO handles ---------------------------------------------------------------------------------------
# C-order [l][p][s][r] (p = primed/input leg) is Julia's column-major Array(T, r, s, p, l)
function _upload(W::SingleSiteMPO)
    n = length(W.data)
    T = promote_type(map(eltype, W.data)...)
    bufs = Vector{Array{T}}(undef, n)
    bond = ones(Int64, n + 1)
    for i in 1:n
        l = i == 1 ? nothing : W.bonds[i-1]
        r = i == n ? nothing : W.bonds[i]
        inds = filter(!isnothing, (r, W.sites[i], prime(W.sites[i]), l))
        bufs[i] = Array{T}(Array(W.data[i], inds...))
        bond[i+1] = i == n ? 1 : dim(W.bonds[i])
    end
    ptrs = [pointer(b) for b in bufs]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve bufs begin
        _check(ccall((:qil_mpo_from_host, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}),
                     ctx(), n, _iscomplex(T), bond, ptrs, out))
    end
    return out[]
end

function _download_mpo(h::Ptr{Cvoid}, sites::Vector{<:Index}; bondtag="bond-%d")
    n = Ref{Cint}(0); ic = Ref{Cint}(0)
    _check(ccall((:qil_mpo_info, LIB), Cint, (Ptr{Cvoid}, Ref{Cint}, Ref{Cint}), h, n, ic))
    N = Int(n[]); T = ic[] == 1 ? ComplexF64 : Float64
    bond = Vector{Int64}(undef, N + 1)
    _check(ccall((:qil_mpo_dims, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h, bond))
    bonds = [Index(Int(bond[i+1]); tags=Printf.format(Printf.Format(bondtag), i)) for i in 1:(N-1)]
    data = Vector{ITensor}(undef, N)
    for i in 1:N
        buf = Array{T}(undef, Int(bond[i+1]), 2, 2, Int(bond[i]))       # (r, s, s', l)
        _check(ccall((:qil_mpo_get_core, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), h, i - 1, buf))
        is = Any[]
        i < N && push!(is, bonds[i]); push!(is, sites[i]); push!(is, prime(sites[i])); i > 1 && push!(is, bonds[i-1])
        data[i] = ITensor(reshape(buf, (dim(x) for x in is)...), is...)
    end
    ccall((:qil_mpo_free, LIB), Cint, (Ptr{Cvoid},), h)
    return SingleSiteMPO(data, sites, bonds)
end

# ---- apply (src/linalg/apply.jl:75-122, 201-218, 233-236): exact, kwargs ignored like the reference -------------
function apply(W::SingleSiteMPO, ψ::SignalMPS; kwargs...)
    length(W.data) == length(ψ.data) ||
        throw(ArgumentError("apply: MPO and MPS must have the same number of sites"))
    hW = _upload(W); hψ = _upload(ψ)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:qil_apply_mpo_mps, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), ctx(), hW, hψ, out)
    ccall((:qil_mpo_free, LIB), Cint, (Ptr{Cvoid},), hW)
    ccall((:qil_mps_free, LIB), Cint, (Ptr{Cvoid},), hψ)
    _check(rc)
    return _download_mps(out[], ψ.sites)
end
function apply(W::PairedSiteMPO, ψ::ZTMPS; kwargs...)                                 # apply.jl:201-218
    result = _writeback_signal_2n(apply(_as_single_site_mpo(W), _as_signal_2n(ψ)))   # 1-argument method, mps.jl:447
    result.amplitude = ψ.amplitude                                                   # apply.jl:214-217
    return result
end
Base.:*(W::Union{SingleSiteMPO,PairedSiteMPO}, ψ::Union{SignalMPS,ZTMPS}) = apply(W, ψ)

# ---- signal_ztmps (SignalConverters.jl:247-283): encode, then the copy-tensor split on the device --------------
function signal_ztmps(x::AbstractVector{<:Number}; cutoff::Real=1e-10, maxdim::Int=typemax(Int), kwargs...)
    ψ = signal_mps(x; cutoff=cutoff, maxdim=maxdim, kwargs...)
    h = _upload(ψ)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:qil_ztmps_split, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int64, Ref{Ptr{Cvoid}}),
               ctx(), h, cutoff, _maxdim(maxdim), out)
    ccall((:qil_mps_free, LIB), Cint, (Ptr{Cvoid},), h)
    _check(rc)
    n = length(ψ.sites)
    flat_sites = Index[]
    for i in 1:n
        push!(flat_sites, Index(2; tags=@sprintf("site-main-%d", i)))
        push!(flat_sites, Index(2; tags=@sprintf("site-copy-%d", i)))
    end
    result = _writeback_signal_2n(_download_mps(out[], flat_sites))          # 2n-site chain -> ZTMPS (mps.jl:447-472)
    result.amplitude = ψ.amplitude
    return result
end

# ---- canonicalize! / compress! / norm (src/mps.jl:754-999): in place on the handle, downloaded back ----------
function _inplace!(f::Function, ψ::SignalMPS)
    h = _upload(ψ)
    rc = f(h)
    rc == 0 || (ccall((:qil_mps_free, LIB), Cint, (Ptr{Cvoid},), h); _check(rc))
    new = _download_mps(h, ψ.sites)
    ψ.data .= new.data; ψ.bonds .= new.bonds; ψ.amplitude = new.amplitude
    return ψ
end
function canonicalize!(ψ::SignalMPS, dir::Symbol; center::Int=0, cutoff::Real=1e-12, maxdim::Int=typemax(Int))
    dir ∈ (:left, :right) || throw(ArgumentError("Direction must be :right or :left"))      # mps.jl:794
    _inplace!(h -> ccall((:qil_canonicalize, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cdouble, Int64),
                         ctx(), h, dir == :right ? 1 : 0, center, cutoff, _maxdim(maxdim)), ψ)
end
function compress!(ψ::SignalMPS; maxdim::Int=typemax(Int), tol::Real=1e-12, sweeps::Int=1)
    sweeps >= 1 || throw(DomainError(sweeps, "compress!: sweeps must be >= 1"))
    _inplace!(h -> ccall((:qil_compress, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cdouble, Cint),
                         ctx(), h, _maxdim(maxdim), tol, sweeps), ψ)
end
function LinearAlgebra.norm(ψ::SignalMPS)            # mps.jl:754-771 (ignores `amplitude`)
    h = _upload(ψ); v = Ref{Cdouble}(0)
    rc = ccall((:qil_norm, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cdouble}), ctx(), h, v)
    ccall((:qil_mps_free, LIB), Cint, (Ptr{Cvoid},), h)
    _check(rc)
    return v[]
end

# ---- transform MPOs (src/transforms/*.jl): built on the device, downloaded over the caller's site indices ------
function build_qft_mpo(n::Int, sites::Vector{<:Index}; cutoff::Real=1e-14, maxdim::Int=1000)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:qil_build_qft_mpo, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble, Int64, Ref{Ptr{Cvoid}}),
                 ctx(), n, cutoff, _maxdim(maxdim), out))
    return _download_mpo(out[], sites)
end
function _build_paired(sym::Symbol, n::Int, ωr::Real, sites_main, sites_copy; cutoff::Real=1e-14, maxdim::Int=1000)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = sym === :dt ?
        ccall((:qil_build_dt_mpo, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Int64, Ref{Ptr{Cvoid}}),
              ctx(), n, ωr, cutoff, _maxdim(maxdim), out) :
        ccall((:qil_build_zt_mpo, LIB), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Int64, Ref{Ptr{Cvoid}}),
              ctx(), n, ωr, cutoff, _maxdim(maxdim), out)
    _check(rc)
    flat = Index[]
    for i in 1:n
        push!(flat, sites_main[i]); push!(flat, sites_copy[i])
    end
    return _paired_from_single(_download_mpo(out[], flat))        # 2n-site chain -> PairedSiteMPO (apply.jl:34-58)
end
build_dt_mpo(n::Int, ωr::Real, sm, sc; kw...) = _build_paired(:dt, n, ωr, sm, sc; kw...)
build_zt_mpo(n::Int, ωr::Real, sm, sc; kw...) = _build_paired(:zt, n, ωr, sm, sc; kw...)

end # module
