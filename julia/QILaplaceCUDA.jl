# QILaplaceCUDA.jl -- the reference-side binding of libqilcuda.so (UNTESTED in this environment: Julia is not
# installed in the build container; the identical C ABI is exercised by the Python ctypes harness).
#
# Drop this file into QILaplace.jl's `src/` and `include` it after `mps.jl`/`mpo.jl`.  It overrides the BODIES of
# the hot-path functions; signatures, return types and exception types stay those of the reference.  ITensor
# `Index` bookkeeping stays here; only flat buffers cross the boundary (C-order [l][s][r] == Array(T, r, s, l)).
module QILaplaceCUDA

using ITensors, Random, Printf
using ..Mps: SignalMPS, ZTMPS, PairCore, _as_signal_2n, _writeback_signal_2n
using ..Mpo: SingleSiteMPO, PairedSiteMPO

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
        # same side effect as the reference: the global RNG is reseeded (rsvd.jl:74); the normal stream is the
        # one `random_itensor(eltype, cR, alpha)` would consume, so Omega matches the reference bit for bit
        Random.seed!(random_seed)
        cols_top = 2^(n - n ÷ 2)
        stream = randn(T, cols_top * min(k + p, cols_top))
        _check(ccall((:qil_encode_rsvd, LIB), Cint,
                     (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Cint, Cint, Cint, Int64, Cdouble, Int64, Int64,
                      Ptr{Cvoid}, Int64, Int64, Ref{Ptr{Cvoid}}),
                     ctx(), _iscomplex(T), xv, length(xv), k, p, q, random_seed, cutoff, _maxdim(maxdim), mindim,
                     stream, length(stream), 0, out))
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
# maxdim, mindim, C_NULL, 0, out) on every rank; all ranks receive the same MPS handle contents.

# ---- apply (src/linalg/apply.jl:75-122, 201-218) ---------------------------------------------------
# W is uploaded with qil_mpo_from_host exactly like _upload (cores Array(T, r, s, s', l)), then
#   qil_apply_mpo_mps(ctx, hW, hψ, out) ; _download_mps(out[], ψ.sites)
# compress!/canonicalize!/norm: _upload(ψ) ; qil_compress / qil_canonicalize / qil_norm ; download in place.
# build_qft_mpo / build_dt_mpo / build_zt_mpo: qil_build_*_mpo(ctx, n, [ωr,] cutoff, maxdim, out) ; download cores
# into ITensors over (bond_l, s', s, bond_r) with the reference's tags ("bond-%d").

end # module
