# QILContainer.jl -- reader / writer of the QILTN001 container (include/qilcuda.h) for a Julia host.
# UNTESTED here (no Julia in the build image): the layout is pinned by tests/test_container.py on the Python side.
#
#   "QILTN001" | u32 kind (0 MPS, 1 MPO) | u32 is_complex | u32 n | u32 0 | f64 amplitude | i64 bond[n+1] | cores
# Cores are C-order [l][s][r] (MPS) or [l][p][s][r] (MPO), i.e. exactly Julia's column-major Array(T, r, s, l) /
# Array(T, r, s, p, l) -- the same arrays QILaplaceCUDA.jl hands to the C ABI.
module QILContainer

export read_container, write_container

function read_container(path::AbstractString)
    open(path, "r") do io
        String(read(io, 8)) == "QILTN001" || error("$path is not a QILTN001 container")
        kind, is_complex, n, _ = ntuple(_ -> ltoh(read(io, UInt32)), 4)
        amplitude = ltoh(read(io, Float64))
        bond = [ltoh(read(io, Int64)) for _ in 1:(n + 1)]
        T = is_complex == 1 ? ComplexF64 : Float64
        cores = Vector{Array{T}}(undef, n)
        for i in 1:n
            dims = kind == 0 ? (bond[i + 1], 2, bond[i]) : (bond[i + 1], 2, 2, bond[i])
            a = Array{T}(undef, dims...)
            read!(io, a)
            cores[i] = a
        end
        return (; kind = Int(kind), amplitude, bond, cores)
    end
end

function write_container(path::AbstractString, cores::Vector{<:Array}, amplitude::Real, kind::Integer)
    T = eltype(cores[1])
    n = length(cores)
    bond = [size(c, ndims(c)) for c in cores]
    push!(bond, size(cores[end], 1))
    open(path, "w") do io
        write(io, "QILTN001")
        for v in (UInt32(kind), UInt32(T <: Complex), UInt32(n), UInt32(0)); write(io, htol(v)); end
        write(io, htol(Float64(amplitude)))
        for b in bond; write(io, htol(Int64(b))); end
        for c in cores; write(io, c); end
    end
end

end # module
