# SGFHECuda.jl -- reference-side binding for libsgfhe_cuda.so (UNEXECUTED: Julia is not installed in the build
# image; the same C ABI is exercised from Python in tests/).  Drop next to src/fhe.jl and `include` it from
# src/SGFHE.jl after fhe.jl.  Keeps every public name and signature of SGFHE.jl; only `bootstrap` changes
# backend, plus a batched method.
#
# Values cross the boundary as canonical residues (`value(x)`), never as raw MgModUInt words.

const libsgfhe = "libsgfhe_cuda"

mutable struct CudaContext
    handle :: Ptr{Cvoid}
    params :: Params
    key_uploaded :: Bool
end

function _check(rc::Cint)
    rc == 0 || error(unsafe_string(ccall((:sgfhe_last_error, libsgfhe), Cstring, ())))
    nothing
end

function CudaContext(params::Params; device::Integer=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:sgfhe_ctx_create, libsgfhe), Cint, (Int32, Int32, Ref{Ptr{Cvoid}}), params.n, device, h))
    ctx = CudaContext(h[], params, false)
    finalizer(c -> ccall((:sgfhe_ctx_destroy, libsgfhe), Cint, (Ptr{Cvoid},), c.handle), ctx)
    ctx
end

_wide(x) = (UInt64(x & typemax(UInt64)), UInt64(x >> 64))

# BootstrapKey.key (src/fhe.jl:176-201) -> C order [n][4][2][m][2] UInt64
function upload!(ctx::CudaContext, bkey::BootstrapKey)
    p = ctx.params
    buf = Array{UInt64}(undef, 2, p.m, 2, 4, p.n)          # Julia is column-major: reversed C order
    for i in 1:p.n, j in 1:4, c in 1:2, k in 1:p.m
        lo, hi = _wide(convert(BigInt, value(bkey.key[i][j, c].coeffs[k])))
        buf[1, k, c, j, i] = lo
        buf[2, k, c, j, i] = hi
    end
    _check(ccall((:sgfhe_bkey_upload, libsgfhe), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Int32), ctx.handle, buf, p.n))
    ctx.key_uploaded = true
    nothing
end

_flat(lwe::LWE) = UInt64[value.(lwe.a); value(lwe.b)]

# Draws in the reference's order: step k, polynomial a then b, coefficient, digit
# (src/fhe.jl:524-525, src/utils.jl:228-230, 257-258), each rand(rng, -xmax:xmax), xmax = 3 (B / 2).
function _draws(rng::AbstractRNG, p::Params, batch::Int)
    xmax = Int64(p.B ÷ 2 * 3)
    out = Array{Int64}(undef, 2, p.m, 2, p.n, batch)
    for g in 1:batch, k in 1:p.n, ab in 1:2, j in 1:p.m, d in 1:2
        out[d, j, ab, k, g] = rand(rng, -xmax:xmax)
    end
    out
end

"""
    bootstrap(ctx, bkey, rng|nothing, bits1::Vector{EncryptedBit}, bits2::Vector{EncryptedBit})

Batched form of `bootstrap` (src/fhe.jl:608-621) on the GPU: returns three vectors of `EncryptedBit`.
"""
function bootstrap(ctx::CudaContext, bkey::BootstrapKey, rng::Union{AbstractRNG, Nothing},
        bits1::AbstractVector{EncryptedBit}, bits2::AbstractVector{EncryptedBit})
    p = ctx.params
    ctx.key_uploaded || upload!(ctx, bkey)
    batch = length(bits1)
    @assert length(bits2) == batch
    l1 = hcat((_flat(b.lwe) for b in bits1)...)            # (n+1) x batch, column-major = C [batch][n+1]
    l2 = hcat((_flat(b.lwe) for b in bits2)...)
    outs = [Array{UInt64}(undef, p.n + 1, batch) for _ in 1:3]
    draws = rng === nothing ? C_NULL : pointer(_draws(rng, p, batch))
    _check(ccall((:sgfhe_bootstrap_batch, libsgfhe), Cint,
        (Ptr{Cvoid}, Int32, Ptr{UInt64}, Ptr{UInt64}, Ptr{Int64}, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}),
        ctx.handle, batch, l1, l2, draws, outs[1], outs[2], outs[3]))
    tp = type_r(p)
    wrap(o, g) = EncryptedBit(LWE(
        [tp(o[k, g], DarkIntegers._verbatim) for k in 1:p.n], tp(o[p.n + 1, g], DarkIntegers._verbatim)))
    ([wrap(outs[1], g) for g in 1:batch], [wrap(outs[2], g) for g in 1:batch], [wrap(outs[3], g) for g in 1:batch])
end

# Same signature as the reference's bootstrap, with a context in front.
function bootstrap(ctx::CudaContext, bkey::BootstrapKey, rng::Union{AbstractRNG, Nothing},
        enc_bit1::EncryptedBit, enc_bit2::EncryptedBit)
    a, o, x = bootstrap(ctx, bkey, rng, [enc_bit1], [enc_bit2])
    a[1], o[1], x[1]
end

"""
`split_ciphertext` (src/fhe.jl:287-290) on the GPU: the n `EncryptedBit`s of one RLWE ciphertext.
"""
function split_ciphertext(ctx::CudaContext, ct::Union{Ciphertext, PackedCiphertext})
    p = ctx.params
    N = length(ct.rlwe.a.coeffs)
    a = UInt64[value(x) for x in ct.rlwe.a.coeffs]
    b = UInt64[value(x) for x in ct.rlwe.b.coeffs]
    lwes = Array{UInt64}(undef, p.n + 1, p.n)              # column-major = C [n][n+1]
    _check(ccall((:sgfhe_split_ciphertext, libsgfhe), Cint,
        (Ptr{Cvoid}, Int32, Int32, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}), ctx.handle, 1, N, a, b, lwes))
    tp = type_r(p)
    [EncryptedBit(LWE([tp(lwes[k, i], DarkIntegers._verbatim) for k in 1:p.n], tp(lwes[p.n + 1, i], DarkIntegers._verbatim)))
     for i in 1:p.n]
end

"""
`decrypt(key, ::EncryptedBit)` (src/fhe.jl:504-507) for a vector of encrypted bits on the GPU.
"""
function decrypt(ctx::CudaContext, key::PrivateKey, bits::AbstractVector{EncryptedBit})
    p = ctx.params
    l = hcat((_flat(b.lwe) for b in bits)...)
    sk = UInt8[value(x) for x in key.key.coeffs]
    out = Vector{UInt8}(undef, length(bits))
    _check(ccall((:sgfhe_decrypt_bits, libsgfhe), Cint,
        (Ptr{Cvoid}, Int32, Ptr{UInt64}, Ptr{UInt8}, Ptr{UInt8}), ctx.handle, length(bits), l, sk, out))
    convert.(Bool, out)                                    # InexactError for values > 1, as at src/fhe.jl:506
end
