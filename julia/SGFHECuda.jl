# SGFHECuda.jl -- reference-side binding for libsgfhe_cuda.so (UNEXECUTED: Julia is not installed in the build
# image; the same C ABI is exercised from Python in tests/).  Drop next to src/fhe.jl and `include` it from
# src/SGFHE.jl after fhe.jl; export `CudaBootstrapKey`.
#
# The public API keeps its names AND signatures.  The backend is selected by the type of the key:
#
#     bkey  = BootstrapKey(rng, sk)                  # reference, CPU                         src/fhe.jl:181
#     cbkey = CudaBootstrapKey(bkey)                 # upload an existing key, or
#     cbkey = CudaBootstrapKey(rng, sk)              # generate it on the device (same draws as src/fhe.jl:192-194)
#     bootstrap(cbkey, rng, enc_bit1, enc_bit2)      # same call as src/fhe.jl:608-610, runs on the GPU
#     pack_encrypted_bits(cbkey, rng, enc_bits)      # same call as src/fhe.jl:660-662
#
# Values cross the boundary as canonical residues (`value(x)`), never as raw MgModUInt words.  Every random number is
# drawn HERE, from the caller's rng, with the reference's range types and in the reference's order, and handed to the
# library: Julia's `rand(rng, range)` stream depends on the element type of the range, so the draws use exactly
# `-xmax_i:xmax_i` with `xmax_i :: signed(encompassing_type(T))` (Int128 for every n >= 128, src/utils.jl:204,216) and
# are narrowed to Int64 afterwards (|x| <= 3B/2 < 2^45).

const libsgfhe = "libsgfhe_cuda"

function _check(rc::Cint)
    rc == 0 || error(unsafe_string(ccall((:sgfhe_last_error, libsgfhe), Cstring, ())))
    nothing
end

mutable struct CudaContext
    handle :: Ptr{Cvoid}
    params :: Params
end

function CudaContext(params::Params; device::Integer=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    _check(ccall((:sgfhe_ctx_create, libsgfhe), Cint, (Int32, Int32, Ref{Ptr{Cvoid}}), params.n, device, h))
    ctx = CudaContext(h[], params)
    finalizer(c -> ccall((:sgfhe_ctx_destroy, libsgfhe), Cint, (Ptr{Cvoid},), c.handle), ctx)
    ctx
end

# one context per (n, device); a context holds ONE transformed key at a time (see _ensure_key!)
const _contexts = Dict{Tuple{Int, Int}, CudaContext}()
_context(params::Params, device::Integer) = get!(() -> CudaContext(params; device=device), _contexts, (params.n, Int(device)))

function _key_token(ctx::CudaContext)
    tok = Ref{UInt64}(0)
    _check(ccall((:sgfhe_bkey_token, libsgfhe), Cint, (Ptr{Cvoid}, Ref{UInt64}, Ptr{Int32}), ctx.handle, tok, C_NULL))
    tok[]
end

"""
A bootstrap key whose transformed copy lives on the GPU.  `bkey` is the reference key it was made from
(`nothing` when it was generated on the device without keeping the coefficient form).
"""
mutable struct CudaBootstrapKey
    params :: Params
    ctx :: CudaContext
    bkey :: Union{BootstrapKey, Nothing}
    token :: UInt64
end

_wide(x) = (UInt64(x & typemax(UInt64)), UInt64(x >> 64))

# BootstrapKey.key (src/fhe.jl:176-201) -> C order [n][4][2][m][2] UInt64
function _upload!(k::CudaBootstrapKey)
    p = k.params
    buf = Array{UInt64}(undef, 2, p.m, 2, 4, p.n)          # Julia is column-major: reversed C order
    for i in 1:p.n, j in 1:4, c in 1:2, kk in 1:p.m
        lo, hi = _wide(convert(BigInt, value(k.bkey.key[i][j, c].coeffs[kk])))
        buf[1, kk, c, j, i] = lo
        buf[2, kk, c, j, i] = hi
    end
    _check(ccall((:sgfhe_bkey_upload, libsgfhe), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Int32), k.ctx.handle, buf, p.n))
    k.token = _key_token(k.ctx)
    nothing
end

"""    CudaBootstrapKey(bkey::BootstrapKey; device=0) -- upload a reference key (transformed once on the device)."""
function CudaBootstrapKey(bkey::BootstrapKey; device::Integer=0)
    k = CudaBootstrapKey(bkey.params, _context(bkey.params, device), bkey, UInt64(0))
    _upload!(k)
    k
end

"""
    CudaBootstrapKey(rng, sk::PrivateKey; device=0)

`BootstrapKey(rng, sk)` (src/fhe.jl:181-201) generated on the device.  The draws are the reference's own, in its order
(per row i: a_1..a_4 from `range_Q`, then e_1..e_4 from `-n:n`, src/fhe.jl:192-194); the products, `+ e_j`, `+ s_i G` and
the transform run in `sgfhe_bkey_generate`.
"""
function CudaBootstrapKey(rng::AbstractRNG, sk::PrivateKey; device::Integer=0, rows_per_call::Int=64)
    p = sk.params
    ctx = _context(p, device)
    tp = encompassing_type(type_Q(p))
    range_Q = zero(tp):convert(tp, p.Q)-one(tp)                                     # src/fhe.jl:187-188
    skb = UInt8[value(x) for x in sk.key.coeffs]
    for row0 in 0:rows_per_call:p.n-1
        rows = min(rows_per_call, p.n - row0)
        a = Array{UInt64}(undef, 2, p.m, 4, rows)                                   # C [rows][4][m][2]
        e = Array{Int64}(undef, p.m, 4, rows)                                       # C [rows][4][m]
        for i in 1:rows
            for j in 1:4                                                            # src/fhe.jl:193
                aj = rand(rng, range_Q, p.m)
                for kk in 1:p.m
                    a[1, kk, j, i], a[2, kk, j, i] = _wide(aj[kk])
                end
            end
            for j in 1:4                                                            # src/fhe.jl:194
                e[:, j, i] = rand(rng, -p.n:p.n, p.m)
            end
        end
        _check(ccall((:sgfhe_bkey_generate, libsgfhe), Cint,
            (Ptr{Cvoid}, Ptr{UInt8}, Ptr{UInt64}, Ptr{Int64}, Int32, Int32, Ptr{UInt64}),
            ctx.handle, skb, a, e, row0, rows, C_NULL))
    end
    CudaBootstrapKey(p, ctx, nothing, _key_token(ctx))
end

# bootstrap is a pure function of its key in the reference; a context holds one key, so check whose it is
function _ensure_key!(k::CudaBootstrapKey)
    _key_token(k.ctx) == k.token && return nothing
    k.bkey === nothing && error("the device copy of this key was replaced by another key and no reference key is attached")
    _upload!(k)
end

_flat(lwe::LWE) = UInt64[value.(lwe.a); value(lwe.b)]

# rand(rng, -xmax_i:xmax_i) exactly as flatten(rng, ...) calls it (src/utils.jl:204-216, 229)
function _xmax_i(p::Params)
    etp = encompassing_type(type_Q(p))
    B_u = convert(etp, p.B)
    xmax = isodd(B_u) ? (B_u - 1) ÷ 2 * 3 : B_u ÷ 2 * 3
    signed(convert(etp, xmax))
end

# Draws of `steps` external products per gate, in the reference's order: step k, polynomial a then b, coefficient,
# digit (src/fhe.jl:524-525, src/utils.jl:228-230, 257-258).  Column-major (2, m, 2, steps, gates) = C [gates][steps][2][m][2].
function _draws(rng::AbstractRNG, p::Params, steps::Int, gates::Int)
    x = _xmax_i(p)
    out = Array{Int64}(undef, 2, p.m, 2, steps, gates)
    for g in 1:gates, k in 1:steps, ab in 1:2, j in 1:p.m, d in 1:2
        out[d, j, ab, k, g] = Int64(rand(rng, -x:x))
    end
    out
end
# Draws of one shortened external product per key row (src/fhe.jl:636, 683-684): C [n][m][2]
function _draws_short(rng::AbstractRNG, p::Params)
    x = _xmax_i(p)
    out = Array{Int64}(undef, 2, p.m, p.n)
    for i in 1:p.n, j in 1:p.m, d in 1:2
        out[d, j, i] = Int64(rand(rng, -x:x))
    end
    out
end

_wrap_r(p::Params, o, g) = (tp = type_r(p);
    EncryptedBit(LWE([tp(o[kk, g], DarkIntegers._verbatim) for kk in 1:p.n], tp(o[p.n + 1, g], DarkIntegers._verbatim))))

"""
    bootstrap(bkey::CudaBootstrapKey, rng|nothing, bits1::Vector{EncryptedBit}, bits2::Vector{EncryptedBit})

Batched form of `bootstrap` (src/fhe.jl:608-621) on the GPU: gate g is `bootstrap(bkey, rng, bits1[g], bits2[g])`, with
the rng consumed gate after gate exactly as a loop over the reference call would.  Returns three vectors.
"""
function bootstrap(bkey::CudaBootstrapKey, rng::Union{AbstractRNG, Nothing},
        bits1::AbstractVector{EncryptedBit}, bits2::AbstractVector{EncryptedBit})
    p = bkey.params
    _ensure_key!(bkey)
    batch = length(bits1)
    @assert length(bits2) == batch
    l1 = hcat((_flat(b.lwe) for b in bits1)...)            # (n+1) x batch, column-major = C [batch][n+1]
    l2 = hcat((_flat(b.lwe) for b in bits2)...)
    outs = [Array{UInt64}(undef, p.n + 1, batch) for _ in 1:3]
    draws = rng === nothing ? C_NULL : _draws(rng, p, p.n, batch)    # an Array argument is rooted for the whole ccall
    _check(ccall((:sgfhe_bootstrap_batch, libsgfhe), Cint,
        (Ptr{Cvoid}, Int32, Ptr{UInt64}, Ptr{UInt64}, Ptr{Int64}, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}),
        bkey.ctx.handle, batch, l1, l2, draws, outs[1], outs[2], outs[3]))
    ([_wrap_r(p, outs[1], g) for g in 1:batch], [_wrap_r(p, outs[2], g) for g in 1:batch],
     [_wrap_r(p, outs[3], g) for g in 1:batch])
end

# The reference's signature (src/fhe.jl:608-610), dispatched on the key type.
function bootstrap(bkey::CudaBootstrapKey, rng::Union{AbstractRNG, Nothing},
        enc_bit1::EncryptedBit, enc_bit2::EncryptedBit)
    a, o, x = bootstrap(bkey, rng, [enc_bit1], [enc_bit2])
    a[1], o[1], x[1]
end

"""
    pack_encrypted_bits(bkey::CudaBootstrapKey, rng|nothing, enc_bits)

The reference's signature (src/fhe.jl:660-662); all of src/fhe.jl:670-693 runs in `sgfhe_pack_encrypted_bits`.  With an
rng the draws are made in the reference's order: the n internal bootstraps (src/fhe.jl:673), then the n shortened
external products (src/fhe.jl:683-684).
"""
function pack_encrypted_bits(bkey::CudaBootstrapKey, rng::Union{AbstractRNG, Nothing},
        enc_bits::AbstractArray{EncryptedBit, 1})
    p = bkey.params
    @assert length(enc_bits) == p.n                                                  # src/fhe.jl:667
    _ensure_key!(bkey)
    bits = hcat((_flat(b.lwe) for b in enc_bits)...)       # C [n][n+1]
    db = rng === nothing ? C_NULL : _draws(rng, p, p.n, p.n)
    ds = rng === nothing ? C_NULL : _draws_short(rng, p)
    w = Vector{UInt64}(undef, p.m)
    v = Vector{UInt64}(undef, p.m)
    _check(ccall((:sgfhe_pack_encrypted_bits, libsgfhe), Cint,
        (Ptr{Cvoid}, Ptr{UInt64}, Ptr{Int64}, Ptr{Int64}, Ptr{UInt64}, Ptr{UInt64}),
        bkey.ctx.handle, bits, db, ds, w, v))
    tp = type_r(p)
    poly(x) = Polynomial([tp(c, DarkIntegers._verbatim) for c in x], negacyclic_modulus)
    Ciphertext(p, RLWE(poly(w), poly(v)))                                            # src/fhe.jl:695
end

"""
`split_ciphertext` (src/fhe.jl:287-290) on the GPU: the n `EncryptedBit`s of one RLWE ciphertext.
"""
function split_ciphertext(bkey::CudaBootstrapKey, ct::Union{Ciphertext, PackedCiphertext})
    p = bkey.params
    N = length(ct.rlwe.a.coeffs)
    a = UInt64[value(x) for x in ct.rlwe.a.coeffs]
    b = UInt64[value(x) for x in ct.rlwe.b.coeffs]
    lwes = Array{UInt64}(undef, p.n + 1, p.n)              # column-major = C [n][n+1]
    _check(ccall((:sgfhe_split_ciphertext, libsgfhe), Cint,
        (Ptr{Cvoid}, Int32, Int32, Ptr{UInt64}, Ptr{UInt64}, Ptr{UInt64}), bkey.ctx.handle, 1, N, a, b, lwes))
    [_wrap_r(p, lwes, i) for i in 1:p.n]
end

"""
`decrypt(key, ::EncryptedBit)` (src/fhe.jl:504-507) for a vector of encrypted bits on the GPU.
"""
function decrypt(bkey::CudaBootstrapKey, key::PrivateKey, bits::AbstractVector{EncryptedBit})
    l = hcat((_flat(b.lwe) for b in bits)...)
    sk = UInt8[value(x) for x in key.key.coeffs]
    out = Vector{UInt8}(undef, length(bits))
    _check(ccall((:sgfhe_decrypt_bits, libsgfhe), Cint,
        (Ptr{Cvoid}, Int32, Ptr{UInt64}, Ptr{UInt8}, Ptr{UInt8}), bkey.ctx.handle, length(bits), l, sk, out))
    convert.(Bool, out)                                    # InexactError for values > 1, as at src/fhe.jl:506
end
