# bench_reference.jl -- times the REAL SGFHE.jl bootstrap on a machine that has Julia + DarkIntegers 0.1.x
# (neither exists in the build image, so this script ships unexecuted).  Same measurement as the reference's own
# test/performance.test.jl:114-139, extended to a chosen n and to all Julia threads; prints gates/s so the number can
# be set beside bench.py's `cpu_baseline` (a C port of the same algorithm).
#
#   julia --project=/path/to/SGFHE.jl -t auto julia/bench_reference.jl 1024 8
#
using Random, BenchmarkTools, SGFHE

n = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 64
gates = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : Threads.nthreads()

rng = MersenneTwister(1)
params = Params(n)
key = PrivateKey(params, rng)
bkey = BootstrapKey(rng, key)
message = rand(rng, Bool, params.n)
enc_bits = split_ciphertext(encrypt(key, rng, message))

# single call, as test/performance.test.jl:137 does
trial = @benchmark bootstrap($bkey, nothing, $(enc_bits[1]), $(enc_bits[2])) samples=3 evals=1
t1 = minimum(trial.times) / 1e9
println("Params($n): one bootstrap, 1 thread: $(round(t1, digits=3)) s  -> $(round(1 / t1, digits=4)) gates/s")

# independent gates over all threads (the reference itself is single threaded; gates are independent)
t = @elapsed Threads.@threads for g in 1:gates
    a, o, x = bootstrap(bkey, nothing, enc_bits[2g - 1], enc_bits[2g])
    @assert decrypt(key, a) == (message[2g - 1] & message[2g])
end
println("Params($n): $gates gates over $(Threads.nthreads()) threads: $(round(gates / t, digits=4)) gates/s")
