#!/usr/bin/env python3
"""bench.py -- bootstrapped gates/sec (AND/OR/XOR) for the SGFHE bootstrapping hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 1024] [--batch 4096] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic gate inputs: `batch` independent
bootstrap(bkey, nothing, y1, y2) calls (reference src/fhe.jl:608-621) at Params(n).  Default workload is
BASELINE.json configs[1]: paper-size Params(1024), 4096 gates per GPU.  Prints ONE JSON line (rank 0).

  value          gates/s with the LWE inputs already resident in HBM (sgfhe_bootstrap_batch_device)
  e2e            gates/s through the public host API (sgfhe_bootstrap_batch: pinned host buffers, H2D + D2H inside)
  roofline       integer-pipe roofline of the fused bootstrap kernel (algorithmic 32-bit
                 multiply-adds per gate / measured IMAD peak) plus the key-streaming HBM figure
  cpu_baseline   the CPU oracle (a C port of the reference's algorithm, literal formulation) on this box's cores
  --impl reference   times that CPU port as the reference arm (Julia + DarkIntegers cannot run in this image)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "bootstrapped gates/sec (AND/OR/XOR)"
UNIT = "gates/s"


def algorithmic_imad_per_gate(n: int, m: int, qbits: int) -> float:
    """W_mul = n m (3 log2 m + 8) modular multiplications in Z_Q, each a w-limb Montgomery
    product of 2 w^2 + w 32-bit multiply-adds (w = ceil(bits(Q) / 32))."""
    w = (qbits + 31) // 32
    return float(n) * m * (3 * (m.bit_length() - 1) + 8) * (2 * w * w + w)


def int_peak_imad_per_s() -> tuple[float, str]:
    """Measured IMAD.lo issue rate of this pool's B200 (tools/microbench/intpipe.cu -> profiles/int_peaks_r01.json);
    MEASURED_PEAKS.json has no integer figure."""
    path = os.path.join(ROOT, "profiles", "int_peaks_r01.json")
    try:
        with open(path) as f:
            return float(json.load(f)["imad_lo"]["thread_ops_per_s"]), "profiles/int_peaks_r01.json (measured, imad_lo)"
    except Exception:
        return 148 * 64 * 1.965e9, "nominal 148 SM x 64 lanes x 1.965 GHz (fallback)"


def ncu_dram_bytes_per_gate(n: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of one bootstrap_kernel launch (ncu --set full capture of a one-wave
    launch at Params(1024), committed under profiles/), per gate; None for other parameter sets."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_r02_traffic.json")) as f:
            t = json.load(f)
        return (t["dram_bytes_read"] + t["dram_bytes_write"]) / t["gates_per_launch"] if n == 1024 else None
    except Exception:
        return None


def hbm_peak_gbs() -> tuple[float, str]:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (a C port of the reference's algorithm; the reference itself is Julia + DarkIntegers,
# neither present in this image) -- bench.py may execute oracle/ only here and in cpu_baseline.
# ----------------------------------------------------------------------------------------------------------
class CpuArm:
    """The CPU arm: `threads` gates side by side, one per thread, in the oracle's literal formulation (24 transforms per
    step, src/fhe.jl:579-582).  Key material and inputs are built once; every sample() is one timed pass."""

    def __init__(self, n: int, threads: int, per_sample_s: float):
        import sgfhe_oracle as so
        so.build()
        self.so, self.n, self.threads = so, n, threads
        self.OP = OP = so.Params(n)
        so.set_setup_threads(max(threads, os.cpu_count() or 1))
        self.sk = so.make_secret(OP, 1)
        probe = min(OP.n, 4)
        key = so.make_bkey(OP, self.sk, 1, rows=probe)
        _, lwes = so.make_lwes(OP, self.sk, 1)
        self.l1 = np.ascontiguousarray(lwes[:threads]); self.l2 = np.ascontiguousarray(lwes[threads:2 * threads])
        t0 = time.perf_counter()
        so.bootstrap_batch(OP, key, self.l1, self.l2, n_steps=probe, literal=True, threads=threads)
        per_step = (time.perf_counter() - t0) / probe
        # all n steps (full gates) when that fits the budget, otherwise a prefix of the strictly sequential loop
        self.steps = OP.n if per_step * OP.n <= 1.5 * per_sample_s else int(max(probe, min(OP.n, per_sample_s / max(per_step, 1e-9))))
        self.key = key if self.steps <= probe else so.make_bkey(OP, self.sk, 1, rows=self.steps)

    def sample(self) -> dict:
        OP = self.OP
        t0 = time.perf_counter()
        self.so.bootstrap_batch(OP, self.key, self.l1, self.l2, n_steps=self.steps, literal=True, threads=self.threads)
        dt = time.perf_counter() - t0
        gates = self.threads * self.steps / OP.n                   # linear in the number of sequential steps
        how = "all %d accumulation steps (full gates)" % OP.n if self.steps == OP.n else \
            "first %d of %d accumulation steps, extrapolated linearly in steps" % (self.steps, OP.n)
        return {"value": gates / dt, "unit": UNIT, "cores": self.threads, "kind": "port", "seconds": dt,
                "sample": f"{self.threads} gate(s) x {how} at Params({self.n}), literal 24-NTT/step formulation (fhe.jl:579-582), "
                          f"{self.threads} thread(s); C port of SGFHE.jl's algorithm (Julia/DarkIntegers unavailable in this image)"}


def cpu_gates_per_s(n: int, target_s: float, threads: int | None = None, with_single: bool = True) -> dict:
    """`value` is the all-threads figure (independent gates side by side: the most the host can do); `single_thread` is the
    figure closest to the reference itself, which is single-threaded (no Threads / Distributed anywhere in src/)."""
    threads = threads or os.cpu_count() or 1
    out = CpuArm(n, threads, target_s).sample()
    out["host_cores"] = os.cpu_count()
    if with_single and threads > 1:
        one = CpuArm(n, 1, max(target_s / 3, 3.0)).sample()
        out["single_thread"] = {k: one[k] for k in ("value", "unit", "cores", "seconds", "sample")}
    return out


def run_reference(args) -> None:
    """--impl reference: K timed passes of the CPU arm.  A full gate is 1024 strictly sequential steps (about 27 s per gate
    and thread at Params(1024)), so with many steps requested each pass is a prefix of the loop sized to keep the whole run
    near three minutes; with few steps it is whole gates."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = args.cpu_threads or os.cpu_count() or 1
    per_step_s = min(30.0, max(3.0, 180.0 / max(args.steps, 1)))
    arm = CpuArm(args.n, threads, per_step_s)
    warm = CpuArm(args.n, threads, 1.0) if args.warmup else None
    for _ in range(args.warmup):
        warm.sample()
    vals, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = arm.sample()
        vals.append(last["value"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    cb = dict(last); cb["value"] = v; cb["host_cores"] = os.cpu_count()
    if threads > 1:
        one = CpuArm(args.n, 1, 5.0).sample()
        cb["single_thread"] = {k: one[k] for k in ("value", "unit", "cores", "seconds", "sample")}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u128 (CPU Montgomery)", "data": "synthetic",
        "config": {"workload": f"Params({args.n}) batch of {args.batch} random gate bootstraps (bounded CPU sample)",
                   "n": args.n, "batch_per_gpu": args.batch},
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------------------
def cuda_alias(ptr: int, nbytes: int, torch):
    class _A:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_A(), device="cuda")


def run_ours(args) -> None:
    import ctypes as C
    import torch
    import sgfhe_jl_b200 as sg
    from sgfhe_jl_b200 import _lib

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n, batch = args.n, args.batch
    if args.scaling == "strong":                         # --batch is the whole job: this rank's contiguous share of it
        from sgfhe_jl_b200.parallel import shard_bounds
        lo, hi = shard_bounds(args.batch, world)[rank]
        batch = hi - lo
    P = sg.Params(n, device=local)
    L = _lib.lib()
    # ---- key: generated and pre-transformed on rank 0, NCCL-broadcast in NTT form --------
    sk_rng = np.random.default_rng([args.seed, 1])
    sk = sg.PrivateKey(P, sk_rng)                       # same secret on every rank (same seed)
    t_key0 = time.perf_counter()
    dptr, nbytes = C.c_void_p(), C.c_uint64()
    _lib.check(L.sgfhe_bkey_device_buffer(P.ctx, n, C.byref(dptr), C.byref(nbytes)))
    if rank == 0:
        bkey = sg.BootstrapKey(np.random.default_rng([args.seed, 2]), sk)
        bkey.upload()
    if world > 1:
        kt = cuda_alias(dptr.value, nbytes.value, torch)
        dist.broadcast(kt, src=0)
        torch.cuda.synchronize()
    _lib.check(L.sgfhe_bkey_adopt(P.ctx, n))
    key_s = time.perf_counter() - t_key0

    # ---- inputs: batch disjoint pairs of valid encrypted bits (test/api.test.jl:61-67 style) --------------
    rng = np.random.default_rng([args.seed, 3, rank])
    blocks = (2 * batch + n - 1) // n
    bits, lw = [], []
    for _ in range(blocks):
        msg = rng.integers(0, 2, size=n, dtype=np.uint8)
        ct = sg.encrypt(sk, rng, msg)
        lw.append(np.stack([e.lwe.flat() for e in sg.split_ciphertext(ct)]))
        bits.append(msg)
    bits = np.concatenate(bits)[: 2 * batch]; lw = np.concatenate(lw)[: 2 * batch]
    h1 = torch.from_numpy(np.ascontiguousarray(lw[:batch])).pin_memory()
    h2 = torch.from_numpy(np.ascontiguousarray(lw[batch:])).pin_memory()
    houts = [torch.empty_like(h1).pin_memory() for _ in range(3)]
    d1, d2 = h1.cuda(), h2.cuda()
    douts = [torch.empty_like(d1) for _ in range(3)]
    stream = torch.cuda.current_stream()

    seed = int(args.rng_seed)                            # 0: rng = nothing; else bootstrap(bkey, rng, ...) with device-side draws
    gate0 = rank * batch                                 # disjoint Philox streams per rank

    def step_device():
        if seed:
            _lib.check(L.sgfhe_bootstrap_batch_rng_device(P.ctx, batch, d1.data_ptr(), d2.data_ptr(), seed, gate0,
                                                          *[o.data_ptr() for o in douts], stream.cuda_stream))
        else:
            _lib.check(L.sgfhe_bootstrap_batch_device(P.ctx, batch, d1.data_ptr(), d2.data_ptr(), None,
                                                      *[o.data_ptr() for o in douts], stream.cuda_stream))

    def step_host():
        if seed:
            _lib.check(L.sgfhe_bootstrap_batch_rng(P.ctx, batch, h1.data_ptr(), h2.data_ptr(), seed, gate0,
                                                   *[o.data_ptr() for o in houts]))
        else:
            _lib.check(L.sgfhe_bootstrap_batch(P.ctx, batch, h1.data_ptr(), h2.data_ptr(), None,
                                               *[o.data_ptr() for o in houts]))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_ms = ev0.elapsed_time(ev1)
        barrier()
        return dev_ms, wall

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = sg.launch_count()
    dev_ms, _ = timed(step_device, args.steps)
    launches = sg.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # e2e: host buffers through the public C ABI; wall clock brackets H2D + kernel + D2H (call is synchronous)
    step_host()
    sampler2 = ClockSampler(local)
    if rank == 0:
        sampler2.start()
    _, e2e_wall = timed(step_host, args.steps)
    clocks_e2e = sampler2.stop() if rank == 0 else None
    t = torch.tensor([dev_ms, e2e_wall * 1e3], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    # ---- verification (untimed): decrypt every output of this rank; device and host paths must agree -------
    o = [x.cpu().numpy() for x in douts]
    skb = sk.key.astype(bool)
    ok = all(np.array_equal(a, b.numpy()) for a, b in zip(o, houts))
    y1, y2 = bits[:batch].astype(np.int64), bits[batch:].astype(np.int64)
    for arr, want in zip(o, (y1 & y2, y1 | y2, y1 ^ y2)):
        b1 = (arr[:, n].astype(np.int64) - arr[:, :n][:, skb].astype(np.int64).sum(axis=1)) % P.r
        ok = ok and np.array_equal(((b1 + P.Dr // 2) % P.r) // P.Dr, want)
    okt = torch.tensor([1 if ok else 0], device="cuda")
    if dist is not None:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    verified = bool(int(okt[0]))

    if rank == 0:
        total_gates = (args.batch if args.scaling == "strong" else batch * world) * args.steps
        value = total_gates / (dev_ms / 1e3)
        e2e = total_gates / (e2e_ms / 1e3)
        qbits = P.Q.bit_length()
        imad_gate = algorithmic_imad_per_gate(n, P.m, qbits)
        peak, peak_src = int_peak_imad_per_s()
        per_gpu = value / world
        hbm, hbm_src = hbm_peak_gbs()
        pc = _lib.ParamsC(); _lib.check(L.sgfhe_params_get(P.ctx, C.byref(pc)))
        key_bytes = n * pc.rns_primes * 8 * P.m * 4
        waves = -(-batch // 148)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32 (RNS residues of Z_Q, exact)", "data": "synthetic",
            "config": {"workload": f"Params({n}) (m={P.m}, {qbits}-bit Q) batch of {batch} random gate bootstraps per GPU, " + ("rng=nothing" if not seed else f"rng=DeviceRng({seed}): randomised flatten, draws made on the device (Philox4x32-10)") +
                                   (f" (strong scaling: {args.batch} gates in total)" if args.scaling == "strong" else ""),
                       "n": n, "batch_per_gpu": batch, "parallelism": f"gates sharded over {world} GPU(s), key NCCL-broadcast once",
                       "l2": "working set (pre-transformed key %.2f GB + per-gate scratch) is larger than L2; no flush needed" % (key_bytes / 1e9)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(2 * h1.numel() * 8), "d2h_bytes_per_step": int(3 * h1.numel() * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "clocks_e2e": clocks_e2e,
            "verified": verified,
            "key_setup_s": key_s,
            "roofline": {"bound": "int32-pipe", "achieved": per_gpu * imad_gate / 1e9, "peak": peak / 1e9, "unit": "GIMAD/s",
                         "frac": per_gpu * imad_gate / peak,
                         "traffic": (ncu_dram_bytes_per_gate(n) * batch) if ncu_dram_bytes_per_gate(n) else None,
                         "traffic_note": "ncu DRAM bytes of a 148-gate launch (profiles/ncu_r02_traffic.json) scaled to this batch; "
                                         "almost all of it is per-gate scratch written back from L2, not algorithmic bytes",
                         "kernel": "bootstrap_kernel (one launch = one step = batch gates x n fused accumulation steps)",
                         "algorithmic_imad_per_gate": imad_gate, "peak_source": peak_src,
                         "hbm": {"algorithmic_key_bytes_per_launch": key_bytes * waves,
                                 "achieved_gbs": key_bytes * waves / (dev_ms / args.steps / 1e3) / 1e9, "peak_gbs": hbm, "peak_source": hbm_src,
                                 "note": "key streamed once per wave of 148 lock-step gates; far below HBM peak by design (integer bound)"}},
        }
        # second half of BASELINE.json's metric: standalone NTT polymuls/s at this ring degree (both operands full size)
        try:
            pb = int(os.environ.get("SGFHE_PM_BATCH", "2368"))       # 16 products per SM: no partial wave for one or two products per CTA
            pa = torch.from_numpy((np.random.default_rng(7).integers(0, 1 << 62, size=(pb, P.m, 2), dtype=np.uint64) &
                                   np.array([0xFFFFFFFFFFFFFFFF, (1 << max(qbits - 65, 0)) - 1], np.uint64)).view(np.int64)).cuda()
            pb2 = torch.from_numpy((np.random.default_rng(8).integers(0, 1 << 62, size=(pb, P.m, 2), dtype=np.uint64) &
                                    np.array([0xFFFFFFFFFFFFFFFF, (1 << max(qbits - 65, 0)) - 1], np.uint64)).view(np.int64)).cuda()
            po = torch.empty_like(pa)
            for _ in range(3):
                _lib.check(L.sgfhe_polymul_device(P.ctx, pb, pa.data_ptr(), pb2.data_ptr(), po.data_ptr(), stream.cuda_stream))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(5):
                _lib.check(L.sgfhe_polymul_device(P.ctx, pb, pa.data_ptr(), pb2.data_ptr(), po.data_ptr(), stream.cuda_stream))
            e1.record(stream); torch.cuda.synchronize()
            out["ntt_polymul"] = {"value": 5 * pb / (e0.elapsed_time(e1) / 1e3), "unit": "polymuls/s",
                                  "workload": f"negacyclic products a * b (a != b) in Z_Q[x]/(x^{P.m}+1), batch {pb}, both operands {qbits}-bit (sgfhe_polymul_device)"}
        except Exception as ex:  # the headline number must not depend on the secondary one
            out["ntt_polymul"] = {"error": str(ex)}
        if not args.no_cpu and world == 1:
            out["cpu_baseline"] = cpu_gates_per_s(n, args.cpu_seconds, args.cpu_threads)
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_depth(args) -> None:
    """BASELINE.json configs[4]: examples/depth.jl -- `layers` sequential gate layers at Params(512), `batch` independent
    bit pairs per layer and GPU, layer l+1 bootstrapping (AND_l, XOR_l); ciphertexts stay in HBM between layers.
    --wiring allgather: the layer outputs are all-gathered and rank r continues with the wires of rank r+1 (one NCCL
    collective per layer); --wiring local: every rank chains its own gates, nothing is exchanged."""
    import torch
    import sgfhe_jl_b200 as sg
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, W, layers = args.n, args.batch, args.layers
    P = sg.Params(n, device=local)
    sk = sg.PrivateKey(P, np.random.default_rng([args.seed, 1]))
    bkey = sg.BootstrapKey(np.random.default_rng([args.seed, 2]), sk) if rank == 0 else None   # generated on the device
    if world > 1:
        sg.broadcast_key(P, n, dist)
        bkey = sg.BootstrapKey.resident(P)

    def inputs(r):
        rng = np.random.default_rng([args.seed, 3, r])
        bits, lw = [], []
        for _ in range((2 * W + n - 1) // n):
            msg = rng.integers(0, 2, size=n, dtype=np.uint8)
            lw.append(np.stack([e.lwe.flat() for e in sg.split_ciphertext(sg.encrypt(sk, rng, msg))])); bits.append(msg)
        return np.concatenate(bits)[: 2 * W], np.concatenate(lw)[: 2 * W]

    bits, lw = inputs(rank)
    wiring = args.wiring if world > 1 else "local"
    for _ in range(args.warmup):
        sg.bootstrap_chain(bkey, lw[:W], lw[W:], 1)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        outs = sg.bootstrap_chain(bkey, lw[:W], lw[W:], layers, dist=dist, wiring=wiring)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # plaintext circuit: after every exchanged layer rank r holds the wires of rank r+1, i.e. after `layers` layers the
    # wires that started on rank (r + layers - 1) % world (the last layer's outputs are not exchanged)
    src = (rank + layers - 1) % world if wiring == "allgather" else rank
    sbits = bits if src == rank else inputs(src)[0]
    y1, y2 = sbits[:W].astype(np.int64), sbits[W:].astype(np.int64)
    for _ in range(layers):
        y1, y2 = y1 & y2, y1 ^ y2                                   # depth.jl:71-75
    skb = sk.key.astype(bool)
    b1 = (outs[0][:, n].astype(np.int64) - outs[0][:, :n][:, skb].astype(np.int64).sum(axis=1)) % P.r
    ok = bool(np.array_equal(((b1 + P.Dr // 2) % P.r) // P.Dr, y1))
    t = torch.tensor([dt, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        dt = float(t[0])
        print(json.dumps({"metric": METRIC, "value": W * layers * world * args.steps / dt, "unit": UNIT, "n_gpus": world,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "u32 (RNS residues of Z_Q, exact)", "data": "synthetic",
                          "config": {"workload": f"examples/depth.jl chain at Params({n}): {layers} sequential layers x {W} gates per GPU, "
                                                 "(AND, XOR) fed back in, ciphertexts resident in HBM", "n": n, "batch_per_gpu": W, "layers": layers,
                                     "wiring": wiring + (" (one all-gather of the layer outputs per layer; rank r continues with the wires of rank r+1)"
                                                         if wiring == "allgather" else " (no exchange between ranks)")},
                          "verified": bool(float(t[1]) == 0.0)}))
    if dist is not None:
        dist.destroy_process_group()


def run_pack(args) -> None:
    """pack_encrypted_bits (src/fhe.jl:660-696) through the public host entry point sgfhe_pack_encrypted_bits: one call = n
    internal bootstraps + n shortened external products + the sums and ModRed, all on the device.  `--steps` calls are timed
    (host buffers, copies inside the timed region); every packed ciphertext is decrypted (src/fhe.jl:471-494) and compared
    with the plaintext bits."""
    import torch
    import sgfhe_jl_b200 as sg
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    n = args.n
    P = sg.Params(n, device=local)
    rng = np.random.default_rng([args.seed, 1])
    sk = sg.PrivateKey(P, rng)
    bkey = sg.BootstrapKey(np.random.default_rng([args.seed, 2]), sk)
    msg = rng.integers(0, 2, size=n, dtype=np.uint8)
    ebits = sg.split_ciphertext(sg.encrypt(sk, rng, msg))
    for _ in range(max(args.warmup, 1)):
        ct = sg.pack_encrypted_bits(bkey, None, ebits)
    l0 = sg.launch_count()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ct = sg.pack_encrypted_bits(bkey, None, ebits)
    dt = time.perf_counter() - t0
    ok = bool(np.array_equal(sg.decrypt(sk, ct), msg.astype(bool)))
    print(json.dumps({"metric": "pack_encrypted_bits calls/sec", "value": args.steps / dt, "unit": "packs/s", "n_gpus": 1,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "u32 (RNS residues of Z_Q, exact)", "data": "synthetic",
                      "gpu_launches": int(sg.launch_count() - l0),
                      "gates_per_s_equivalent": n * args.steps / dt,
                      "config": {"workload": f"pack_encrypted_bits at Params({n}), rng=nothing: {n} internal bootstraps + {n} shortened "
                                             "external products + sums + ModRed per call, host buffers in and out", "n": n},
                      "verified": ok}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", "--params-n", dest="n", type=int, default=1024,
                    help="Params(n); under torchrun spell it --params-n (torchrun's own parser finds a bare --n ambiguous)")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rng-seed", type=int, default=0, help="gates workload: 0 = rng=nothing (the default, BASELINE's config); S > 0 = the randomised flatten of bootstrap(bkey, rng, ...) with draws made on the device from seed S")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=30.0, help="budget of the cpu_baseline sample: 30 s fits whole gates at Params(1024)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="threads of the CPU arm (0 = all host cores); a one-thread figure is reported next to it")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: --batch is the TOTAL over all GPUs")
    ap.add_argument("--wiring", default="local", choices=["local", "allgather"], help="depth workload: layer inputs from this rank only, or all-gathered and permuted across ranks")
    ap.add_argument("--workload", default="gates", choices=["gates", "depth", "pack"],
                    help="depth = BASELINE configs[4] (use --n 512 --batch W --layers L); pack = pack_encrypted_bits calls (src/fhe.jl:660-696)")
    ap.add_argument("--layers", type=int, default=100)
    args = ap.parse_args()
    args.cpu_threads = args.cpu_threads or None
    if args.workload == "depth" and args.impl == "ours":
        run_depth(args)
        return
    if args.workload == "pack" and args.impl == "ours":
        run_pack(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner when the
    # environment sets NCCL_DEBUG), so everything except that line is sent to stderr: fd 1 points at stderr while the
    # benchmark runs and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = _real_stdout
    try:
        main()
    finally:
        _real_stdout.flush()
