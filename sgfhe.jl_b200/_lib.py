"""Loader for libsgfhe_cuda.so (the C ABI declared in include/sgfhe_cuda.h).

The library is built in-tree by `build()` (nvcc, sm_100a only).  There is no CPU fallback: if the
library is missing or no CUDA device is present, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("SGFHE_CUDA_LIB") or os.path.join(_HERE, "libsgfhe_cuda.so")   # override: A/B builds only
_LIB = None

# every symbol include/sgfhe_cuda.h declares
SYMBOLS = [
    "sgfhe_ctx_create", "sgfhe_ctx_destroy", "sgfhe_params_get", "sgfhe_params_derive", "sgfhe_last_error",
    "sgfhe_bkey_upload", "sgfhe_bootstrap_batch", "sgfhe_bootstrap_batch_device", "sgfhe_bootstrap_trace",
    "sgfhe_polymul", "sgfhe_polymul_device", "sgfhe_flatten_poly", "sgfhe_external_product",
    "sgfhe_launch_count", "sgfhe_bkey_device_buffer", "sgfhe_bkey_adopt",
    "sgfhe_bootstrap_internal_batch", "sgfhe_shortened_products",
    "sgfhe_scheme2_params_derive", "sgfhe_rns2_op", "sgfhe_rns2_op_device",
    "sgfhe_bkey_export_size", "sgfhe_bkey_export", "sgfhe_bkey_import",
    "sgfhe_split_ciphertext", "sgfhe_split_ciphertext_device", "sgfhe_decrypt_bits", "sgfhe_decrypt_bits_device",
    "sgfhe_bkey_generate", "sgfhe_bkey_token", "sgfhe_pack_encrypted_bits", "sgfhe_pack_from_lwes", "sgfhe_bootstrap_batch_rng", "sgfhe_bootstrap_batch_rng_device", "sgfhe_device_draws",
    "sgfhe_s2_ctx_create", "sgfhe_s2_ctx_destroy", "sgfhe_s2_params_get", "sgfhe_s2_polymul", "sgfhe_s2_polymul_device",
    "sgfhe_s2_ntt_device", "sgfhe_s2_mac8_device", "sgfhe_s2_bkey_generate",
]


class ParamsC(C.Structure):
    _fields_ = [("n", C.c_int32), ("t", C.c_int32), ("m", C.c_int32), ("rns_primes", C.c_int32),
                ("r", C.c_uint64), ("q", C.c_uint64), ("Dr", C.c_uint64), ("Dq", C.c_uint64),
                ("Q", C.c_uint64 * 2), ("B", C.c_uint64 * 2), ("DQ_tilde", C.c_uint64 * 2)]


def build(force: bool = False) -> str:
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in ("sgfhe_cuda.cu", "device_math.cuh", "host_math.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "sgfhe_cuda.h"))
    stale = (not os.path.exists(SO_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", src_dir, "-s"] + (["-B"] if force else []))
    return SO_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(SO_PATH)
        vp, i32, u64p = C.c_void_p, C.c_int32, C.c_void_p
        L.sgfhe_last_error.restype = C.c_char_p
        L.sgfhe_launch_count.restype = C.c_uint64
        L.sgfhe_ctx_create.argtypes = [i32, i32, C.POINTER(vp)]
        L.sgfhe_ctx_destroy.argtypes = [vp]
        L.sgfhe_params_get.argtypes = [vp, C.POINTER(ParamsC)]
        L.sgfhe_params_derive.argtypes = [i32, C.POINTER(ParamsC)]
        L.sgfhe_bkey_upload.argtypes = [vp, u64p, i32]
        L.sgfhe_bootstrap_batch.argtypes = [vp, i32, u64p, u64p, vp, u64p, u64p, u64p]
        L.sgfhe_bootstrap_batch_device.argtypes = [vp, i32, u64p, u64p, vp, u64p, u64p, u64p, vp]
        L.sgfhe_bootstrap_trace.argtypes = [vp, u64p, u64p, vp, i32, u64p, u64p, u64p, u64p]
        L.sgfhe_polymul.argtypes = [vp, i32, u64p, u64p, u64p]
        L.sgfhe_polymul_device.argtypes = [vp, i32, u64p, u64p, u64p, vp]
        L.sgfhe_flatten_poly.argtypes = [vp, u64p, vp, u64p]
        L.sgfhe_external_product.argtypes = [vp, u64p, u64p, u64p, vp, u64p, u64p]
        L.sgfhe_bkey_device_buffer.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(C.c_uint64)]
        L.sgfhe_bkey_adopt.argtypes = [vp, i32]
        L.sgfhe_bootstrap_internal_batch.argtypes = [vp, i32, u64p, u64p, vp, u64p, u64p, u64p]
        L.sgfhe_shortened_products.argtypes = [vp, i32, u64p, vp, u64p]
        L.sgfhe_bkey_export_size.argtypes = [vp, i32, C.POINTER(C.c_uint64)]
        L.sgfhe_bkey_export.argtypes = [vp, i32, vp, C.c_uint64]
        L.sgfhe_bkey_import.argtypes = [vp, vp, C.c_uint64]
        L.sgfhe_split_ciphertext.argtypes = [vp, i32, i32, u64p, u64p, u64p]
        L.sgfhe_split_ciphertext_device.argtypes = [vp, i32, i32, u64p, u64p, u64p, vp]
        L.sgfhe_decrypt_bits.argtypes = [vp, i32, u64p, vp, vp]
        L.sgfhe_decrypt_bits_device.argtypes = [vp, i32, u64p, vp, vp, vp]
        L.sgfhe_bkey_generate.argtypes = [vp, vp, u64p, vp, i32, i32, u64p]
        L.sgfhe_bkey_token.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_int32)]
        L.sgfhe_pack_encrypted_bits.argtypes = [vp, u64p, vp, vp, u64p, u64p]
        L.sgfhe_pack_from_lwes.argtypes = [vp, u64p, vp, u64p, u64p]
        L.sgfhe_bootstrap_batch_rng.argtypes = [vp, i32, u64p, u64p, C.c_uint64, C.c_uint64, u64p, u64p, u64p]
        L.sgfhe_bootstrap_batch_rng_device.argtypes = [vp, i32, u64p, u64p, C.c_uint64, C.c_uint64, u64p, u64p, u64p, vp]
        L.sgfhe_device_draws.argtypes = [vp, C.c_uint64, C.c_uint64, i32, i32, vp]
        L.sgfhe_s2_ctx_create.argtypes = [i32, i32, C.POINTER(vp)]
        L.sgfhe_s2_ctx_destroy.argtypes = [vp]
        L.sgfhe_s2_params_get.argtypes = [vp, C.POINTER(Scheme2ParamsC)]
        L.sgfhe_s2_polymul.argtypes = [vp, i32, u64p, u64p, u64p, u64p, u64p, u64p]
        L.sgfhe_s2_polymul_device.argtypes = [vp, i32, u64p, u64p, u64p, u64p, i32, u64p, u64p, vp]
        L.sgfhe_s2_ntt_device.argtypes = [vp, i32, i32, u64p, u64p, vp]
        L.sgfhe_s2_mac8_device.argtypes = [vp, i32, u64p, u64p, u64p, u64p, u64p, u64p, vp]
        L.sgfhe_s2_bkey_generate.argtypes = [vp, vp, u64p, vp, i32, i32, u64p]
        L.sgfhe_scheme2_params_derive.argtypes = [i32, C.POINTER(Scheme2ParamsC)]
        L.sgfhe_rns2_op.argtypes = [i32, i32, C.c_uint64, u64p, u64p, u64p, u64p, C.c_uint64, C.c_uint64, u64p, u64p]
        L.sgfhe_rns2_op_device.argtypes = [i32, i32, C.c_uint64, u64p, u64p, u64p, u64p, C.c_uint64, C.c_uint64, u64p, u64p, vp]
        _LIB = L
    return _LIB


class Scheme2ParamsC(C.Structure):
    _fields_ = [("n", C.c_int32), ("k", C.c_int32), ("t", C.c_int32), ("pad", C.c_int32)] + \
               [(f, C.c_uint64) for f in ("r", "m", "q", "tau", "B", "Bp", "Dr", "Dq")]


class SgfheError(RuntimeError):
    """Raised where the reference would throw (AssertionError / ErrorException)."""


def check(rc: int):
    if rc != 0:
        raise SgfheError(f"libsgfhe_cuda status {rc}: {lib().sgfhe_last_error().decode()}")
