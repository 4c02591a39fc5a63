"""sgfhe.jl_b200 -- B200-native backend for SGFHE.jl's bootstrapping hot path.

Imported as `sgfhe_jl_b200` through the shim at the repository root.
"""
from ._lib import SO_PATH, SgfheError, build
from .chain import bootstrap_chain
from .parallel import bootstrap_sharded, broadcast_key, exchange_layer, shard_bounds
from .api import (DeviceRng, Scheme2Params, Scheme2Context, scheme2_bootstrap_key, rns2_op, BootstrapKey, Ciphertext, EncryptedBit, LWE, PackedCiphertext, Params, PrivateKey, bootstrap,
                  bootstrap_batch, bootstrap_trace, decrypt, encrypt, external_product, flatten_poly,
                  launch_count, pack_encrypted_bits, polymul, split_ciphertext, split_ciphertexts, decrypt_bits)

__all__ = ["DeviceRng", "Scheme2Params", "Scheme2Context", "scheme2_bootstrap_key", "rns2_op", "Params", "PrivateKey", "BootstrapKey", "encrypt", "decrypt", "split_ciphertext", "bootstrap",
           "bootstrap_batch", "bootstrap_trace", "pack_encrypted_bits", "Ciphertext", "polymul", "flatten_poly", "external_product",
           "bootstrap_chain", "bootstrap_sharded", "broadcast_key", "exchange_layer", "shard_bounds", "EncryptedBit", "LWE", "PackedCiphertext", "SgfheError", "build", "launch_count", "SO_PATH",
           "split_ciphertexts", "decrypt_bits"]
