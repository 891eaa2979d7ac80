"""Builds serialised server keys byte for byte from the serde field order of the reference's structs (bincode 1.x
default options: little endian, fixed-width integers, usize and sequence lengths as u64, enum variant index as u32,
u128 as 16 bytes).  PARITY UNPINNED: no Rust toolchain here, so these bytes were never produced by the reference
itself; this script and tfhe_rs_string_b200/csrc/key_import.hpp are two independent readings of

  shortint/server_key/mod.rs:283-297          ServerKey { key_switching_key, bootstrapping_key, message_modulus,
                                               carry_modulus, max_degree, max_noise_level, ciphertext_modulus, pbs_order }
  shortint/server_key/mod.rs:113-121          SerializableShortintBootstrappingKey::Classic = variant 0
  core_crypto/entities/lwe_keyswitch_key.rs:76-86, fft_impl/fft64/crypto/bootstrap.rs:24-32,
  fft_impl/fft64/math/fft/mod.rs:588-630      FourierPolynomialList: seq(2 + count): polynomial_size, count, polynomials
  core_crypto/commons/ciphertext_modulus.rs:41-63  { modulus: u128 (0 = native), scalar_bits: usize }
  core_crypto/entities/ggsw_ciphertext_list.rs:9-20  (standard-domain bundle)

Run from the repo root to regenerate tests/golden/server_key_toy_{fourier,std}.bin (toy parameters, a few KB):
    python tests/golden/make_server_key_fixture.py
The GPU test builds the full-size PARAM_MESSAGE_2_CARRY_2 blobs in memory with the same functions."""
import os
import struct
import sys

import numpy as np

U64 = "<Q"


def _u64(x):
    return struct.pack(U64, int(x))


def _modulus_native_u64():
    return struct.pack("<QQ", 0, 0) + _u64(64)          # u128 = 0 (native), scalar_bits = 64


def _ksk(p, ksk):
    ksk = np.ascontiguousarray(ksk, dtype="<u8")
    return (_u64(ksk.size) + ksk.tobytes() + _u64(p.ks_base_log) + _u64(p.ks_level) + _u64(p.lwe_dimension + 1)
            + _modulus_native_u64())


def fourier_natural_order(p, bsk_standard):
    """Every polynomial of the standard key in the Fourier domain, natural frequency order:
    F[k] = sum_j z_j exp(-2 pi i j k / (N/2)), z_j = (p_j + i p_{j+N/2}) / 2^64 * exp(i pi j / N)  (fft/mod.rs:197-218)."""
    N = p.polynomial_size
    polys = np.ascontiguousarray(bsk_standard, dtype=np.uint64).reshape(-1, N).view(np.int64)
    half = N // 2
    z = (polys[:, :half].astype(np.float64) + 1j * polys[:, half:].astype(np.float64)) * 2.0 ** -64
    z = z * np.exp(1j * np.pi * np.arange(half) / N)[None, :]
    return np.fft.fft(z, axis=1)


def serialize_server_key_fourier(p, ksk, bsk_standard, pbs_order=0):
    """bincode(shortint::ServerKey) with a Classic Fourier bootstrap key."""
    N, half = p.polynomial_size, p.polynomial_size // 2
    f = fourier_natural_order(p, bsk_standard)
    count = f.shape[0]
    out = [_ksk(p, ksk), struct.pack("<I", 0), _u64(2 + count), _u64(N), _u64(count)]
    inter = np.empty((count, half, 2), dtype="<f8")
    inter[:, :, 0], inter[:, :, 1] = f.real, f.imag
    for c in range(count):
        out.append(_u64(half))
        out.append(inter[c].tobytes())
    out += [_u64(p.lwe_dimension), _u64(p.glwe_dimension + 1), _u64(p.pbs_base_log), _u64(p.pbs_level)]
    ms = p.message_modulus * p.carry_modulus
    out += [_u64(p.message_modulus), _u64(p.carry_modulus), _u64(ms - 1), _u64((ms - 1) // (p.message_modulus - 1) if p.message_modulus > 1 else 1),
            _modulus_native_u64(), struct.pack("<I", pbs_order)]
    return b"".join(out)


def serialize_std_bundle(p, ksk, bsk_standard, pbs_order=0):
    """bincode((LweKeyswitchKey<Vec<u64>>, LweBootstrapKey<Vec<u64>>, MessageModulus, CarryModulus, PBSOrder))."""
    bsk = np.ascontiguousarray(bsk_standard, dtype="<u8")
    return b"".join([_ksk(p, ksk), _u64(bsk.size), bsk.tobytes(), _u64(p.glwe_dimension + 1), _u64(p.polynomial_size),
                     _u64(p.pbs_base_log), _u64(p.pbs_level), _modulus_native_u64(),
                     _u64(p.message_modulus), _u64(p.carry_modulus), struct.pack("<I", pbs_order)])


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, root)
    from oracle import oracle as O
    p = O.params_toy(4, 256)
    keys = O.Keyset(p, seed=11)
    here = os.path.dirname(os.path.abspath(__file__))
    open(os.path.join(here, "server_key_toy_fourier.bin"), "wb").write(serialize_server_key_fourier(p, keys.ksk, keys.bsk_standard))
    open(os.path.join(here, "server_key_toy_std.bin"), "wb").write(serialize_std_bundle(p, keys.ksk, keys.bsk_standard, pbs_order=0))
    print("toy params:", p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.pbs_base_log, p.pbs_level, p.ks_base_log, p.ks_level,
          p.message_modulus, p.carry_modulus)
