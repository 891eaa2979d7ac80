"""ctypes binding of oracle/_build/liboracle.so (CPU restatement of the reference's KS+PBS path).

TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


class Params(C.Structure):
    """Mirror of orc_params (shortint/parameters/mod.rs:62-76 of the reference)."""
    _fields_ = [
        ("lwe_dimension", C.c_uint32),
        ("glwe_dimension", C.c_uint32),
        ("polynomial_size", C.c_uint32),
        ("lwe_modular_std_dev", C.c_double),
        ("glwe_modular_std_dev", C.c_double),
        ("pbs_base_log", C.c_uint32),
        ("pbs_level", C.c_uint32),
        ("ks_base_log", C.c_uint32),
        ("ks_level", C.c_uint32),
        ("message_modulus", C.c_uint32),
        ("carry_modulus", C.c_uint32),
    ]

    @property
    def big_lwe_size(self):
        return self.glwe_dimension * self.polynomial_size + 1

    @property
    def small_lwe_size(self):
        return self.lwe_dimension + 1

    @property
    def glwe_len(self):
        return (self.glwe_dimension + 1) * self.polynomial_size

    @property
    def delta(self):
        return (1 << 63) // (self.message_modulus * self.carry_modulus)


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("tfhe_oracle.cpp", "tfhe_oracle.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    PP = C.POINTER(Params)
    KS = C.c_void_p
    sig = {
        "orc_params_message_2_carry_2": (None, [PP]),
        "orc_params_toy": (None, [PP, C.c_uint32, C.c_uint32]),
        "orc_keyset_create": (KS, [PP, C.c_uint64, C.c_int]),
        "orc_keyset_destroy": (None, [KS]),
        "orc_keyset_small_sk": (C.POINTER(C.c_uint64), [KS]),
        "orc_keyset_big_sk": (C.POINTER(C.c_uint64), [KS]),
        "orc_keyset_ksk": (C.POINTER(C.c_uint64), [KS]),
        "orc_keyset_ksk_len": (C.c_size_t, [KS]),
        "orc_keyset_bsk_standard": (C.POINTER(C.c_uint64), [KS]),
        "orc_keyset_bsk_len": (C.c_size_t, [KS]),
        "orc_keyset_bsk_fourier": (C.POINTER(C.c_double), [KS]),
        "orc_closest_representable": (C.c_uint64, [C.c_uint64, C.c_uint32, C.c_uint32]),
        "orc_decompose": (None, [C.c_uint64, C.c_uint32, C.c_uint32, _i64p]),
        "orc_modulus_switch": (C.c_uint64, [C.c_uint64, C.c_uint32]),
        "orc_monomial_div": (None, [_u64p, _u64p, C.c_size_t, C.c_size_t]),
        "orc_monomial_mul": (None, [_u64p, _u64p, C.c_size_t, C.c_size_t]),
        "orc_monomial_mul_and_subtract": (None, [_u64p, _u64p, C.c_size_t, C.c_size_t]),
        "orc_sample_extract0": (None, [_u64p, _u64p, C.c_uint32, C.c_uint32]),
        "orc_fft_forward_integer": (None, [_f64p, _u64p, C.c_uint32]),
        "orc_fft_forward_torus": (None, [_f64p, _u64p, C.c_uint32]),
        "orc_fft_add_backward_torus": (None, [_u64p, _f64p, C.c_uint32]),
        "orc_from_torus": (C.c_uint64, [C.c_double]),
        "orc_keyswitch": (None, [KS, _u64p, _u64p]),
        "orc_keyswitch_raw": (None, [_u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _u64p, _u64p]),
        "orc_blind_rotate": (None, [KS, _u64p, _u64p]),
        "orc_bootstrap": (None, [KS, _u64p, _u64p, _u64p]),
        "orc_ks_pbs": (None, [KS, _u64p, _u64p, _u64p]),
        "orc_bootstrap_steps": (None, [KS, _u64p, C.c_uint32, _u64p, _u64p]),
        "orc_ks_pbs_batch": (None, [KS, _u64p, _u64p, _u32p, _u64p, C.c_size_t, C.c_int]),
        "orc_fill_accumulator": (C.c_uint64, [PP, _u64p, _u64p]),
        "orc_trivial_pbs": (C.c_uint64, [PP, C.c_uint64, _u64p]),
        "orc_encrypt_seeded": (None, [KS, C.c_uint64, C.c_uint64, _u64p]),
        "orc_encrypt_batch_seeded": (None, [KS, _u64p, C.c_size_t, C.c_uint64, _u64p]),
        "orc_decrypt_phase": (C.c_uint64, [KS, _u64p]),
        "orc_decrypt_small_phase": (C.c_uint64, [KS, _u64p]),
        "orc_decrypt_message_and_carry": (C.c_uint64, [KS, _u64p]),
        "orc_decrypt_batch": (None, [KS, _u64p, C.c_size_t, _u64p]),
        "orc_circuit_last_error": (C.c_char_p, []),
        "orc_circuit_info": (C.c_int, [C.c_char_p, _u64p, C.c_size_t, C.c_uint32, C.c_uint32, _u64p]),
        "orc_circuit_run_cleartext": (C.c_int, [C.c_char_p, _u64p, C.c_size_t, C.c_uint32, C.c_uint32, _u64p, _u64p]),
        "orc_circuit_run_encrypted": (C.c_int, [KS, C.c_char_p, _u64p, C.c_size_t, _u64p, _u64p, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def params_message_2_carry_2():
    p = Params()
    lib().orc_params_message_2_carry_2(C.byref(p))
    return p


def params_message_2_carry_2_pbs_ks():
    """PARAM_MESSAGE_2_CARRY_2_PBS_KS (shortint/parameters/mod.rs:1155-1169): PBS -> KS order."""
    p = params_message_2_carry_2()
    p.lwe_dimension = 870
    p.lwe_modular_std_dev = 0.0000006791658447437413
    p.glwe_modular_std_dev = 0.00000000000000029403601535432533
    p.ks_base_log, p.ks_level = 4, 4
    return p


def params_message_1_carry_1():
    """PARAM_MESSAGE_1_CARRY_1_KS_PBS (shortint/parameters/mod.rs:613-627): k = 3, N = 512."""
    p = params_message_2_carry_2()
    p.lwe_dimension, p.glwe_dimension, p.polynomial_size = 684, 3, 512
    p.lwe_modular_std_dev, p.glwe_modular_std_dev = 0.00002043784477291318, 0.0000000000034525330484572114
    p.pbs_base_log, p.pbs_level, p.ks_base_log, p.ks_level = 18, 1, 4, 3
    p.message_modulus, p.carry_modulus = 2, 2
    return p


def params_message_3_carry_3():
    """PARAM_MESSAGE_3_CARRY_3_KS_PBS (shortint/parameters/mod.rs:853-867): N = 8192, two PBS levels."""
    p = params_message_2_carry_2()
    p.lwe_dimension, p.glwe_dimension, p.polynomial_size = 864, 1, 8192
    p.lwe_modular_std_dev, p.glwe_modular_std_dev = 0.000000757998020150446, 0.0000000000000000002168404344971009
    p.pbs_base_log, p.pbs_level, p.ks_base_log, p.ks_level = 15, 2, 3, 6
    p.message_modulus, p.carry_modulus = 8, 8
    return p


def params_toy(n=16, N=256):
    p = Params()
    lib().orc_params_toy(C.byref(p), n, N)
    return p


def _view(ptr, n, dtype):
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)


class Keyset:
    """Client + server keys of the oracle; arrays are views into oracle-owned memory."""

    def __init__(self, params, seed=0xB200, n_threads=None):
        self.L = lib()
        self.params = params
        n_threads = n_threads or os.cpu_count() or 1
        self.h = self.L.orc_keyset_create(C.byref(params), seed, n_threads)
        p = params
        self.small_sk = _view(self.L.orc_keyset_small_sk(self.h), p.lwe_dimension, np.uint64)
        self.big_sk = _view(self.L.orc_keyset_big_sk(self.h), p.glwe_dimension * p.polynomial_size, np.uint64)
        self.ksk = _view(self.L.orc_keyset_ksk(self.h), self.L.orc_keyset_ksk_len(self.h), np.uint64)
        self.bsk_standard = _view(self.L.orc_keyset_bsk_standard(self.h), self.L.orc_keyset_bsk_len(self.h), np.uint64)
        self.bsk_fourier = _view(self.L.orc_keyset_bsk_fourier(self.h), self.L.orc_keyset_bsk_len(self.h), np.float64)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_keyset_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # -- shortint layer
    def lut(self, f):
        """generate_lookup_table (shortint/server_key/mod.rs:383-399): GLWE accumulator of f."""
        p = self.params
        table = np.array([f(i) for i in range(p.message_modulus * p.carry_modulus)], dtype=np.uint64)
        out = np.zeros(p.glwe_len, dtype=np.uint64)
        self.L.orc_fill_accumulator(C.byref(p), table, out)
        return out

    def encrypt_batch(self, messages, seed=0xC0FFEE):
        m = np.ascontiguousarray(messages, dtype=np.uint64)
        out = np.zeros((len(m), self.params.big_lwe_size), dtype=np.uint64)
        self.L.orc_encrypt_batch_seeded(self.h, m, len(m), seed, out)
        return out

    def decrypt_batch(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64)
        out = np.zeros(cts.shape[0], dtype=np.uint64)
        self.L.orc_decrypt_batch(self.h, cts, cts.shape[0], out)
        return out

    def phase_batch(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64)
        return np.array([self.L.orc_decrypt_phase(self.h, cts[i]) for i in range(cts.shape[0])], dtype=np.uint64)

    def small_phase_batch(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64)
        return np.array([self.L.orc_decrypt_small_phase(self.h, cts[i]) for i in range(cts.shape[0])], dtype=np.uint64)

    # -- hot path
    def keyswitch_batch(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64)
        out = np.zeros((cts.shape[0], self.params.small_lwe_size), dtype=np.uint64)
        for i in range(cts.shape[0]):
            self.L.orc_keyswitch(self.h, cts[i], out[i])
        return out

    def bootstrap_batch(self, small_cts, luts, lut_idx=None):
        small_cts = np.ascontiguousarray(small_cts, dtype=np.uint64)
        luts = np.ascontiguousarray(luts, dtype=np.uint64).reshape(-1, self.params.glwe_len)
        out = np.zeros((small_cts.shape[0], self.params.big_lwe_size), dtype=np.uint64)
        for i in range(small_cts.shape[0]):
            li = 0 if lut_idx is None else int(lut_idx[i])
            self.L.orc_bootstrap(self.h, small_cts[i], luts[li], out[i])
        return out

    def bootstrap_steps_batch(self, small_prefix, steps, lut):
        """The bootstrap stopped after `steps` CMUX steps (small_prefix: batch x (steps + 1): mask prefix, body)."""
        cts = np.ascontiguousarray(small_prefix, dtype=np.uint64).reshape(-1, steps + 1)
        lut = np.ascontiguousarray(lut, dtype=np.uint64)
        out = np.zeros((cts.shape[0], self.params.big_lwe_size), dtype=np.uint64)
        for i in range(cts.shape[0]):
            self.L.orc_bootstrap_steps(self.h, cts[i], steps, lut, out[i])
        return out

    def ks_pbs_batch(self, cts, luts, lut_idx=None, n_threads=None):
        cts = np.ascontiguousarray(cts, dtype=np.uint64)
        luts = np.ascontiguousarray(luts, dtype=np.uint64).reshape(-1, self.params.glwe_len)
        B = cts.shape[0]
        idx = np.zeros(B, dtype=np.uint32) if lut_idx is None else np.ascontiguousarray(lut_idx, dtype=np.uint32)
        out = np.zeros((B, self.params.big_lwe_size), dtype=np.uint64)
        self.L.orc_ks_pbs_batch(self.h, cts, luts, idx, out, B, n_threads or os.cpu_count() or 1)
        return out


# ---- leveled circuits (host logic of the product, executed on the CPU for tests) -----------------
def _shape(shape):
    return np.ascontiguousarray(shape, dtype=np.uint64)


def circuit_info(op, shape, message_modulus=4, carry_modulus=4):
    info = np.zeros(6, dtype=np.uint64)
    sh = _shape(shape)
    if lib().orc_circuit_info(op.encode(), sh, len(sh), message_modulus, carry_modulus, info):
        raise RuntimeError(lib().orc_circuit_last_error().decode())
    return dict(zip(["n_inputs", "n_outputs", "n_pbs", "depth", "n_stages", "n_luts"], (int(x) for x in info)))


def circuit_run_cleartext(op, shape, in_msgs, message_modulus=4, carry_modulus=4):
    """Runs a named program (tfhe_rs_string_b200/csrc/programs.hpp) on clear message values."""
    info = circuit_info(op, shape, message_modulus, carry_modulus)
    m = np.ascontiguousarray(in_msgs, dtype=np.uint64).ravel()
    assert m.size == info["n_inputs"], (m.size, info)
    out = np.zeros(info["n_outputs"], dtype=np.uint64)
    sh = _shape(shape)
    if lib().orc_circuit_run_cleartext(op.encode(), sh, len(sh), message_modulus, carry_modulus, m, out):
        raise RuntimeError(lib().orc_circuit_last_error().decode())
    return out


def circuit_run_encrypted(keys, op, shape, in_cts, n_threads=None):
    """Same program on ciphertexts with the CPU oracle's KS+PBS (use toy parameters for speed)."""
    p = keys.params
    info = circuit_info(op, shape, p.message_modulus, p.carry_modulus)
    cts = np.ascontiguousarray(in_cts, dtype=np.uint64).reshape(-1, p.big_lwe_size)
    assert cts.shape[0] == info["n_inputs"]
    out = np.zeros((info["n_outputs"], p.big_lwe_size), dtype=np.uint64)
    sh = _shape(shape)
    if lib().orc_circuit_run_encrypted(keys.h, op.encode(), sh, len(sh), cts, out, n_threads or os.cpu_count() or 1):
        raise RuntimeError(lib().orc_circuit_last_error().decode())
    return out
