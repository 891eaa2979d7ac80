"""ctypes binding of libb200tfhe.so (C ABI declared in include/b200tfhe.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# every symbol include/b200tfhe.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "b200tfhe_ctx_create", "b200tfhe_ctx_create_multi", "b200tfhe_ctx_device_count", "b200tfhe_ctx_destroy", "b200tfhe_last_error", "b200tfhe_last_global_error",
    "b200tfhe_load_ksk", "b200tfhe_load_bsk_standard", "b200tfhe_key_arena", "b200tfhe_keys_adopt",
    "b200tfhe_register_lut", "b200tfhe_register_lut_from_table",
    "b200tfhe_keyswitch_batch", "b200tfhe_pbs_batch", "b200tfhe_ks_pbs_batch", "b200tfhe_pbs_ks_batch",
    "b200tfhe_keyswitch_batch_device", "b200tfhe_pbs_batch_device", "b200tfhe_ks_pbs_batch_device",
    "b200tfhe_pbs_ks_batch_device",
    "b200tfhe_lwe_linear_batch_device", "b200tfhe_sync", "b200tfhe_stream",
    "b200tfhe_ks_pbs_batch_device_multi",
    "b200tfhe_set_profiling", "b200tfhe_get_kernel_times", "b200tfhe_kernel_launch_count",
    "b200tfhe_parse_server_key", "b200tfhe_load_server_key_bytes",
    "b200tfhe_boolean_ctx_create", "b200tfhe_boolean_ctx_destroy", "b200tfhe_boolean_last_error",
    "b200tfhe_boolean_load_ksk", "b200tfhe_boolean_load_bsk_standard", "b200tfhe_boolean_gate_batch",
    "b200tfhe_debug_negacyclic_mul", "b200tfhe_debug_pbs_steps", "b200tfhe_debug_from_torus",
    "b200tfhe_program_create", "b200tfhe_program_create_from_circuit", "b200tfhe_program_info", "b200tfhe_program_run", "b200tfhe_program_run_device",
    "b200tfhe_program_destroy",
]


class B200TfheError(RuntimeError):
    pass


class KeyView(C.Structure):
    """b200tfhe_key_view: where the key material sits inside a serialised server key."""
    _fields_ = [
        ("ksk_offset", C.c_uint64), ("ksk_len", C.c_uint64), ("bsk_offset", C.c_uint64), ("bsk_len", C.c_uint64),
        ("bsk_poly_stride_bytes", C.c_uint64), ("bsk_is_fourier", C.c_uint32), ("pbs_order", C.c_uint32),
        ("max_degree", C.c_uint64), ("max_noise_level", C.c_uint64),
    ]


class CircuitDesc(C.Structure):
    """b200tfhe_circuit_desc: a caller-built level schedule (see include/b200tfhe.h)."""
    _fields_ = [
        ("n_inputs", C.c_size_t), ("n_nodes", C.c_size_t), ("n_luts", C.c_size_t), ("n_outputs", C.c_size_t),
        ("node_term_begin", C.c_void_p), ("term_block", C.c_void_p), ("term_coeff", C.c_void_p),
        ("node_plaintext", C.c_void_p), ("node_lut", C.c_void_p), ("luts", C.c_void_p), ("outputs", C.c_void_p),
    ]


class Params(C.Structure):
    """b200tfhe_params == ClassicPBSParameters of the reference (shortint/parameters/mod.rs:62-76)."""
    _fields_ = [
        ("lwe_dimension", C.c_uint32), ("glwe_dimension", C.c_uint32), ("polynomial_size", C.c_uint32),
        ("pbs_base_log", C.c_uint32), ("pbs_level", C.c_uint32),
        ("ks_base_log", C.c_uint32), ("ks_level", C.c_uint32),
        ("message_modulus", C.c_uint32), ("carry_modulus", C.c_uint32),
    ]

    @classmethod
    def message_2_carry_2_pbs_ks(cls):
        """PARAM_MESSAGE_2_CARRY_2_PBS_KS (shortint/parameters/mod.rs:1155-1169)."""
        return cls(870, 1, 2048, 23, 1, 4, 4, 4, 4)

    @classmethod
    def message_2_carry_2(cls):
        """PARAM_MESSAGE_2_CARRY_2_KS_PBS (shortint/parameters/mod.rs:703-717)."""
        return cls(742, 1, 2048, 23, 1, 3, 5, 4, 4)

    @property
    def big_lwe_size(self):
        return self.glwe_dimension * self.polynomial_size + 1

    @property
    def small_lwe_size(self):
        return self.lwe_dimension + 1

    @property
    def glwe_len(self):
        return (self.glwe_dimension + 1) * self.polynomial_size


def lib_path():
    return os.path.join(_HERE, "libb200tfhe.so")


def load_library():
    """Loads libb200tfhe.so; raises (never falls back) when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise B200TfheError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    L = C.CDLL(path)
    u64p, u32p, vp = C.c_void_p, C.c_void_p, C.c_void_p   # raw addresses (host or device)
    ctx = C.c_void_p
    sig = {
        "b200tfhe_ctx_create": [C.POINTER(Params), C.c_int, C.POINTER(ctx)],
        "b200tfhe_ctx_create_multi": [C.POINTER(Params), C.POINTER(C.c_int), C.c_int, C.POINTER(ctx)],
        "b200tfhe_ctx_device_count": [ctx, C.POINTER(C.c_int)],
        "b200tfhe_ks_pbs_batch_device_multi": [ctx, vp, vp, vp, vp],
        "b200tfhe_kernel_launch_count": [ctx, C.POINTER(C.c_uint64)],
        "b200tfhe_parse_server_key": [vp, C.c_size_t, C.POINTER(Params), C.POINTER(KeyView)],
        "b200tfhe_load_server_key_bytes": [ctx, vp, C.c_size_t],
        "b200tfhe_debug_pbs_steps": [ctx, u64p, u32p, u64p, C.c_size_t, C.c_uint32],
        "b200tfhe_boolean_ctx_create": [C.POINTER(Params), C.c_int, C.c_int, C.POINTER(ctx)],
        "b200tfhe_boolean_ctx_destroy": [ctx],
        "b200tfhe_boolean_last_error": [ctx, C.c_char_p, C.c_size_t],
        "b200tfhe_boolean_load_ksk": [ctx, vp, C.c_size_t],
        "b200tfhe_boolean_load_bsk_standard": [ctx, vp, C.c_size_t],
        "b200tfhe_boolean_gate_batch": [ctx, C.c_int, vp, vp, vp, C.c_size_t],
        "b200tfhe_program_create_from_circuit": [ctx, C.POINTER(CircuitDesc), C.POINTER(C.c_void_p)],
        "b200tfhe_ctx_destroy": [ctx],
        "b200tfhe_last_error": [ctx, C.c_char_p, C.c_size_t],
        "b200tfhe_last_global_error": [C.c_char_p, C.c_size_t],
        "b200tfhe_load_ksk": [ctx, u64p, C.c_size_t],
        "b200tfhe_load_bsk_standard": [ctx, u64p, C.c_size_t],
        "b200tfhe_key_arena": [ctx, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)],
        "b200tfhe_keys_adopt": [ctx],
        "b200tfhe_register_lut": [ctx, u64p, C.POINTER(C.c_uint32)],
        "b200tfhe_register_lut_from_table": [ctx, u64p, C.c_size_t, C.POINTER(C.c_uint32)],
        "b200tfhe_keyswitch_batch": [ctx, u64p, u64p, C.c_size_t],
        "b200tfhe_pbs_batch": [ctx, u64p, u32p, u64p, C.c_size_t],
        "b200tfhe_ks_pbs_batch": [ctx, u64p, u32p, u64p, C.c_size_t],
        "b200tfhe_keyswitch_batch_device": [ctx, u64p, u64p, C.c_size_t],
        "b200tfhe_pbs_batch_device": [ctx, u64p, u32p, u64p, C.c_size_t],
        "b200tfhe_ks_pbs_batch_device": [ctx, u64p, u32p, u64p, C.c_size_t],
        "b200tfhe_pbs_ks_batch": [ctx, u64p, u32p, u64p, C.c_size_t],
        "b200tfhe_pbs_ks_batch_device": [ctx, u64p, u32p, u64p, C.c_size_t],
        "b200tfhe_lwe_linear_batch_device": [ctx, vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, C.c_size_t],
        "b200tfhe_sync": [ctx],
        "b200tfhe_stream": [ctx, C.POINTER(C.c_void_p)],
        "b200tfhe_set_profiling": [ctx, C.c_int],
        "b200tfhe_get_kernel_times": [ctx, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_double),
                                      C.POINTER(C.c_uint64), C.c_int],
        "b200tfhe_debug_negacyclic_mul": [ctx, u64p, u64p, u64p, C.c_size_t],
        "b200tfhe_debug_from_torus": [ctx, C.c_void_p, u64p, u64p, C.c_size_t],
        "b200tfhe_program_create": [ctx, C.c_char_p, u64p, C.c_size_t, C.POINTER(C.c_void_p)],
        "b200tfhe_program_info": [C.c_void_p, u64p],
        "b200tfhe_program_run": [C.c_void_p, u64p, u64p],
        "b200tfhe_program_run_device": [C.c_void_p, u64p, u64p],
        "b200tfhe_program_destroy": [C.c_void_p],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _LIB = L
    return L


def _ptr(x):
    """Address of a numpy array (host) or torch tensor (host or device); None -> NULL."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return x.data_ptr()
    if isinstance(x, int):
        return x
    raise TypeError(type(x))


class Engine:
    """One context per GPU.  Mirrors the call surface of shortint::ServerKey for the KS+PBS path:
    keyswitch / programmable bootstrap / apply_lookup_table, batched."""

    def __init__(self, params=None, device=0, devices=None):
        """device: one GPU; devices=[...]: one context over several GPUs of the box (b200tfhe_ctx_create_multi)."""
        self.L = load_library()
        self.params = params or Params.message_2_carry_2()
        self.devices = list(devices) if devices is not None else [device]
        self.device = self.devices[0]
        h = C.c_void_p()
        devs = (C.c_int * len(self.devices))(*self.devices)
        rc = self.L.b200tfhe_ctx_create_multi(C.byref(self.params), devs, len(self.devices), C.byref(h))
        if rc != 0:
            buf = C.create_string_buffer(1024)
            self.L.b200tfhe_last_global_error(buf, 1024)
            raise B200TfheError(buf.value.decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.b200tfhe_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            buf = C.create_string_buffer(1024)
            self.L.b200tfhe_last_error(self.h, buf, 1024)
            raise B200TfheError(buf.value.decode())

    # ---- keys / LUTs
    def load_ksk(self, ksk):
        ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
        self._check(self.L.b200tfhe_load_ksk(self.h, _ptr(ksk), ksk.size))

    def load_bsk_standard(self, bsk):
        bsk = np.ascontiguousarray(bsk, dtype=np.uint64)
        self._check(self.L.b200tfhe_load_bsk_standard(self.h, _ptr(bsk), bsk.size))

    def key_arena(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.L.b200tfhe_key_arena(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def keys_adopt(self):
        self._check(self.L.b200tfhe_keys_adopt(self.h))

    def register_lut(self, glwe_acc):
        acc = np.ascontiguousarray(glwe_acc, dtype=np.uint64)
        assert acc.size == self.params.glwe_len
        i = C.c_uint32()
        self._check(self.L.b200tfhe_register_lut(self.h, _ptr(acc), C.byref(i)))
        return i.value

    def register_lut_from_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.uint64)
        i = C.c_uint32()
        self._check(self.L.b200tfhe_register_lut_from_table(self.h, _ptr(t), t.size, C.byref(i)))
        return i.value

    def generate_lookup_table(self, f):
        """ServerKey::generate_lookup_table (shortint/server_key/mod.rs:383-399) -> device LUT id."""
        m = self.params.message_modulus * self.params.carry_modulus
        return self.register_lut_from_table([f(x) for x in range(m)])

    # ---- host-buffer hot path
    def keyswitch_batch(self, cts):
        cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, self.params.big_lwe_size)
        out = np.empty((cts.shape[0], self.params.small_lwe_size), dtype=np.uint64)
        self._check(self.L.b200tfhe_keyswitch_batch(self.h, _ptr(cts), _ptr(out), cts.shape[0]))
        return out

    def pbs_batch(self, small_cts, lut_ids=None):
        cts = np.ascontiguousarray(small_cts, dtype=np.uint64).reshape(-1, self.params.small_lwe_size)
        ids = None if lut_ids is None else np.ascontiguousarray(lut_ids, dtype=np.uint32)
        out = np.empty((cts.shape[0], self.params.big_lwe_size), dtype=np.uint64)
        self._check(self.L.b200tfhe_pbs_batch(self.h, _ptr(cts), _ptr(ids), _ptr(out), cts.shape[0]))
        return out

    def ks_pbs_batch(self, cts, lut_ids=None, out=None):
        """apply_lookup_table over a batch (host buffers; numpy arrays or pinned torch tensors)."""
        if isinstance(cts, np.ndarray):
            cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(-1, self.params.big_lwe_size)
            batch = cts.shape[0]
            if out is None:
                out = np.empty((batch, self.params.big_lwe_size), dtype=np.uint64)
        else:
            batch = cts.shape[0]
            assert out is not None
        ids = lut_ids
        if ids is not None and isinstance(ids, (list, tuple)):
            ids = np.ascontiguousarray(ids, dtype=np.uint32)
        self._check(self.L.b200tfhe_ks_pbs_batch(self.h, _ptr(cts), _ptr(ids), _ptr(out), batch))
        return out

    def pbs_ks_batch(self, small_cts, lut_ids=None):
        """apply_lookup_table for PBSOrder::BootstrapKeyswitch over a batch of small-key ciphertexts."""
        cts = np.ascontiguousarray(small_cts, dtype=np.uint64).reshape(-1, self.params.small_lwe_size)
        ids = None if lut_ids is None else np.ascontiguousarray(lut_ids, dtype=np.uint32)
        out = np.empty_like(cts)
        self._check(self.L.b200tfhe_pbs_ks_batch(self.h, _ptr(cts), _ptr(ids), _ptr(out), cts.shape[0]))
        return out

    def pbs_ks_batch_device(self, d_in, d_lut_ids, d_out, batch):
        self._check(self.L.b200tfhe_pbs_ks_batch_device(self.h, _ptr(d_in), _ptr(d_lut_ids), _ptr(d_out), batch))

    # ---- device-buffer hot path (torch CUDA tensors or raw device addresses)
    def keyswitch_batch_device(self, d_in, d_out, batch):
        self._check(self.L.b200tfhe_keyswitch_batch_device(self.h, _ptr(d_in), _ptr(d_out), batch))

    def pbs_batch_device(self, d_in, d_lut_ids, d_out, batch):
        self._check(self.L.b200tfhe_pbs_batch_device(self.h, _ptr(d_in), _ptr(d_lut_ids), _ptr(d_out), batch))

    def ks_pbs_batch_device(self, d_in, d_lut_ids, d_out, batch):
        self._check(self.L.b200tfhe_ks_pbs_batch_device(self.h, _ptr(d_in), _ptr(d_lut_ids), _ptr(d_out), batch))

    def lwe_linear_batch_device(self, d_x, d_y, d_ia, d_ib, d_ca, d_cb, d_pt, d_out, batch, lwe_size):
        self._check(self.L.b200tfhe_lwe_linear_batch_device(
            self.h, _ptr(d_x), _ptr(d_y), _ptr(d_ia), _ptr(d_ib), _ptr(d_ca), _ptr(d_cb), _ptr(d_pt),
            _ptr(d_out), batch, lwe_size))

    def sync(self):
        self._check(self.L.b200tfhe_sync(self.h))

    def stream(self):
        s = C.c_void_p()
        self._check(self.L.b200tfhe_stream(self.h, C.byref(s)))
        return s.value

    # ---- measurement
    def set_profiling(self, on):
        self._check(self.L.b200tfhe_set_profiling(self.h, 1 if on else 0))

    def kernel_times(self, reset=False):
        a, b = C.c_double(), C.c_double()
        na, nb = C.c_uint64(), C.c_uint64()
        self._check(self.L.b200tfhe_get_kernel_times(self.h, C.byref(a), C.byref(na), C.byref(b), C.byref(nb),
                                                     1 if reset else 0))
        return {"ks_ms": a.value, "ks_launches": na.value, "pbs_ms": b.value, "pbs_launches": nb.value}

    def kernel_launch_count(self):
        n = C.c_uint64()
        self._check(self.L.b200tfhe_kernel_launch_count(self.h, C.byref(n)))
        return n.value

    def load_server_key_bytes(self, blob):
        buf = np.frombuffer(blob, dtype=np.uint8)
        self._check(self.L.b200tfhe_load_server_key_bytes(self.h, buf.ctypes.data, buf.size))

    def ks_pbs_batch_device_multi(self, d_ins, d_lut_ids, d_outs, batches):
        """One device-resident shard per GPU of the context (lists of torch tensors / device addresses)."""
        n = len(self.devices)
        arr = lambda xs: (C.c_void_p * n)(*[_ptr(x) for x in xs])
        a_in, a_out = arr(d_ins), arr(d_outs)
        a_id = arr(d_lut_ids) if d_lut_ids is not None else None
        a_b = (C.c_size_t * n)(*batches)
        self._check(self.L.b200tfhe_ks_pbs_batch_device_multi(self.h, a_in, a_id, a_out, a_b))

    def debug_pbs_steps(self, small_prefix, steps, lut_ids=None):
        """The production PBS kernel stopped after `steps` CMUX steps; small_prefix: batch x (steps + 1)."""
        cts = np.ascontiguousarray(small_prefix, dtype=np.uint64).reshape(-1, steps + 1)
        ids = None if lut_ids is None else np.ascontiguousarray(lut_ids, dtype=np.uint32)
        out = np.empty((cts.shape[0], self.params.big_lwe_size), dtype=np.uint64)
        self._check(self.L.b200tfhe_debug_pbs_steps(self.h, _ptr(cts), _ptr(ids), _ptr(out), cts.shape[0], steps))
        return out

    def debug_from_torus(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).ravel()
        a = np.empty(x.size, dtype=np.uint64); b = np.empty(x.size, dtype=np.uint64)
        self._check(self.L.b200tfhe_debug_from_torus(self.h, x.ctypes.data_as(C.c_void_p), _ptr(a), _ptr(b), x.size))
        return a, b

    def debug_negacyclic_mul(self, a_int, b_torus, out):
        a = np.ascontiguousarray(a_int, dtype=np.uint64).reshape(-1, 2048)
        b = np.ascontiguousarray(b_torus, dtype=np.uint64).reshape(-1, 2048)
        assert out.dtype == np.uint64 and out.shape == a.shape and out.flags["C_CONTIGUOUS"]
        self._check(self.L.b200tfhe_debug_negacyclic_mul(self.h, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out


class BooleanEngine:
    """The reference's u32 boolean gate path (boolean::ServerKey: and / nand / or / nor / xor / xnor) on one GPU."""
    GATES = {"and": 0, "nand": 1, "or": 2, "nor": 3, "xor": 4, "xnor": 5}

    def __init__(self, params, keyswitch_first, device=0):
        self.L = load_library()
        self.params, self.keyswitch_first = params, bool(keyswitch_first)
        h = C.c_void_p()
        if self.L.b200tfhe_boolean_ctx_create(C.byref(params), 1 if keyswitch_first else 0, device, C.byref(h)) != 0:
            buf = C.create_string_buffer(1024)
            self.L.b200tfhe_last_global_error(buf, 1024)
            raise B200TfheError(buf.value.decode())
        self.h = h

    def _check(self, rc):
        if rc != 0:
            buf = C.create_string_buffer(1024)
            self.L.b200tfhe_boolean_last_error(self.h, buf, 1024)
            raise B200TfheError(buf.value.decode())

    def load_ksk(self, ksk):
        k = np.ascontiguousarray(ksk, dtype=np.uint32)
        self._check(self.L.b200tfhe_boolean_load_ksk(self.h, _ptr(k), k.size))

    def load_bsk_standard(self, bsk):
        b = np.ascontiguousarray(bsk, dtype=np.uint32)
        self._check(self.L.b200tfhe_boolean_load_bsk_standard(self.h, _ptr(b), b.size))

    def gate(self, gate, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        out = np.empty_like(a)
        g = self.GATES[gate] if isinstance(gate, str) else gate
        self._check(self.L.b200tfhe_boolean_gate_batch(self.h, g, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.L.b200tfhe_boolean_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parse_server_key(blob):
    """b200tfhe_parse_server_key: (Params, KeyView) of a serialised server key; needs no GPU."""
    L = load_library()
    buf = np.frombuffer(blob, dtype=np.uint8)
    p, v = Params(), KeyView()
    if L.b200tfhe_parse_server_key(buf.ctypes.data, buf.size, C.byref(p), C.byref(v)) != 0:
        msg = C.create_string_buffer(1024)
        L.b200tfhe_last_global_error(msg, 1024)
        raise B200TfheError(msg.value.decode())
    return p, v


class Program:
    """A batched call site (radix / FheString / Trivium operation) compiled for one workload shape:
    b200tfhe_program_* in include/b200tfhe.h; names and shapes in csrc/programs.hpp."""

    def __init__(self, engine, op, shape, _handle=None):
        self.engine = engine
        self.op, self.shape = op, list(shape)
        h = C.c_void_p()
        if _handle is not None:
            h = _handle
        else:
            sh = np.ascontiguousarray(shape, dtype=np.uint64)
            engine._check(engine.L.b200tfhe_program_create(engine.h, op.encode(), _ptr(sh), len(sh), C.byref(h)))
        self.h = h
        info = np.zeros(6, dtype=np.uint64)
        engine._check(engine.L.b200tfhe_program_info(self.h, _ptr(info)))
        self.info = dict(zip(["n_inputs", "n_outputs", "n_pbs", "depth", "n_stages", "n_luts"], (int(x) for x in info)))

    @classmethod
    def from_circuit(cls, engine, n_inputs, nodes, luts, outputs):
        """A caller-built schedule (b200tfhe_program_create_from_circuit).  nodes: list of
        (terms=[(block, coeff), ...], plaintext, lut_index or -1); luts: list of function tables."""
        tbeg, tb, tc, pt, nl = [0], [], [], [], []
        for terms, plain, lut in nodes:
            for b, c in terms:
                tb.append(b); tc.append(c)
            tbeg.append(len(tb)); pt.append(plain); nl.append(lut)
        keep = [np.ascontiguousarray(tbeg, dtype=np.uint32), np.ascontiguousarray(tb, dtype=np.int32),
                np.ascontiguousarray(tc, dtype=np.int64), np.ascontiguousarray(pt, dtype=np.uint64),
                np.ascontiguousarray(nl, dtype=np.int32), np.ascontiguousarray(luts, dtype=np.uint64).ravel(),
                np.ascontiguousarray(outputs, dtype=np.int32)]
        d = CircuitDesc(n_inputs, len(nodes), len(luts), len(outputs), *[k.ctypes.data for k in keep])
        h = C.c_void_p()
        engine._check(engine.L.b200tfhe_program_create_from_circuit(engine.h, C.byref(d), C.byref(h)))
        return cls(engine, "custom", [], _handle=h)

    def run(self, cts, out=None):
        """Host buffers (numpy or pinned torch): H2D, all levels, D2H; synchronous."""
        n_in, n_out, big = self.info["n_inputs"], self.info["n_outputs"], self.engine.params.big_lwe_size
        if isinstance(cts, np.ndarray):
            cts = np.ascontiguousarray(cts, dtype=np.uint64).reshape(n_in, big)
            if out is None:
                out = np.empty((n_out, big), dtype=np.uint64)
        self.engine._check(self.engine.L.b200tfhe_program_run(self.h, _ptr(cts), _ptr(out)))
        return out

    def run_device(self, d_in, d_out):
        self.engine._check(self.engine.L.b200tfhe_program_run_device(self.h, _ptr(d_in), _ptr(d_out)))

    def close(self):
        if getattr(self, "h", None) and getattr(self.engine, "h", None):
            self.engine.L.b200tfhe_program_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
