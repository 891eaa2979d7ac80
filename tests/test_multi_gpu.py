"""Multi-GPU scheduler behind the C ABI (b200tfhe_ctx_create_multi): contiguous shards, one per GPU, keys copied GPU
to GPU, programs split over their independent units.  Runs on however many GPUs the box has; the sharding logic is
also exercised with a one-GPU "multi" context."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
U64 = np.uint64


def _devices():
    import torch
    return list(range(min(torch.cuda.device_count(), 8)))


def _engine(keys, devices):
    import tfhe_rs_string_b200 as T
    p = keys.params
    e = T.Engine(T.Params(p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.pbs_base_log, p.pbs_level,
                          p.ks_base_log, p.ks_level, p.message_modulus, p.carry_modulus), devices=devices)
    e.load_ksk(keys.ksk)
    e.load_bsk_standard(keys.bsk_standard)
    return e


def test_multi_context_matches_single(engine, real_keys):
    devs = _devices()
    multi = _engine(real_keys, devs)
    try:
        for B in (1, 7, 640, 1301):
            msgs = np.arange(B) % 16
            cts = real_keys.encrypt_batch(msgs, seed=50 + B)
            f = lambda x: (x * 3 + 2) % 16
            ids_m = np.full(B, multi.generate_lookup_table(f), dtype=np.uint32)
            out = multi.ks_pbs_batch(cts, ids_m)
            assert list(real_keys.decrypt_batch(out)) == [f(int(m)) for m in msgs]
            assert np.array_equal(multi.keyswitch_batch(cts), real_keys.keyswitch_batch(cts))
            if len(devs) == 1:   # same kernels, same shard: bit-identical to the single-GPU context
                assert np.array_equal(out, engine.ks_pbs_batch(cts, np.full(B, engine.generate_lookup_table(f), dtype=np.uint32)))
    finally:
        multi.close()


def test_multi_context_programs_split_over_units(real_keys):
    import tfhe_rs_string_b200 as T
    devs = _devices()
    multi = _engine(real_keys, devs)
    try:
        rng = np.random.default_rng(8)
        n = 37   # not a multiple of the GPU count: ragged shards
        a, b = rng.integers(0, 256, n), rng.integers(0, 256, n)
        b[::3] = a[::3]
        blk = lambda v: np.stack([(np.asarray(v, dtype=U64) >> U64(2 * k)) & U64(3) for k in range(4)], axis=-1).ravel()
        cts = real_keys.encrypt_batch(np.concatenate([blk(a), blk(b)]), seed=91)
        prog = T.Program(multi, "radix_eq", [n, 4])
        assert prog.info["n_pbs"] == 3 * n
        assert np.array_equal(real_keys.decrypt_batch(prog.run(cts)), (a == b).astype(U64))
        prog.close()
        prog = T.Program(multi, "radix_add", [n, 4])
        out = real_keys.decrypt_batch(prog.run(cts)).reshape(n, 4)
        assert np.array_equal(sum(out[:, k] << U64(2 * k) for k in range(4)), ((a + b) % 256).astype(U64))
        prog.close()
    finally:
        multi.close()


def test_two_gpu_c_client():
    """examples/multi_gpu_example.c drives every GPU of the box (at least 2 when present) through
    b200tfhe_ctx_create_multi from plain C."""
    import tfhe_rs_string_b200 as T
    inc, libdir = os.path.join(ROOT, "include"), os.path.dirname(T.lib_path())
    n = max(1, min(len(_devices()), 2))
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "multi_gpu_example")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-I", inc, os.path.join(ROOT, "examples", "multi_gpu_example.c"),
                               "-L", libdir, "-lb200tfhe", "-Wl,-rpath," + libdir, "-o", exe])
        r = subprocess.run([exe, str(n)], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert f"{n} GPU" in r.stdout, r.stdout


def test_single_gpu_c_client_runs_on_hardware():
    """The GPU copy of tests/test_abi.py::test_header_is_plain_c_and_example_links: the C client must succeed here."""
    import tfhe_rs_string_b200 as T
    inc, libdir = os.path.join(ROOT, "include"), os.path.dirname(T.lib_path())
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "ks_pbs_example")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-I", inc, os.path.join(ROOT, "examples", "ks_pbs_example.c"),
                               "-L", libdir, "-lb200tfhe", "-Wl,-rpath," + libdir, "-o", exe])
        r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        assert "bootstrapped 8 ciphertexts" in r.stdout
