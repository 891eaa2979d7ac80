"""Classic parameter sets outside the specialised kernels' shape: PARAM_MESSAGE_1_CARRY_1_KS_PBS
(k = 3, N = 512, shortint/parameters/mod.rs:613-627) and PARAM_MESSAGE_3_CARRY_3_KS_PBS (N = 8192, two PBS levels,
:853-867) run on csrc/pbs_generic.cuh (multi-level external product, ggsw.rs:524-572, is_output_uninit :652-676).
Same three checks as the headline set: keyswitch bit-exact, decrypt exact, phase close to the oracle's."""
import re
import os

import numpy as np
import pytest

U64 = np.uint64
SETS = {
    # name: (n, k, N, lwe_std, glwe_std, pbs_base_log, pbs_level, ks_base_log, ks_level, msg, carry)
    "PARAM_MESSAGE_1_CARRY_1_KS_PBS": (684, 3, 512, 0.00002043784477291318, 0.0000000000034525330484572114, 18, 1, 4, 3, 2, 2),
    "PARAM_MESSAGE_3_CARRY_3_KS_PBS": (864, 1, 8192, 0.000000757998020150446, 0.0000000000000000002168404344971009, 15, 2, 3, 6, 8, 8),
}


def test_reference_values_match_parameter_file(oracle_mod):
    path = "/root/reference/tfhe/src/shortint/parameters/mod.rs"
    if not os.path.exists(path):
        pytest.skip("reference not mounted")
    src = open(path).read()
    for name, (n, k, N, ls, gs, pb, pl, kb, kl, mm, cm) in SETS.items():
        body = re.search(r"pub const " + name + r": ClassicPBSParameters = ClassicPBSParameters \{(.*?)\};", src, re.S).group(1)
        g = lambda key: re.search(key + r"\(([-0-9.e]+)\)", body).group(1)
        assert (int(g("LweDimension")), int(g("GlweDimension")), int(g("PolynomialSize"))) == (n, k, N)
        assert float(g("lwe_modular_std_dev: StandardDev")) == ls and float(g("glwe_modular_std_dev: StandardDev")) == gs
        assert (int(g("pbs_base_log: DecompositionBaseLog")), int(g("pbs_level: DecompositionLevelCount"))) == (pb, pl)
        assert (int(g("ks_base_log: DecompositionBaseLog")), int(g("ks_level: DecompositionLevelCount"))) == (kb, kl)
        assert (int(g("MessageModulus")), int(g("CarryModulus"))) == (mm, cm)
    O = oracle_mod
    for name, mk in (("PARAM_MESSAGE_1_CARRY_1_KS_PBS", O.params_message_1_carry_1), ("PARAM_MESSAGE_3_CARRY_3_KS_PBS", O.params_message_3_carry_3)):
        p = mk()
        assert (p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.lwe_modular_std_dev, p.glwe_modular_std_dev, p.pbs_base_log,
                p.pbs_level, p.ks_base_log, p.ks_level, p.message_modulus, p.carry_modulus) == SETS[name]


def test_oracle_multi_level_and_k3(oracle_mod):
    """The oracle itself on a k = 3 set and on a two-level toy set: decrypts f(m) for every message."""
    O = oracle_mod
    p = O.params_message_1_carry_1()
    keys = O.Keyset(p, seed=5)
    msgs = np.arange(8) % 4
    f = lambda x: (x + 1) % 4
    assert list(keys.decrypt_batch(keys.ks_pbs_batch(keys.encrypt_batch(msgs, seed=1), keys.lut(f)))) == [f(int(m)) for m in msgs]
    q = O.params_toy(8, 512)
    q.pbs_base_log, q.pbs_level = 12, 2
    k2 = O.Keyset(q, seed=6)
    msgs = np.arange(16)
    g = lambda x: (3 * x + 2) % 16
    assert list(k2.decrypt_batch(k2.ks_pbs_batch(k2.encrypt_batch(msgs, seed=2), k2.lut(g)))) == [g(int(m)) for m in msgs]


@pytest.mark.gpu
@pytest.mark.parametrize("name,batch", [("PARAM_MESSAGE_1_CARRY_1_KS_PBS", 200), ("PARAM_MESSAGE_3_CARRY_3_KS_PBS", 24)])
def test_gpu_generic_kernel_matches_oracle(oracle_mod, name, batch):
    import time
    import tfhe_rs_string_b200 as T
    O = oracle_mod
    p = (O.params_message_1_carry_1 if "1_CARRY_1" in name else O.params_message_3_carry_3)()
    keys = O.Keyset(p, seed=0xB400 + p.polynomial_size)
    ms = p.message_modulus * p.carry_modulus
    eng = T.Engine(T.Params(p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.pbs_base_log, p.pbs_level,
                            p.ks_base_log, p.ks_level, p.message_modulus, p.carry_modulus), device=0)
    try:
        eng.load_ksk(keys.ksk)
        eng.load_bsk_standard(keys.bsk_standard)
        msgs = np.arange(batch) % (2 * ms)          # includes the padding bit
        cts = keys.encrypt_batch(msgs, seed=77)
        f = lambda x: (5 * x + 3) % ms
        assert np.array_equal(eng.keyswitch_batch(cts), keys.keyswitch_batch(cts))
        ids = np.full(batch, eng.generate_lookup_table(f), dtype=np.uint32)
        got = eng.ks_pbs_batch(cts, ids)
        t0 = time.perf_counter()
        got = eng.ks_pbs_batch(cts, ids)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        eng.ks_pbs_batch(cts[:1], ids[:1])
        dt1 = time.perf_counter() - t0
        print(f"{name}: {batch} KS+PBS in {dt * 1e3:.1f} ms, one KS+PBS in {dt1 * 1e3:.2f} ms on the generic kernel "
              f"(reference CPU, 1 thread: {'7.28' if p.polynomial_size == 512 else '121'} ms, docs/getting_started/benchmarks.md:42)")
        ref = keys.ks_pbs_batch(cts, keys.lut(f))
        exp = [(f(int(m)) if m < ms else (2 * ms - f(int(m) - ms)) % (2 * ms)) for m in msgs]
        assert list(keys.decrypt_batch(got)) == exp
        assert np.array_equal(keys.decrypt_batch(got), keys.decrypt_batch(ref))
        dphase = (keys.phase_batch(got) - keys.phase_batch(ref)).astype(np.int64)
        assert np.abs(dphase).max() < p.delta // 16, np.log2(float(np.abs(dphase).max()))
        ideal = np.array(exp, dtype=U64) * U64(p.delta)
        e_gpu = (keys.phase_batch(got) - ideal).astype(np.int64).astype(np.float64)
        e_ref = (keys.phase_batch(ref) - ideal).astype(np.int64).astype(np.float64)
        assert 0.5 < e_gpu.std() / e_ref.std() < 2.0, (e_gpu.std(), e_ref.std())
        # a batched program on this parameter set: the level executor is parameter independent
        if ms == 4:
            prog = T.Program(eng, "radix_bitxor", [8, 4])
            a, b = np.arange(8) * 13 % 16, np.arange(8) * 7 % 16
            blk = lambda v: np.stack([(v >> k) & 1 for k in range(4)], axis=-1).ravel()
            out = keys.decrypt_batch(prog.run(keys.encrypt_batch(np.concatenate([blk(a), blk(b)]), seed=5)))
            assert np.array_equal(out.reshape(8, 4), np.stack([((a ^ b) >> k) & 1 for k in range(4)], axis=-1))
            prog.close()
    finally:
        eng.close()
