"""Every classic parameter set of the reference with N = 2048 (shortint/parameters/mod.rs: 1_3, 2_2, 3_1, 4_0
KS_PBS at :688-747 and 2_2 PBS_KS at :1155-1169) shares k = 1, one PBS level of base 2^23; they differ in the
small dimension, the keyswitch decomposition and the message / carry split.  The kernels are generic in
those, so the sets other than the headline 2_2 are covered here (2_2 PBS_KS: tests/test_pbs_ks_order.py)."""
import numpy as np
import pytest

SETS = {
    # name: (lwe_dimension, lwe_std, ks_base_log, ks_level, message_modulus, carry_modulus)
    "PARAM_MESSAGE_1_CARRY_3_KS_PBS": (745, 0.000006692125069956277, 3, 5, 2, 8),
    "PARAM_MESSAGE_3_CARRY_1_KS_PBS": (742, 0.000007069849454709433, 3, 5, 8, 2),
    "PARAM_MESSAGE_4_CARRY_0_KS_PBS": (742, 0.000007069849454709433, 3, 5, 16, 1),
}


def make_params(O, name):
    n, std, kb, kl, mm, cm = SETS[name]
    p = O.params_message_2_carry_2()
    p.lwe_dimension, p.lwe_modular_std_dev = n, std
    p.ks_base_log, p.ks_level, p.message_modulus, p.carry_modulus = kb, kl, mm, cm
    return p


def test_reference_values_match_parameter_file():
    """The table above restates the reference's constants; check it against the source when it is there."""
    import os, re
    path = "/root/reference/tfhe/src/shortint/parameters/mod.rs"
    if not os.path.exists(path):
        pytest.skip("reference not mounted")
    src = open(path).read()
    for name, (n, std, kb, kl, mm, cm) in SETS.items():
        body = re.search(r"pub const " + name + r": ClassicPBSParameters = ClassicPBSParameters \{(.*?)\};", src, re.S).group(1)
        g = lambda k: re.search(k + r"\(([-0-9.e]+)\)", body).group(1)
        assert int(g("LweDimension")) == n and int(g("PolynomialSize")) == 2048 and int(g("GlweDimension")) == 1
        assert float(g("lwe_modular_std_dev: StandardDev")) == std
        assert (int(g("pbs_base_log: DecompositionBaseLog")), int(g("pbs_level: DecompositionLevelCount"))) == (23, 1)
        assert (int(g("ks_base_log: DecompositionBaseLog")), int(g("ks_level: DecompositionLevelCount"))) == (kb, kl)
        assert (int(g("MessageModulus")), int(g("CarryModulus"))) == (mm, cm)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(SETS))
def test_gpu_matches_oracle_on_other_parameter_sets(oracle_mod, name):
    import tfhe_rs_string_b200 as T
    O = oracle_mod
    p = make_params(O, name)
    keys = O.Keyset(p, seed=0xB300 + p.lwe_dimension + p.message_modulus)
    eng = T.Engine(T.Params(p.lwe_dimension, 1, 2048, 23, 1, p.ks_base_log, p.ks_level, p.message_modulus, p.carry_modulus), device=0)
    try:
        eng.load_ksk(keys.ksk)
        eng.load_bsk_standard(keys.bsk_standard)
        msgs = np.arange(64) % 16
        cts = keys.encrypt_batch(msgs, seed=31)
        f = lambda x: (11 * x + 2) % 16
        assert np.array_equal(eng.keyswitch_batch(cts), keys.keyswitch_batch(cts))
        got = eng.ks_pbs_batch(cts, np.full(64, eng.generate_lookup_table(f), dtype=np.uint32))
        ref = keys.ks_pbs_batch(cts, keys.lut(f))
        assert list(keys.decrypt_batch(got)) == [f(int(m)) for m in msgs]
        assert np.array_equal(keys.decrypt_batch(got), keys.decrypt_batch(ref))
        dphase = (keys.phase_batch(got) - keys.phase_batch(ref)).astype(np.int64)
        assert np.abs(dphase).max() < (1 << 53)
    finally:
        eng.close()
