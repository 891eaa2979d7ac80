"""u32 boolean gate path: boolean::ServerKey gates on the GPU against the numpy oracle and the truth
tables, for an EncryptionKeyChoice::Small set (bootstrap then keyswitch) and a ::Big set (keyswitch then bootstrap)."""
import os
import re

import numpy as np
import pytest


def test_parameters_match_reference():
    from oracle import boolean_oracle as BO
    path = "/root/reference/tfhe/src/boolean/parameters/mod.rs"
    if not os.path.exists(path):
        pytest.skip("reference not mounted")
    src = open(path).read()
    for name, p in (("DEFAULT_PARAMETERS", BO.default_parameters()), ("DEFAULT_PARAMETERS_KS_PBS", BO.default_parameters_ks_pbs())):
        body = re.search(r"pub const " + name + r": BooleanParameters = BooleanParameters \{(.*?)\};", src, re.S).group(1)
        g = lambda key: re.search(key + r"\(([-0-9.e]+)\)", body).group(1)
        assert (int(g("LweDimension")), int(g("GlweDimension")), int(g("PolynomialSize"))) == (p.lwe_dimension, p.glwe_dimension, p.polynomial_size)
        assert float(g("lwe_modular_std_dev: StandardDev")) == p.lwe_std and float(g("glwe_modular_std_dev: StandardDev")) == p.glwe_std
        assert (int(g("pbs_base_log: DecompositionBaseLog")), int(g("pbs_level: DecompositionLevelCount"))) == (p.pbs_base_log, p.pbs_level)
        assert (int(g("ks_base_log: DecompositionBaseLog")), int(g("ks_level: DecompositionLevelCount"))) == (p.ks_base_log, p.ks_level)
        assert ("Big" in body) == p.ks_first


def test_oracle_gates_truth_tables():
    """The numpy oracle alone (insecure small parameters so that it runs in seconds): every gate, every input pair."""
    from oracle import boolean_oracle as BO
    for ks_first in (False, True):
        p = BO.BooleanParams(24, 2, 256, 2.0**-25, 2.0**-30, 6, 3, 3, 4, ks_first)
        keys = BO.BooleanKeyset(p, seed=5)
        a = np.array([0, 0, 1, 1] * 2); b = np.array([0, 1, 0, 1] * 2)
        ca, cb = keys.encrypt(a, seed=1), keys.encrypt(b, seed=2)
        assert list(keys.decrypt(ca)) == list(a)
        for name, g in BO.GATES.items():
            out = keys.gate(name, ca, cb)
            assert out.shape == ca.shape
            assert list(keys.decrypt(out)) == list(BO.TRUTH[g](a, b)), name


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["small_key_pbs_ks", "big_key_ks_pbs"])
def test_gpu_boolean_gates(which):
    import tfhe_rs_string_b200 as T
    from oracle import boolean_oracle as BO
    p = BO.default_parameters() if which == "small_key_pbs_ks" else BO.default_parameters_ks_pbs()
    keys = BO.BooleanKeyset(p, seed=77)
    eng = T.BooleanEngine(T.Params(p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.pbs_base_log, p.pbs_level,
                                   p.ks_base_log, p.ks_level, 2, 1), keyswitch_first=p.ks_first, device=0)
    try:
        eng.load_ksk(keys.ksk)
        eng.load_bsk_standard(keys.bsk_standard)
        rng = np.random.default_rng(3)
        B = 40
        a, b = rng.integers(0, 2, B), rng.integers(0, 2, B)
        ca, cb = keys.encrypt(a, seed=11), keys.encrypt(b, seed=12)
        for name, g in BO.GATES.items():
            got = eng.gate(name, ca, cb)
            assert got.shape == (B, p.ct_size)
            assert list(keys.decrypt(got)) == list(BO.TRUTH[g](a, b)), name
        # against the oracle on identical inputs: same decrypted bits, phases agree to well inside the 1/8 margin
        ref = keys.gate("nand", ca[:8], cb[:8])
        got = eng.gate("nand", ca[:8], cb[:8])
        assert list(keys.decrypt(got)) == list(keys.decrypt(ref))
        d = (keys.phase(got) - keys.phase(ref)).astype(np.int32)
        assert np.abs(d).max() < (1 << 26), np.abs(d).max()
        # a chain of gates: outputs are valid inputs (full adder on 1-bit values)
        cc = keys.encrypt(rng.integers(0, 2, B), seed=13)
        c = keys.decrypt(cc)
        s1 = eng.gate("xor", ca, cb)
        total = eng.gate("xor", s1, cc)
        carry = eng.gate("or", eng.gate("and", ca, cb), eng.gate("and", s1, cc))
        assert list(keys.decrypt(total)) == list((a + b + c) % 2)
        assert list(keys.decrypt(carry)) == list((a + b + c) // 2)
        with pytest.raises(T.B200TfheError):
            eng.L.b200tfhe_boolean_gate_batch.restype
            eng._check(eng.L.b200tfhe_boolean_gate_batch(eng.h, 9, ca.ctypes.data, cb.ctypes.data, ca.ctypes.data, 1))
    finally:
        eng.close()
