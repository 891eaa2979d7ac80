"""PBSOrder::BootstrapKeyswitch (shortint/server_key/mod.rs:859-933) at PARAM_MESSAGE_2_CARRY_2_PBS_KS
(shortint/parameters/mod.rs:1155-1169): the oracle on the CPU, and the GPU entry point against it."""
import numpy as np
import pytest

U64 = np.uint64


def decode(phase, delta):
    return ((np.asarray(phase, dtype=U64) + U64(delta // 2)) // U64(delta)) % U64(16)


@pytest.fixture(scope="module")
def pbs_ks_keys(oracle_mod):
    return oracle_mod.Keyset(oracle_mod.params_message_2_carry_2_pbs_ks(), seed=0xB201)


def small_inputs(keys, msgs, seed):
    """Small-key encryptions of msgs: a big-key encryption keyswitched by the oracle (what a PBS->KS
    ciphertext is between two operations)."""
    return keys.keyswitch_batch(keys.encrypt_batch(msgs, seed=seed))


def test_oracle_pbs_then_ks_decrypts(pbs_ks_keys):
    keys = pbs_ks_keys
    p = keys.params
    assert (p.lwe_dimension, p.ks_base_log, p.ks_level) == (870, 4, 4)
    msgs = np.arange(16)
    f = lambda x: (3 * x + 5) % 16
    small = small_inputs(keys, msgs, 11)
    out = keys.keyswitch_batch(keys.bootstrap_batch(small, keys.lut(f)))
    assert out.shape == (16, 871)
    delta = (1 << 63) // 16
    assert list(decode(keys.small_phase_batch(out), delta)) == [f(int(m)) for m in msgs]


@pytest.mark.gpu
def test_gpu_pbs_ks_matches_oracle(pbs_ks_keys):
    import tfhe_rs_string_b200 as T
    keys = pbs_ks_keys
    eng = T.Engine(T.Params.message_2_carry_2_pbs_ks(), device=0)
    try:
        eng.load_ksk(keys.ksk)
        eng.load_bsk_standard(keys.bsk_standard)
        rng = np.random.default_rng(5)
        msgs = rng.integers(0, 16, 200)
        fs = [lambda x: x, lambda x: (x * x) % 16]
        ids = np.array([eng.generate_lookup_table(f) for f in fs], dtype=np.uint32)
        idx = (np.arange(200) % 2).astype(np.uint32)
        small = small_inputs(keys, msgs, 12)
        got = eng.pbs_ks_batch(small, ids[idx])
        # keyswitch alone stays bit-exact at these parameters (base 2^4, 4 levels)
        big = keys.bootstrap_batch(small, np.stack([keys.lut(f) for f in fs]), idx)
        assert np.array_equal(eng.keyswitch_batch(big), keys.keyswitch_batch(big))
        ref = keys.keyswitch_batch(big)
        delta = (1 << 63) // 16
        exp = [fs[i](int(m)) for m, i in zip(msgs, idx)]
        assert list(decode(keys.small_phase_batch(got), delta)) == exp
        assert list(decode(keys.small_phase_batch(ref), delta)) == exp
        # After the keyswitch the two results carry independent decomposition noise (the digits of two
        # big ciphertexts that differ by FFT rounding differ), so the bound is on the noise itself: same
        # spread as the oracle's, far below delta / 2 = 2^58.
        ideal = np.array(exp, dtype=U64) * U64(delta)
        e_gpu = (keys.small_phase_batch(got) - ideal).astype(np.int64).astype(np.float64)
        e_ref = (keys.small_phase_batch(ref) - ideal).astype(np.int64).astype(np.float64)
        assert 0.6 < e_gpu.std() / e_ref.std() < 1.6, (e_gpu.std(), e_ref.std())
        assert np.abs(e_gpu).max() < (1 << 56) and np.abs(e_ref).max() < (1 << 56)
    finally:
        eng.close()
