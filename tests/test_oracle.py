"""CPU tests: pin the oracle against the reference's own vectors / properties."""
import ctypes as C

import numpy as np
import pytest

U64 = np.uint64


def test_monomial_doc_vectors(oracle_mod):
    # core_crypto/algorithms/polynomial_algorithms.rs:313 (div) and :373 (mul), u8 examples
    L = oracle_mod.lib()
    inp = np.array([1, 2, 3], dtype=U64)
    out = np.zeros(3, dtype=U64)
    L.orc_monomial_div(out, inp, 3, 2)
    assert list(out & U64(0xFF)) == [3, 255, 254]
    L.orc_monomial_mul(out, inp, 3, 2)
    assert list(out & U64(0xFF)) == [254, 253, 1]


def test_monomial_mul_div_roundtrip_and_subtract(oracle_mod):
    L = oracle_mod.lib()
    rng = np.random.default_rng(0)
    N = 64
    p = rng.integers(0, 2**64, N, dtype=U64)
    for deg in [0, 1, 17, N - 1, N, N + 5, 2 * N - 1, 2 * N]:
        a = np.zeros(N, dtype=U64); b = np.zeros(N, dtype=U64); c = np.zeros(N, dtype=U64)
        L.orc_monomial_mul(a, p, N, deg)
        L.orc_monomial_div(b, a, N, deg)
        assert np.array_equal(b, p)
        L.orc_monomial_mul_and_subtract(c, p, N, deg)
        assert np.array_equal(c, a - p)
    # X^N = -1
    a = np.zeros(N, dtype=U64)
    L.orc_monomial_mul(a, p, N, N)
    assert np.array_equal(a, U64(0) - p)


def test_decomposer_properties(oracle_mod):
    # commons/math/decomposition/tests.rs:32-136: digit range, recompose == closest, idempotence
    L = oracle_mod.lib()
    rng = np.random.default_rng(1)
    for base_log, level in [(3, 5), (23, 1), (4, 3), (8, 2), (1, 1), (15, 2), (7, 9)]:
        for x in rng.integers(0, 2**64, 2000, dtype=U64):
            x = int(x)
            d = np.zeros(level, dtype=np.int64)
            L.orc_decompose(x, base_log, level, d)
            half = 1 << (base_log - 1)
            assert all(-half <= int(t) <= half for t in d)
            # recompose: digit of level l (d[level - l]) weighs 2^(64 - base_log*l)
            rec = sum(int(d[level - l]) << (64 - base_log * l) for l in range(1, level + 1)) % (1 << 64)
            closest = L.orc_closest_representable(x, base_log, level)
            assert rec == closest
            assert L.orc_closest_representable(closest, base_log, level) == closest
            eps = (1 << (64 - base_log * level - 1)) // 2
            assert L.orc_closest_representable((closest + eps) % (1 << 64), base_log, level) == closest
            assert L.orc_closest_representable((closest - eps) % (1 << 64), base_log, level) == closest


def test_modulus_switch(oracle_mod):
    # fft_impl/common.rs:26-43; may return 2N
    L = oracle_mod.lib()
    assert L.orc_modulus_switch(0, 11) == 0
    assert L.orc_modulus_switch(2**64 - 1, 11) == 4096
    assert L.orc_modulus_switch(1 << 52, 11) == 1
    assert L.orc_modulus_switch((1 << 51), 11) == 1      # exactly half rounds up
    assert L.orc_modulus_switch((1 << 51) - 1, 11) == 0
    rng = np.random.default_rng(2)
    for x in rng.integers(0, 2**64, 1000, dtype=U64):
        x = int(x)
        assert L.orc_modulus_switch(x, 11) == (x * 4096 + (1 << 63)) >> 64


def test_sample_extract(oracle_mod):
    L = oracle_mod.lib()
    rng = np.random.default_rng(3)
    N, k = 16, 2
    glwe = rng.integers(0, 2**64, (k + 1) * N, dtype=U64)
    sk = rng.integers(0, 2, k * N, dtype=U64)
    lwe = np.zeros(k * N + 1, dtype=U64)
    L.orc_sample_extract0(lwe, glwe, k, N)
    # constant coefficient of body - sum_p A_p * S_p  ==  lwe.body - <lwe.mask, sk>
    const = int(glwe[k * N])
    for p in range(k):
        a = glwe[p * N:(p + 1) * N].astype(object); s = sk[p * N:(p + 1) * N].astype(object)
        const -= int(a[0]) * int(s[0])
        for j in range(1, N):
            const += int(a[N - j]) * int(s[j])   # X^N = -1
    got = int(lwe[k * N]) - sum(int(lwe[i]) * int(sk[i]) for i in range(k * N))
    assert const % 2**64 == got % 2**64


def _negacyclic_exact(a, b):
    n = len(a)
    full = np.convolve(np.array(a, dtype=object), np.array(b, dtype=object))
    res = full[:n].copy()
    res[:n - 1] -= full[n:]
    return res


@pytest.mark.parametrize("N", [256, 2048])
def test_fft_roundtrip_tolerance(oracle_mod, N):
    # fft/tests.rs:9-80: forward_as_torus then backward: distance < 2^(64-50)
    L = oracle_mod.lib()
    rng = np.random.default_rng(4)
    poly = rng.integers(0, 2**64, N, dtype=U64)
    f = np.zeros(N, dtype=np.float64)
    L.orc_fft_forward_torus(f, poly, N)
    back = np.zeros(N, dtype=U64)
    L.orc_fft_add_backward_torus(back, f, N)
    dist = np.abs((back - poly).astype(np.int64))
    assert dist.max() < 2**14


@pytest.mark.parametrize("N", [256, 2048])
def test_fft_product_tolerance(oracle_mod, N):
    # fft/tests.rs:82-222: integer poly (16-bit) times torus poly vs schoolbook,
    # tolerance 2^(64 - (52 - 16 - log2 N))
    L = oracle_mod.lib()
    rng = np.random.default_rng(5)
    a = rng.integers(-2**15, 2**15, N)
    b = rng.integers(0, 2**64, N, dtype=U64)
    fa = np.zeros(N); fb = np.zeros(N)
    L.orc_fft_forward_integer(fa, a.astype(np.int64).view(U64).copy(), N)
    L.orc_fft_forward_torus(fb, b, N)
    za = fa[0::2] + 1j * fa[1::2]; zb = fb[0::2] + 1j * fb[1::2]
    prod = za * zb
    fp = np.empty(N); fp[0::2] = prod.real; fp[1::2] = prod.imag
    got = np.zeros(N, dtype=U64)
    L.orc_fft_add_backward_torus(got, fp, N)
    ref = _negacyclic_exact(a.astype(object), b.astype(object))
    ref = np.array([int(x) % 2**64 for x in ref], dtype=U64)
    dist = np.abs((got - ref).astype(np.int64)).max()
    log2N = int(np.log2(N))
    assert dist <= 2 ** (64 - (52 - 16 - log2N))


def test_from_torus(oracle_mod):
    L = oracle_mod.lib()
    assert L.orc_from_torus(0.0) == 0
    assert L.orc_from_torus(0.25) == 1 << 62
    assert L.orc_from_torus(-0.25) == (1 << 64) - (1 << 62)
    assert L.orc_from_torus(3.25) == 1 << 62
    assert L.orc_from_torus(2.0 ** -64) == 1


def test_lut_layout(toy_keys, real_keys):
    # shortint/engine/mod.rs:72-128: boxes of N/16, first half box negated and rotated to the end
    for ks in (toy_keys, real_keys):
        p = ks.params
        N, box, delta = p.polynomial_size, p.polynomial_size // 16, p.delta
        lut = ks.lut(lambda x: (3 * x + 1) % 16)
        assert not lut[:N].any()
        body = lut[N:]
        for i in range(16):
            f = (3 * i + 1) % 16
            lo, hi = i * box - box // 2, i * box + box // 2
            if i == 0:
                assert (body[:hi] == U64(f * delta)).all()
                assert (body[N - box // 2:] == U64((-f * delta) % 2**64)).all()
            else:
                assert (body[lo:hi] == U64(f * delta)).all()


def test_keyswitch_decrypts(real_keys):
    # algorithms/test/lwe_keyswitch.rs:8-109: encrypt -> KS -> decrypt under the small key
    msgs = np.arange(16)
    cts = real_keys.encrypt_batch(msgs, seed=11)
    small = real_keys.keyswitch_batch(cts)
    ph = real_keys.small_phase_batch(small)
    delta = real_keys.params.delta
    dec = ((ph.astype(object) + delta // 2) // delta) % 32
    assert list(dec) == list(msgs)


@pytest.mark.parametrize("which", ["toy", "real"])
def test_pbs_identity_and_padding_bit(which, toy_keys, real_keys):
    # shortint/server_key/tests/shortint.rs:366-462 (keyswitch_bootstrap / programmable bootstrap)
    # + negacyclic behaviour the comparator relies on (integer/server_key/comparator.rs:198-217)
    ks = toy_keys if which == "toy" else real_keys
    msgs = np.arange(32) if which == "toy" else np.array([0, 1, 5, 15, 16, 17, 23, 31])
    cts = ks.encrypt_batch(msgs, seed=12)
    lut = ks.lut(lambda x: x)
    out = ks.ks_pbs_batch(cts, lut)
    dec = ks.decrypt_batch(out)
    exp = [m if m < 16 else (32 - (m - 16)) % 32 for m in msgs]
    assert list(dec) == exp


def test_trivial_pbs_matches_encrypted(toy_keys, oracle_mod):
    # shortint/server_key/tests/shortint.rs:3233-3296 (shortint_trivial_pbs), incl. dirty padding bit
    ks, L, p = toy_keys, oracle_mod.lib(), toy_keys.params
    lut = ks.lut(lambda x: (x * x + 3) % 16)
    msgs = np.arange(32)
    cts = ks.encrypt_batch(msgs, seed=13)
    dec = ks.decrypt_batch(ks.ks_pbs_batch(cts, lut))
    for m in msgs:
        body = (int(m) * p.delta) % 2**64
        triv = L.orc_trivial_pbs(C.byref(p), body, lut)
        v = ((triv + p.delta // 2) // p.delta) % 32
        assert v == dec[m]


def test_many_luts_batch(toy_keys):
    ks = toy_keys
    luts = np.stack([ks.lut(lambda x: x % 4), ks.lut(lambda x: x // 4), ks.lut(lambda x: int(x == 5))])
    msgs = np.arange(48) % 16
    idx = (np.arange(48) // 16).astype(np.uint32)
    out = ks.ks_pbs_batch(ks.encrypt_batch(msgs, seed=14), luts, idx)
    dec = ks.decrypt_batch(out)
    exp = [[m % 4, m // 4, int(m == 5)][i] for m, i in zip(msgs, idx)]
    assert list(dec) == exp


def test_lat4_fft_data_flow_model():
    """The data flow of pbs_lat4_kernel (four warps per polynomial: two cross-warp radix-2 levels + 8-point transforms,
    tools/proto_fft8x4.py) reproduces the negacyclic FFT-1024 in the frequency layout of fft.cuh, and its inverse is the inverse."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "proto_fft8x4.py")
    spec = importlib.util.spec_from_file_location("proto_fft8x4", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fwd_err, rt_err = mod.run(seed=3)
    assert fwd_err < 1e-12
    assert rt_err < 1e-6     # values up to 2^22, unnormalised transforms of 1024 points
