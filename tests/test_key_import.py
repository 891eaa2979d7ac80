"""Key wire-format import: bincode(shortint::ServerKey) and the shim's standard-domain bundle.
PARITY UNPINNED (no Rust toolchain in the image): the fixtures are built by tests/golden/make_server_key_fixture.py
from the serde field order of the reference's structs, independently of the C++ parser."""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
U64 = np.uint64


def _fixture_module():
    spec = importlib.util.spec_from_file_location("make_server_key_fixture", os.path.join(HERE, "golden", "make_server_key_fixture.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("kind", ["fourier", "std"])
def test_parse_committed_fixture(oracle_mod, kind):
    import tfhe_rs_string_b200 as T
    blob = open(os.path.join(HERE, "golden", f"server_key_toy_{kind}.bin"), "rb").read()
    p, v = T.parse_server_key(blob)
    assert (p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.pbs_base_log, p.pbs_level, p.ks_base_log, p.ks_level,
            p.message_modulus, p.carry_modulus) == (4, 1, 256, 23, 1, 3, 5, 4, 4)
    assert v.bsk_is_fourier == (1 if kind == "fourier" else 0) and v.pbs_order == 0
    keys = oracle_mod.Keyset(oracle_mod.params_toy(4, 256), seed=11)   # the seed the fixture was built with
    raw = np.frombuffer(blob, dtype=np.uint8)
    ksk = raw[v.ksk_offset:v.ksk_offset + 8 * v.ksk_len].view(U64)
    assert v.ksk_len == keys.ksk.size and np.array_equal(ksk, keys.ksk)
    if kind == "std":
        bsk = raw[v.bsk_offset:v.bsk_offset + 8 * v.bsk_len].view(U64)
        assert np.array_equal(bsk, keys.bsk_standard)
    else:
        assert v.bsk_len == keys.bsk_standard.size // 2 and v.bsk_poly_stride_bytes == 8 + 128 * 16
        assert (v.max_degree, v.max_noise_level) == (15, 5)
        first = raw[v.bsk_offset:v.bsk_offset + 128 * 16].view(np.float64).reshape(128, 2)
        ref = _fixture_module().fourier_natural_order(keys.params, keys.bsk_standard)[0]
        assert np.allclose(first[:, 0] + 1j * first[:, 1], ref)


def test_parser_rejects_malformed_input():
    import tfhe_rs_string_b200 as T
    blob = open(os.path.join(HERE, "golden", "server_key_toy_std.bin"), "rb").read()
    for bad in (blob[:-1], blob + b"\0", blob[:100], b""):
        with pytest.raises(T.B200TfheError):
            T.parse_server_key(bad)
    # non-native ciphertext modulus in the keyswitch key header
    p, v = T.parse_server_key(blob)
    off = v.ksk_offset + 8 * v.ksk_len + 24
    with pytest.raises(T.B200TfheError, match="modulus"):
        T.parse_server_key(blob[:off] + b"\x01" + blob[off + 1:])


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["fourier", "std"])
def test_imported_key_bootstraps_correctly(real_keys, kind):
    """Full-size PARAM_MESSAGE_2_CARRY_2 key through the wire format, then KS+PBS must decrypt exactly.  For the Fourier
    layout this also checks that this library's Fourier key order is the natural frequency order."""
    import tfhe_rs_string_b200 as T
    m = _fixture_module()
    p = real_keys.params
    blob = (m.serialize_server_key_fourier if kind == "fourier" else m.serialize_std_bundle)(p, real_keys.ksk, real_keys.bsk_standard)
    pp, v = T.parse_server_key(blob)
    eng = T.Engine(pp, device=0)
    try:
        eng.load_server_key_bytes(blob)
        msgs = np.arange(96) % 16
        cts = real_keys.encrypt_batch(msgs, seed=4242)
        f = lambda x: (7 * x + 5) % 16
        out = eng.ks_pbs_batch(cts, np.full(96, eng.generate_lookup_table(f), dtype=np.uint32))
        assert list(real_keys.decrypt_batch(out)) == [f(int(x)) for x in msgs]
        assert np.array_equal(eng.keyswitch_batch(cts), real_keys.keyswitch_batch(cts))
        ph = real_keys.phase_batch(out)
        err = (ph - np.array([f(int(x)) for x in msgs], dtype=U64) * U64(p.delta)).astype(np.int64)
        assert np.abs(err).max() < (1 << 54)
        with pytest.raises(T.B200TfheError, match="differ"):
            other = T.Engine(T.Params.message_2_carry_2_pbs_ks(), device=0)
            try:
                other.load_server_key_bytes(blob)
            finally:
                other.close()
    finally:
        eng.close()
