"""GPU tests of the batched call sites (BASELINE.json configs 2-5) through the C ABI
(b200tfhe_program_*): decrypted results must be bit-exact vs the clear operation / known answers."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
U64 = np.uint64


def blocks_of(values, nb=4):
    v = np.asarray(values, dtype=U64)
    return np.stack([(v >> U64(2 * k)) & U64(3) for k in range(nb)], axis=-1)


def from_blocks(b):
    b = np.asarray(b, dtype=U64)
    return sum(b[..., k] << U64(2 * k) for k in range(b.shape[-1]))


def chars(strings):
    return blocks_of(np.array([[ord(ch) for ch in s] for s in strings], dtype=U64))


def run(engine, keys, op, shape, msgs, seed):
    import tfhe_rs_string_b200 as T
    prog = T.Program(engine, op, shape)
    cts = keys.encrypt_batch(np.asarray(msgs, dtype=U64).ravel(), seed=seed)
    out = keys.decrypt_batch(prog.run(cts))
    info = prog.info
    prog.close()
    return out, info


def test_config2_uint8_eq_and_add(engine, real_keys):
    # BASELINE configs[1] shape scaled to 256 pairs (parity case; the full 1024-pair shape is the bench)
    rng = np.random.default_rng(21)
    n = 256
    a = rng.integers(0, 256, n); b = rng.integers(0, 256, n); b[::2] = a[::2]
    msgs = np.concatenate([blocks_of(a).ravel(), blocks_of(b).ravel()])
    out, info = run(engine, real_keys, "radix_eq", [n, 4], msgs, 600)
    assert info["n_pbs"] == 3 * n and info["depth"] == 2   # two packed-pair equality flags + one all-true per integer
    assert np.array_equal(out, (a == b).astype(U64))
    out, info = run(engine, real_keys, "radix_add", [n, 4], msgs, 601)
    assert np.array_equal(from_blocks(out.reshape(n, 4)), ((a + b) % 256).astype(U64))
    out, _ = run(engine, real_keys, "radix_sub", [n, 4], msgs, 602)
    assert np.array_equal(from_blocks(out.reshape(n, 4)), ((a - b) % 256).astype(U64))


def test_scalar_comparisons_all_bytes(engine, real_keys):
    a = np.arange(256)
    for op, f in (("radix_scalar_gt", lambda x: x > 96), ("radix_scalar_lt", lambda x: x < 123), ("radix_scalar_eq", lambda x: x == 96)):
        scalar = 96 if op != "radix_scalar_lt" else 123
        out, _ = run(engine, real_keys, op, [256, 4, scalar], blocks_of(a).ravel(), 610)
        assert np.array_equal(out, f(a).astype(U64)), op


def test_encrypted_ordering_min_max(engine, real_keys):
    # unchecked_compare_parallelized / unchecked_min_or_max_parallelized (integer/server_key/comparator.rs:383-463,849-875)
    rng = np.random.default_rng(23)
    n = 192
    a = rng.integers(0, 256, n); b = rng.integers(0, 256, n)
    b[::4] = a[::4]; b[1::8] = a[1::8] ^ 1; a[2::16] = 255; b[3::16] = 0
    msgs = np.concatenate([blocks_of(a).ravel(), blocks_of(b).ravel()])
    for k, (op, f) in enumerate((("radix_gt", np.greater), ("radix_lt", np.less), ("radix_ge", np.greater_equal), ("radix_le", np.less_equal))):
        out, info = run(engine, real_keys, op, [n, 4], msgs, 630 + k)
        assert info["depth"] == 3 and info["n_pbs"] == 4 * n    # 2 packed signs + 1 reduction + 1 map per integer
        assert np.array_equal(out, f(a, b).astype(U64)), op
    out, info = run(engine, real_keys, "radix_max", [n, 4], msgs, 636)
    assert info["depth"] == 4
    assert out.max() < 4 and np.array_equal(from_blocks(out.reshape(n, 4)), np.maximum(a, b).astype(U64))
    out, _ = run(engine, real_keys, "radix_min", [n, 4], msgs, 637)
    assert out.max() < 4 and np.array_equal(from_blocks(out.reshape(n, 4)), np.minimum(a, b).astype(U64))


def test_config3_string_eq_and_uppercase(engine, real_keys):
    rng = np.random.default_rng(22)
    n, L = 16, 64
    a = ["".join(chr(rng.integers(32, 127)) for _ in range(L)) for _ in range(n)]
    b = list(a)
    for i in range(0, n, 2):
        p = int(rng.integers(0, L)); ch = chr(32 + (ord(b[i][p]) - 32 + 1) % 95)
        b[i] = b[i][:p] + ch + b[i][p + 1:]
    out, info = run(engine, real_keys, "string_eq", [n, L, L, 4], np.concatenate([chars(a).ravel(), chars(b).ravel()]), 620)
    assert list(out) == [int(x == y) for x, y in zip(a, b)]
    out, _ = run(engine, real_keys, "string_to_uppercase", [n, L, 4], chars(a).ravel(), 621)
    got = ["".join(chr(int(v)) for v in row) for row in from_blocks(out.reshape(n, L, 4))]
    assert got == [s.upper() for s in a]


def test_config4_contains_and_find(engine, real_keys):
    rng = np.random.default_rng(23)
    hay = "".join(chr(rng.integers(97, 123)) for _ in range(256))
    for k, pat in enumerate((hay[171:179], "qqqqqqqq")):
        msgs = np.concatenate([chars([hay]).ravel(), chars([pat]).ravel()])
        out, info = run(engine, real_keys, "string_contains", [1, 256, 8, 4], msgs, 630 + k)
        assert int(out[0]) == int(pat in hay)
        out, info = run(engine, real_keys, "string_find", [1, 256, 8, 4], msgs, 640 + k)
        assert int(out[0]) == int(pat in hay)
        assert int(from_blocks(out[1:].reshape(1, 4))[0]) == (hay.find(pat) if pat in hay else 256) % 256


def test_config5_trivium_known_answer(engine, real_keys):
    # ECRYPT vector of apps/trivium/benches/trivium_bool.rs:12,24 / test.rs:152-193, 64 keystream bytes
    k = json.load(open(os.path.join(HERE, "golden", "trivium_kat.json")))["kats"][3]
    iv = k["iv_bits"]
    iv_lo = sum(b << i for i, b in enumerate(iv[:64])); iv_hi = sum(b << i for i, b in enumerate(iv[64:]))
    out, info = run(engine, real_keys, "trivium", [8, iv_lo, iv_hi], np.array(k["key_bits"]), 650)
    by = bytes(sum(int(out[8 * i + j]) << j for j in range(8)) for i in range(64))
    assert by.hex().upper() == k["keystream_bytes_0_63_hex"]


def test_program_matches_cleartext_executor_and_errors(engine, real_keys):
    import tfhe_rs_string_b200 as T
    from oracle import oracle as O
    rng = np.random.default_rng(24)
    msgs = rng.integers(0, 4, 2 * 8 * 4)
    out, info = run(engine, real_keys, "radix_add", [8, 4], msgs, 660)
    assert np.array_equal(out, O.circuit_run_cleartext("radix_add", [8, 4], msgs))
    assert info == O.circuit_info("radix_add", [8, 4])
    with pytest.raises(T.B200TfheError):
        T.Program(engine, "no_such_op", [1])
    with pytest.raises(T.B200TfheError):
        T.Program(engine, "radix_eq", [1])


def test_extra_ops_on_gpu(engine, real_keys):
    """A sample of the remaining call sites through the GPU executor: ne, bitxor, shl, scalar ge,
    to_lowercase, starts_with, ends_with."""
    rng = np.random.default_rng(25)
    n = 32
    a = rng.integers(0, 256, n); b = rng.integers(0, 256, n); b[::2] = a[::2]
    msgs = np.concatenate([blocks_of(a).ravel(), blocks_of(b).ravel()])
    out, _ = run(engine, real_keys, "radix_ne", [n, 4], msgs, 670)
    assert np.array_equal(out, (a != b).astype(U64))
    out, _ = run(engine, real_keys, "radix_bitxor", [n, 4], msgs, 671)
    assert np.array_equal(from_blocks(out.reshape(n, 4)), (a ^ b).astype(U64))
    out, _ = run(engine, real_keys, "radix_shl", [n, 4, 3], blocks_of(a).ravel(), 672)
    assert np.array_equal(from_blocks(out.reshape(n, 4)), ((a << 3) % 256).astype(U64))
    out, _ = run(engine, real_keys, "radix_scalar_ge", [n, 4, 100], blocks_of(a).ravel(), 673)
    assert np.array_equal(out, (a >= 100).astype(U64))
    s = ["Hello, World Q", "tfhe-RS string", "ABCxyz 019 ~!@", "mixedCASEinput"]
    out, _ = run(engine, real_keys, "string_to_lowercase", [4, 14, 4], chars(s).ravel(), 674)
    assert ["".join(chr(int(v)) for v in row) for row in from_blocks(out.reshape(4, 14, 4))] == [x.lower() for x in s]
    pats = ["Hell", "ring", "ABCx", "nput"]
    msgs = np.concatenate([chars(s).ravel(), chars(pats).ravel()])
    out, _ = run(engine, real_keys, "string_starts_with", [4, 14, 4, 4], msgs, 675)
    assert list(out) == [int(x.startswith(q)) for x, q in zip(s, pats)]
    out, _ = run(engine, real_keys, "string_ends_with", [4, 14, 4, 4], msgs, 676)
    assert list(out) == [int(x.endswith(q)) for x, q in zip(s, pats)]


# ---- BASELINE.json shapes at FULL size (the tests above use reduced batches) -----------------------------
def test_full_size_config2_1024_uint8_pairs(engine, real_keys):
    rng = np.random.default_rng(31)
    n = 1024
    a = rng.integers(0, 256, n); b = rng.integers(0, 256, n); b[::2] = a[::2]
    msgs = np.concatenate([blocks_of(a).ravel(), blocks_of(b).ravel()])
    out, info = run(engine, real_keys, "radix_eq", [n, 4], msgs, 700)
    assert info["n_pbs"] == 3072 and np.array_equal(out, (a == b).astype(U64))
    out, info = run(engine, real_keys, "radix_add", [n, 4], msgs, 701)
    assert np.array_equal(from_blocks(out.reshape(n, 4)), ((a + b) % 256).astype(U64))


def test_full_size_config3_256_strings_of_64_chars(engine, real_keys):
    rng = np.random.default_rng(32)
    n, L = 256, 64
    a = ["".join(chr(rng.integers(32, 127)) for _ in range(L)) for _ in range(n)]
    b = list(a)
    for i in range(0, n, 2):   # half equal, half differing in one random position
        q = int(rng.integers(0, L)); b[i] = b[i][:q] + chr(32 + (ord(b[i][q]) - 31) % 95) + b[i][q + 1:]
    out, info = run(engine, real_keys, "string_eq", [n, L, L, 4], np.concatenate([chars(a).ravel(), chars(b).ravel()]), 710)
    assert info["n_pbs"] == 35072 and info["depth"] == 3   # 256 x (128 packed-pair comparisons -> 8 -> 1)
    assert list(out) == [int(x == y) for x, y in zip(a, b)]


def test_full_size_config3_to_uppercase_256_strings(engine, real_keys):
    """BASELINE configs[2], second half: to_uppercase on 256 strings of 64 chars: 49,152 bootstraps at depth 2 with the
    batch schedule, and the reference's own operator decomposition (294,912 bootstraps, depth 7) on the same ciphertexts."""
    rng = np.random.default_rng(33)
    n, L = 256, 64
    a = ["".join(chr(rng.integers(32, 127)) for _ in range(L)) for _ in range(n)]
    out, info = run(engine, real_keys, "string_to_uppercase", [n, L, 4], chars(a).ravel(), 711)
    assert info["n_pbs"] == 49152 and info["depth"] == 2
    got = from_blocks(out.reshape(n, L, 4))
    assert ["".join(chr(int(c)) for c in row) for row in got] == [s.upper() for s in a]
    out, info = run(engine, real_keys, "string_to_uppercase_reference", [n, L, 4], chars(a).ravel(), 711)
    assert info["n_pbs"] == 294912 and info["depth"] == 7
    assert np.array_equal(from_blocks(out.reshape(n, L, 4)), got)


def test_full_size_config5_trivium_1024_bits(engine, real_keys):
    from oracle import oracle as O
    k = json.load(open(os.path.join(HERE, "golden", "trivium_kat.json")))["kats"][3]
    iv = k["iv_bits"]
    iv_lo = sum(b << i for i, b in enumerate(iv[:64])); iv_hi = sum(b << i for i, b in enumerate(iv[64:]))
    out, info = run(engine, real_keys, "trivium", [16, iv_lo, iv_hi], np.array(k["key_bits"]), 720)
    assert len(out) == 1024
    by = bytes(sum(int(out[8 * i + j]) << j for j in range(8)) for i in range(128))
    assert by[:64].hex().upper() == k["keystream_bytes_0_63_hex"]          # the reference's known answer
    clear = O.circuit_run_cleartext("trivium", [16, iv_lo, iv_hi], k["key_bits"])
    assert np.array_equal(out, clear)                                        # and the cleartext executor beyond it
