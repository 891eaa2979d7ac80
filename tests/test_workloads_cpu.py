"""CPU tests of the host-side batched call sites (tfhe_rs_string_b200/csrc/{circuit,workloads,programs}.hpp)
run through the test-only cleartext / CPU-oracle executors: results must equal the clear operation
(`==`, wrapping add/sub, str.upper(), `in`, str.find) exactly, as the reference's own tests require
(integer/server_key/radix_parallel/tests_cases_unsigned.rs, examples/fhe_strings/test_generating_macros.rs)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def blocks_of(values, nb=4):
    v = np.asarray(values, dtype=np.uint64)
    return np.stack([(v >> np.uint64(2 * k)) & np.uint64(3) for k in range(nb)], axis=-1)


def from_blocks(b):
    b = np.asarray(b, dtype=np.uint64)
    return sum(b[..., k] << np.uint64(2 * k) for k in range(b.shape[-1]))


def chars(strings):
    return blocks_of(np.array([[ord(ch) for ch in s] for s in strings], dtype=np.uint64))


def test_trivium_known_answers_cleartext():
    # ECRYPT vectors pinned by the reference: apps/trivium/src/trivium/test.rs:79-193
    kats = json.load(open(os.path.join(HERE, "golden", "trivium_kat.json")))["kats"]
    assert len(kats) == 4
    for k in kats:
        iv = k["iv_bits"]
        iv_lo = sum(b << i for i, b in enumerate(iv[:64])); iv_hi = sum(b << i for i, b in enumerate(iv[64:]))
        out = O.circuit_run_cleartext("trivium", [8, iv_lo, iv_hi], k["key_bits"])
        by = bytes(sum(int(out[8 * i + j]) << j for j in range(8)) for i in range(64))
        assert by.hex().upper() == k["keystream_bytes_0_63_hex"], k["name"]


def test_radix_eq_add_sub_cleartext():
    rng = np.random.default_rng(1)
    n = 500
    a = rng.integers(0, 256, n); b = rng.integers(0, 256, n)
    b[::3] = a[::3]
    inp = np.concatenate([blocks_of(a).ravel(), blocks_of(b).ravel()])
    assert np.array_equal(O.circuit_run_cleartext("radix_eq", [n, 4], inp), (a == b).astype(np.uint64))
    s = O.circuit_run_cleartext("radix_add", [n, 4], inp).reshape(n, 4)
    assert s.max() < 4 and np.array_equal(from_blocks(s), ((a + b) % 256).astype(np.uint64))
    d = O.circuit_run_cleartext("radix_sub", [n, 4], inp).reshape(n, 4)
    assert d.max() < 4 and np.array_equal(from_blocks(d), ((a - b) % 256).astype(np.uint64))


def test_encrypted_ordering_min_max_cleartext():
    """unchecked_compare_parallelized / unchecked_min_or_max_parallelized (integer/server_key/comparator.rs:383-463,849-875):
    every pair of a small exhaustive grid plus random bytes, against Python's operators."""
    rng = np.random.default_rng(5)
    g = np.array([0, 1, 3, 4, 15, 16, 17, 63, 64, 127, 128, 200, 254, 255])
    a = np.concatenate([np.repeat(g, len(g)), rng.integers(0, 256, 400)])
    b = np.concatenate([np.tile(g, len(g)), rng.integers(0, 256, 400)])
    n = len(a)
    inp = np.concatenate([blocks_of(a).ravel(), blocks_of(b).ravel()])
    for op, f in (("radix_gt", np.greater), ("radix_lt", np.less), ("radix_ge", np.greater_equal), ("radix_le", np.less_equal)):
        assert np.array_equal(O.circuit_run_cleartext(op, [n, 4], inp), f(a, b).astype(np.uint64)), op
    mx = O.circuit_run_cleartext("radix_max", [n, 4], inp).reshape(n, 4)
    mn = O.circuit_run_cleartext("radix_min", [n, 4], inp).reshape(n, 4)
    assert mx.max() < 4 and np.array_equal(from_blocks(mx), np.maximum(a, b).astype(np.uint64))
    assert mn.max() < 4 and np.array_equal(from_blocks(mn), np.minimum(a, b).astype(np.uint64))
    # odd block counts leave the top block unpacked (comparator.rs:430-443 chunks of 2)
    a3, b3 = a % 64, b % 64
    inp3 = np.concatenate([blocks_of(a3)[:, :3].ravel(), blocks_of(b3)[:, :3].ravel()])
    assert np.array_equal(O.circuit_run_cleartext("radix_gt", [n, 3], inp3), (a3 > b3).astype(np.uint64))


@pytest.mark.parametrize("scalar", [0, 1, 96, 123, 200, 255, 256, 1000])
def test_scalar_comparisons_cleartext(scalar):
    a = np.arange(256)
    inp = blocks_of(a).ravel()
    assert np.array_equal(O.circuit_run_cleartext("radix_scalar_gt", [256, 4, scalar], inp), (a > scalar).astype(np.uint64))
    assert np.array_equal(O.circuit_run_cleartext("radix_scalar_lt", [256, 4, scalar], inp), (a < scalar).astype(np.uint64))
    assert np.array_equal(O.circuit_run_cleartext("radix_scalar_eq", [256, 4, scalar], inp), (a == scalar).astype(np.uint64))


def test_to_uppercase_every_byte_cleartext():
    s = ["".join(chr(c) for c in range(128))]
    out = O.circuit_run_cleartext("string_to_uppercase", [1, 128, 4], chars(s).ravel()).reshape(128, 4)
    assert "".join(chr(int(v)) for v in from_blocks(out)) == s[0].upper()
    # bytes >= 128 must be left alone like u8 arithmetic does
    hi = blocks_of(np.arange(128, 256))
    out = O.circuit_run_cleartext("string_to_uppercase", [1, 128, 4], hi.ravel()).reshape(128, 4)
    assert np.array_equal(from_blocks(out), np.arange(128, 256).astype(np.uint64))


def test_case_change_fast_schedule_equals_reference_decomposition():
    """The 3-bootstrap schedule (case_change_char, workloads.hpp) and the reference's operator decomposition
    (change_case.rs:53-82: two scalar comparisons, and, shift, add / sub with propagation) on every byte value."""
    allb = blocks_of(np.arange(256)).ravel()
    want_u = np.array([ord(chr(v).upper()) if 97 <= v <= 122 else v for v in range(256)], dtype=np.uint64)
    want_l = np.array([v + 32 if 65 <= v <= 90 else v for v in range(256)], dtype=np.uint64)
    for op, want in (("string_to_uppercase", want_u), ("string_to_lowercase", want_l)):
        fast = O.circuit_run_cleartext(op, [1, 256, 4], allb).reshape(256, 4)
        ref = O.circuit_run_cleartext(op + "_reference", [1, 256, 4], allb).reshape(256, 4)
        assert fast.max() < 4 and np.array_equal(fast, ref) and np.array_equal(from_blocks(fast), want), op
    assert O.circuit_info("string_to_uppercase", [1, 64, 4])["n_pbs"] == 3 * 64
    assert O.circuit_info("string_to_uppercase", [1, 64, 4])["depth"] == 2
    assert O.circuit_info("string_to_uppercase_reference", [1, 64, 4])["n_pbs"] == 18 * 64


def test_string_eq_cleartext_config3_shape():
    # BASELINE configs[2]: 256 pairs of 64-char strings, half equal / half differing in one position
    rng = np.random.default_rng(3)
    n, L = 256, 64
    a = ["".join(chr(rng.integers(32, 127)) for _ in range(L)) for _ in range(n)]
    b = list(a)
    for i in range(0, n, 2):
        p = int(rng.integers(0, L)); ch = chr(32 + (ord(b[i][p]) - 32 + 1) % 95)
        b[i] = b[i][:p] + ch + b[i][p + 1:]
    inp = np.concatenate([chars(a).ravel(), chars(b).ravel()])
    out = O.circuit_run_cleartext("string_eq", [n, L, L, 4], inp)
    assert list(out) == [int(x == y) for x, y in zip(a, b)]
    # different clear lengths: never equal for unpadded strings (comparisons.rs:195-212)
    out = O.circuit_run_cleartext("string_eq", [1, 3, 2, 4], np.concatenate([chars(["abc"]).ravel(), chars(["ab"]).ravel()]))
    assert list(out) == [0]
    # Padding::Final (zero characters after the content, eq_encrypted -> eq_no_init_padding, comparisons.rs:101-124,184-215):
    # the same content behind different amounts of padding is equal, an extra character is not
    for sa, sb, want in (("abc\0\0", "abc\0\0", 1), ("abc\0\0", "abc", 1), ("abc", "abc\0\0\0", 1), ("abcd\0", "abc", 0),
                         ("abc\0\0", "abd\0\0", 0), ("\0\0", "\0\0\0", 1), ("ab\0\0", "abc\0", 0)):
        out = O.circuit_run_cleartext("string_eq", [1, len(sa), len(sb), 4], np.concatenate([chars([sa]).ravel(), chars([sb]).ravel()]))
        assert list(out) == [want], (sa, sb)


@pytest.mark.parametrize("hay,pat", [("the quick brown fox jumps over the lazy dog", "brown"), ("aaaaaaaab", "aab"), ("abc", "abcd"),
                                     ("hello world", "xyz"), ("zzzz", "z"), ("abcabcabd", "abd"), ("x" * 40 + "needle", "needle")])
def test_contains_find_cleartext(hay, pat):
    inp = np.concatenate([chars([hay]).ravel(), chars([pat]).ravel()])
    shape = [1, len(hay), len(pat), 4]
    assert int(O.circuit_run_cleartext("string_contains", shape, inp)[0]) == int(pat in hay)
    out = O.circuit_run_cleartext("string_find", shape, inp)
    found, index = int(out[0]), int(from_blocks(out[1:].reshape(1, 4))[0])
    assert found == int(pat in hay)
    assert index == (hay.find(pat) if pat in hay else len(hay)) % 256   # find.rs:139-160: index counts the positions before the first match


def test_contains_find_config4_shape_cleartext():
    rng = np.random.default_rng(4)
    hay = "".join(chr(rng.integers(97, 123)) for _ in range(256))
    for pat in (hay[171:179], "qqqqqqqq", hay[248:256], hay[0:8]):
        inp = np.concatenate([chars([hay]).ravel(), chars([pat]).ravel()])
        out = O.circuit_run_cleartext("string_find", [1, 256, 8, 4], inp)
        found, index = int(out[0]), int(from_blocks(out[1:].reshape(1, 4))[0])
        assert found == int(pat in hay) and index == (hay.find(pat) if pat in hay else 256) % 256
        assert int(O.circuit_run_cleartext("string_contains", [1, 256, 8, 4], inp)[0]) == int(pat in hay)


def test_workloads_encrypted_on_toy_parameters(toy_keys):
    # same programs on real ciphertexts with the CPU oracle's KS+PBS (insecure fast parameters,
    # examples/fhe_strings/ciphertext.rs:76-90) -- checks that degrees / noise stay decryptable
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, 6); b = rng.integers(0, 256, 6); b[:2] = a[:2]
    cts = toy_keys.encrypt_batch(np.concatenate([blocks_of(a).ravel(), blocks_of(b).ravel()]), seed=50)
    assert list(toy_keys.decrypt_batch(O.circuit_run_encrypted(toy_keys, "radix_eq", [6, 4], cts))) == list((a == b).astype(int))
    assert list(toy_keys.decrypt_batch(O.circuit_run_encrypted(toy_keys, "radix_gt", [6, 4], cts))) == list((a > b).astype(int))
    mx = toy_keys.decrypt_batch(O.circuit_run_encrypted(toy_keys, "radix_max", [6, 4], cts)).reshape(6, 4)
    assert np.array_equal(from_blocks(mx), np.maximum(a, b).astype(np.uint64))
    s = toy_keys.decrypt_batch(O.circuit_run_encrypted(toy_keys, "radix_add", [6, 4], cts)).reshape(6, 4)
    assert np.array_equal(from_blocks(s), ((a + b) % 256).astype(np.uint64))
    text = ["Hello, FHE world {z}~`a"]
    out = toy_keys.decrypt_batch(O.circuit_run_encrypted(toy_keys, "string_to_uppercase", [1, len(text[0]), 4],
                                                         toy_keys.encrypt_batch(chars(text).ravel(), seed=51)))
    assert "".join(chr(int(v)) for v in from_blocks(out.reshape(-1, 4))) == text[0].upper()
    hay, pat = "find the needle here", "needle"
    cts = toy_keys.encrypt_batch(np.concatenate([chars([hay]).ravel(), chars([pat]).ravel()]), seed=52)
    out = toy_keys.decrypt_batch(O.circuit_run_encrypted(toy_keys, "string_find", [1, len(hay), len(pat), 4], cts))
    assert int(out[0]) == 1 and int(from_blocks(out[1:].reshape(1, 4))[0]) == hay.find(pat)


def test_trivium_encrypted_on_toy_parameters(toy_keys):
    k = json.load(open(os.path.join(HERE, "golden", "trivium_kat.json")))["kats"][3]
    iv = k["iv_bits"]
    iv_lo = sum(b << i for i, b in enumerate(iv[:64])); iv_hi = sum(b << i for i, b in enumerate(iv[64:]))
    cts = toy_keys.encrypt_batch(np.array(k["key_bits"]), seed=53)
    out = toy_keys.decrypt_batch(O.circuit_run_encrypted(toy_keys, "trivium", [1, iv_lo, iv_hi], cts))
    by = bytes(sum(int(out[8 * i + j]) << j for j in range(8)) for i in range(8))
    assert by.hex().upper() == k["keystream_bytes_0_63_hex"][:16]


def test_extra_radix_ops_cleartext():
    """ne / bitand / bitor / bitxor / shl / scalar le, ge (radix_parallel/comparison.rs:39-62,
    bitwise_op.rs, scalar_shift.rs, comparator.rs) against the clear operations."""
    rng = np.random.default_rng(7)
    n = 300
    a = rng.integers(0, 256, n); b = rng.integers(0, 256, n)
    b[::4] = a[::4]
    inp = np.concatenate([blocks_of(a).ravel(), blocks_of(b).ravel()])
    assert np.array_equal(O.circuit_run_cleartext("radix_ne", [n, 4], inp), (a != b).astype(np.uint64))
    for op, f in (("radix_bitand", np.bitwise_and), ("radix_bitor", np.bitwise_or), ("radix_bitxor", np.bitwise_xor)):
        r = O.circuit_run_cleartext(op, [n, 4], inp).reshape(n, 4)
        assert r.max() < 4 and np.array_equal(from_blocks(r), f(a, b).astype(np.uint64)), op
    for bits in (0, 1, 2, 3, 5, 7):
        r = O.circuit_run_cleartext("radix_shl", [n, 4, bits], blocks_of(a).ravel()).reshape(n, 4)
        assert np.array_equal(from_blocks(r), ((a << bits) % 256).astype(np.uint64)), bits
    v = np.arange(256)
    for scalar in (0, 64, 91, 255):
        assert np.array_equal(O.circuit_run_cleartext("radix_scalar_le", [256, 4, scalar], blocks_of(v).ravel()), (v <= scalar).astype(np.uint64))
        assert np.array_equal(O.circuit_run_cleartext("radix_scalar_ge", [256, 4, scalar], blocks_of(v).ravel()), (v >= scalar).astype(np.uint64))


def test_extra_string_ops_cleartext():
    """ne, to_lowercase, starts_with, ends_with (examples/fhe_strings/server_key/comparisons.rs:37-40,
    change_case.rs:40-82, contains.rs:96-134, ends_with.rs) on unpadded strings."""
    rng = np.random.default_rng(8)
    n, L, P = 24, 12, 4
    a = ["".join(chr(rng.integers(32, 127)) for _ in range(L)) for _ in range(n)]
    b = list(a)
    for i in range(0, n, 2):
        q = int(rng.integers(0, L)); b[i] = b[i][:q] + chr(32 + (ord(b[i][q]) - 31) % 95) + b[i][q + 1:]
    inp = np.concatenate([chars(a).ravel(), chars(b).ravel()])
    assert list(O.circuit_run_cleartext("string_ne", [n, L, L, 4], inp)) == [int(x != y) for x, y in zip(a, b)]
    low = O.circuit_run_cleartext("string_to_lowercase", [n, L, 4], chars(a).ravel()).reshape(n, L, 4)
    assert ["".join(chr(int(v)) for v in row) for row in from_blocks(low)] == [s.lower() for s in a]
    pats = [s[:P] if i % 3 == 0 else (s[-P:] if i % 3 == 1 else "zz" + s[:P - 2]) for i, s in enumerate(a)]
    inp = np.concatenate([chars(a).ravel(), chars(pats).ravel()])
    assert list(O.circuit_run_cleartext("string_starts_with", [n, L, P, 4], inp)) == [int(s.startswith(q)) for s, q in zip(a, pats)]
    assert list(O.circuit_run_cleartext("string_ends_with", [n, L, P, 4], inp)) == [int(s.endswith(q)) for s, q in zip(a, pats)]
    # pattern longer than the string: never a prefix / suffix of an unpadded string without zero characters
    inp = np.concatenate([chars(pats).ravel(), chars(a).ravel()])
    assert not O.circuit_run_cleartext("string_starts_with", [n, P, L, 4], inp).any()
    assert not O.circuit_run_cleartext("string_ends_with", [n, P, L, 4], inp).any()
