"""numpy restatement of the reference's u32 boolean gate path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (only tests/
may import it).  Paths relative to /root/reference/tfhe/src/.

  parameters            boolean/parameters/mod.rs:123-192
  encrypt / decrypt     boolean/engine/mod.rs:236-262 (PLAINTEXT_TRUE = 1/8, PLAINTEXT_FALSE = 7/8, boolean/mod.rs:77-80),
                        client decryption boolean/engine/mod.rs:330-372 (phase rounded to the nearest multiple of 1/4... sign test)
  gates                 boolean/engine/mod.rs:606-850 (linear pre-combination), bootstrapping.rs:257-391 (pattern)
  key switch            core_crypto/algorithms/lwe_keyswitch.rs:96-170
  bootstrap             fft_impl/fft64/crypto/bootstrap.rs:242-364, ggsw.rs:477-598 (multi-level external product),
                        decomposition commons/math/decomposition/{decomposer.rs:98-116, iter.rs:120-127}
  key generation        lwe_keyswitch_key_generation.rs:65-130, lwe_bootstrap_key_generation.rs:76-141, ggsw_encryption.rs:72-151

Parity status: integer stages are exact restatements; the Fourier stage uses numpy's FFT (the reference uses
concrete-fft, absent from the tree), so post-bootstrap ciphertext bits are unpinned, decrypted bits are pinned by the
truth tables (the reference's own boolean tests check exactly that, boolean/server_key/tests.rs)."""
import numpy as np

U32 = np.uint32
U64 = np.uint64
TRUE = U32(1 << 29)
FALSE = U32(7 << 29)
GATES = {"and": 0, "nand": 1, "or": 2, "nor": 3, "xor": 4, "xnor": 5}
TRUTH = {0: lambda a, b: a & b, 1: lambda a, b: 1 - (a & b), 2: lambda a, b: a | b, 3: lambda a, b: 1 - (a | b),
         4: lambda a, b: a ^ b, 5: lambda a, b: 1 - (a ^ b)}


class BooleanParams:
    def __init__(self, n, k, N, lwe_std, glwe_std, pbs_base_log, pbs_level, ks_base_log, ks_level, ks_first):
        self.lwe_dimension, self.glwe_dimension, self.polynomial_size = n, k, N
        self.lwe_std, self.glwe_std = lwe_std, glwe_std
        self.pbs_base_log, self.pbs_level, self.ks_base_log, self.ks_level = pbs_base_log, pbs_level, ks_base_log, ks_level
        self.ks_first = ks_first          # EncryptionKeyChoice::Big

    @property
    def big(self):
        return self.glwe_dimension * self.polynomial_size + 1

    @property
    def small(self):
        return self.lwe_dimension + 1

    @property
    def ct_size(self):
        return self.big if self.ks_first else self.small


def default_parameters():            # DEFAULT_PARAMETERS (EncryptionKeyChoice::Small)
    return BooleanParams(722, 2, 512, 0.000013071021089943935, 0.00000004990272175010415, 6, 3, 3, 4, False)


def default_parameters_ks_pbs():     # DEFAULT_PARAMETERS_KS_PBS (EncryptionKeyChoice::Big)
    return BooleanParams(664, 2, 512, 0.00003808282923459771, 0.00000004990272175010415, 6, 3, 3, 4, True)


def _negacyclic_matrix(s):
    """M with (a @ M)[j] = coefficient j of a(X) * s(X) mod X^N + 1, for a binary polynomial s (entries 0, +-1)."""
    N = len(s)
    idx = (np.arange(N)[None, :] - np.arange(N)[:, None]) % (2 * N)      # idx[i][j] = (j - i) mod 2N: a_i * s_{j-i}
    m = s[idx % N].astype(np.float64)
    m[idx >= N] *= -1.0
    return m


def _gauss(rng, std, shape):
    return np.rint(rng.normal(0.0, std, shape) * 2.0**32).astype(np.int64).astype(U32)


class BooleanKeyset:
    def __init__(self, params, seed=1):
        p = self.params = params
        rng = self.rng = np.random.default_rng(seed)
        n, k, N = p.lwe_dimension, p.glwe_dimension, p.polynomial_size
        self.small_sk = rng.integers(0, 2, n).astype(U32)
        self.glwe_sk = rng.integers(0, 2, (k, N)).astype(U32)
        self.big_sk = self.glwe_sk.reshape(-1)
        mats = [_negacyclic_matrix(self.glwe_sk[r].astype(np.int64)) for r in range(k)]
        # ---- bootstrap key: GGSW(small_sk[i]), layout [n][level 1..l][k+1 rows][k+1 polys][N]  (ggsw_encryption.rs:116-151)
        L = p.pbs_level
        bsk = np.zeros((n, L, k + 1, k + 1, N), dtype=U32)
        mask = rng.integers(0, 2**32, (n, L, k + 1, k, N), dtype=np.int64)
        bsk[:, :, :, :k, :] = mask.astype(U32)
        body = np.zeros((n, L, k + 1, N), dtype=np.float64)
        for r in range(k):
            body += (mask[:, :, :, r, :].astype(np.float64).reshape(-1, N) @ mats[r]).reshape(n, L, k + 1, N)
        body = np.mod(body, 2.0**32).astype(np.int64).astype(U32)
        body += _gauss(rng, p.glwe_std, body.shape)
        for lvl in range(1, L + 1):
            factor = (U32(0) - self.small_sk) << U32(32 - p.pbs_base_log * lvl)                   # -s_i * 2^(32 - b l)
            for r in range(k):                                                                  # row r: factor * S_r(X)
                body[:, lvl - 1, r, :] += factor[:, None] * self.glwe_sk[r][None, :]
            body[:, lvl - 1, k, 0] += U32(0) - factor                                            # last row: -factor at X^0
        bsk[:, :, :, k, :] = body
        self.bsk_standard = np.ascontiguousarray(bsk).reshape(-1)
        # ---- keyswitch key: [k N][level l..1][n + 1]  (lwe_keyswitch_key_generation.rs:109-128)
        KL = p.ks_level
        a = rng.integers(0, 2**32, (k * N, KL, n), dtype=np.int64)
        b = (a.reshape(-1, n).astype(np.float64) @ self.small_sk.astype(np.float64)).reshape(k * N, KL)
        b = np.mod(b, 2.0**32).astype(np.int64).astype(U32) + _gauss(rng, p.lwe_std, (k * N, KL))
        for li in range(KL):
            lvl = KL - li
            b[:, li] += self.big_sk << U32(32 - p.ks_base_log * lvl)
        ksk = np.zeros((k * N, KL, n + 1), dtype=U32)
        ksk[:, :, :n] = a.astype(U32)
        ksk[:, :, n] = b
        self.ksk = np.ascontiguousarray(ksk).reshape(-1)
        # ---- Fourier bootstrap key for the oracle's own bootstrap
        self._tw = np.exp(1j * np.pi * np.arange(N // 2) / N)
        std = self.bsk_standard.reshape(-1, N).astype(np.int32).astype(np.float64) * 2.0**-32
        self._bsk_f = np.fft.fft((std[:, :N // 2] + 1j * std[:, N // 2:]) * self._tw[None, :], axis=1).reshape(n, L, k + 1, k + 1, N // 2)

    # ---- client side
    def _key(self):
        return self.big_sk if self.params.ks_first else self.small_sk

    def encrypt(self, bits, seed=2):
        rng = np.random.default_rng(seed)
        p, sk = self.params, self._key()
        bits = np.asarray(bits)
        a = rng.integers(0, 2**32, (len(bits), len(sk)), dtype=np.int64)
        std = p.glwe_std if p.ks_first else p.lwe_std
        body = np.mod(a.astype(np.float64) @ sk.astype(np.float64), 2.0**32).astype(np.int64).astype(U32)
        body += np.where(bits != 0, TRUE, FALSE).astype(U32) + _gauss(rng, std, len(bits))
        ct = np.zeros((len(bits), len(sk) + 1), dtype=U32)
        ct[:, :-1], ct[:, -1] = a.astype(U32), body
        return ct

    def phase(self, cts):
        cts = np.asarray(cts, dtype=U32)
        sk = self._key()
        dot = np.mod(cts[:, :-1].astype(np.float64) @ sk.astype(np.float64), 2.0**32).astype(np.int64).astype(U32)
        return cts[:, -1] - dot

    def decrypt(self, cts):
        """true iff the phase is in the upper half-plane of the encoding: closest of {1/8, 7/8} (engine/mod.rs:330-372)."""
        return (self.phase(cts).astype(np.int32) > 0).astype(np.int64)

    # ---- server side (the reference algorithm)
    def _keyswitch(self, cts):
        p = self.params
        n, KL, bl = p.lwe_dimension, p.ks_level, p.ks_base_log
        ksk = self.ksk.reshape(-1, KL, n + 1)
        out = np.zeros((cts.shape[0], n + 1), dtype=U32)
        out[:, n] = cts[:, -1]
        non_rep = 32 - bl * KL
        state = ((cts[:, :-1] >> U32(non_rep - 1)) + U32(1)) >> U32(1)
        mask = U32((1 << bl) - 1)
        for li in range(KL):
            res = state & mask
            state = state >> U32(bl)
            carry = (((res - U32(1)) | state) & res) >> U32(bl - 1)
            state = state + carry
            digit = (res - (carry << U32(bl))).astype(np.int32).astype(np.float64)           # [B, kN], |d| <= 2^(bl-1)
            # out -= sum_i digit_i * ksk[i][li]: exact in float64 (|d| <= 4, 2^32 words, 1024 terms < 2^45)
            out -= np.mod(digit @ ksk[:, li, :].astype(np.float64), 2.0**32).astype(np.int64).astype(U32)
        return out

    def _bootstrap(self, small):
        p = self.params
        n, k, N, L, bl = p.lwe_dimension, p.glwe_dimension, p.polynomial_size, p.pbs_level, p.pbs_base_log
        B, half = small.shape[0], N // 2
        ms = lambda x: (((x >> U32(32 - (N.bit_length() - 1) - 2)) + U32(1)) >> U32(1)).astype(np.int64)   # fast_pbs_modulus_switch, common.rs:26-43
        j = np.arange(N)[None, :]

        def rot(poly, deg):           # poly * X^deg, deg in [0, 2N], batched: poly [B, N], deg [B]
            idx = (j - deg[:, None]) % (2 * N)
            v = np.take_along_axis(poly, idx % N, axis=1)
            return np.where(idx >= N, U32(0) - v, v)

        acc = np.zeros((B, k + 1, N), dtype=U32)
        acc[:, k, :] = rot(np.full((B, N), TRUE, dtype=U32), 2 * N - ms(small[:, n]))
        non_rep = 32 - bl * L
        mask = U32((1 << bl) - 1)
        for i in range(n):
            ah = ms(small[:, i])
            ct1 = np.stack([rot(acc[:, r, :], ah) - acc[:, r, :] for r in range(k + 1)], axis=1)
            state = ((ct1 >> U32(non_rep - 1)) + U32(1)) >> U32(1)
            outf = np.zeros((B, k + 1, half), dtype=np.complex128)
            for lvl in range(L, 0, -1):
                res = state & mask
                state = state >> U32(bl)
                carry = (((res - U32(1)) | state) & res) >> U32(bl - 1)
                state = state + carry
                d = (res - (carry << U32(bl))).astype(np.int32).astype(np.float64)             # [B, k+1, N]
                f = np.fft.fft((d[:, :, :half] + 1j * d[:, :, half:]) * self._tw[None, None, :], axis=2)
                outf += np.einsum("brf,rcf->bcf", f, self._bsk_f[i, lvl - 1])
            y = np.fft.ifft(outf, axis=2) * np.conj(self._tw)[None, None, :]
            for part, sl in ((y.real, slice(0, half)), (y.imag, slice(half, N))):
                fr = part - np.rint(part)
                acc[:, :, sl] += np.rint(fr * 2.0**32).astype(np.int64).astype(U32)
        out = np.zeros((B, k * N + 1), dtype=U32)
        for r in range(k):
            out[:, r * N] = acc[:, r, 0]
            out[:, r * N + 1:(r + 1) * N] = U32(0) - acc[:, r, :0:-1]
        out[:, k * N] = acc[:, k, 0]
        return out

    def gate(self, gate, a, b):
        g = GATES[gate] if isinstance(gate, str) else gate
        coeff = [1, -1, 1, -1, 2, -2][g]
        cst = [FALSE, TRUE, TRUE, FALSE, U32(2 << 29), U32(6 << 29)][g]
        pre = (np.asarray(a, dtype=U32) + np.asarray(b, dtype=U32)) * U32(coeff % 2**32)
        pre[:, -1] += cst
        if self.params.ks_first:
            return self._bootstrap(self._keyswitch(pre))
        return self._keyswitch(self._bootstrap(pre))
