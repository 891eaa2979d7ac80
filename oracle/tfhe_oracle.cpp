/*
 * tfhe_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).  See tfhe_oracle.h.
 *
 * Restates, function by function, the reference's CPU algorithm for shortint KS+PBS.
 * Citations are relative to /root/reference/tfhe/src/.  Nothing here is copied: the reference is
 * Rust over generic containers, this is flat C++ over u64 arrays.
 *
 * Third-party arithmetic that is NOT in /root/reference: the complex FFT itself comes from the
 * crate concrete-fft 0.3.0 (tfhe/Cargo.toml:60; call sites core_crypto/fft_impl/fft64/math/fft/
 * mod.rs:161,513,533,553).  Its published algorithm is an ordinary power-of-two complex DFT
 * (unnormalised forward, unnormalised inverse, output order plan dependent).  We restate it as
 * a textbook iterative radix-2 transform in f64; frequency ordering is free because the Fourier
 * BSK and the operand go through the same transform.  Fourier-domain bits are therefore
 * "parity unpinned" (tfhe_oracle.h); the reference's own tests only pin tolerances there
 * (fft/tests.rs:9-80,82-222) and we re-run those tolerances in tests/test_oracle.py.
 */
#include "tfhe_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace {

/* ---------------------------------------------------------------- seeded PRNG (test fixture) */
struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) {
        for (auto &w : s) w = splitmix(seed);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() { /* xoshiro256** */
        uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    /* commons/math/random/gaussian.rs:15-52: polar Box-Muller on two signed 64-bit uniforms
     * scaled to [-1,1); returns the pair. */
    void gaussian_pair(double std, double &g0, double &g1) {
        for (;;) {
            double u = (double)(int64_t)next() * 0x1p-63;
            double v = (double)(int64_t)next() * 0x1p-63;
            double s2 = u * u + v * v;
            if (s2 > 0.0 && s2 < 1.0) {
                double cst = std * std::sqrt(-2.0 * std::log(s2) / s2);
                g0 = u * cst; g1 = v * cst;
                return;
            }
        }
    }
};

uint64_t derive_seed(uint64_t seed, uint64_t tag, uint64_t idx) {
    uint64_t x = seed ^ (tag * 0xD6E8FEB86659FD93ull) ^ (idx * 0xA0761D6478BD642Full);
    Rng::splitmix(x);
    return Rng::splitmix(x);
}

/* commons/math/torus/mod.rs:72-78 (from_torus): fractional part centred on 0, scaled by 2^64,
 * rounded, cast to i64 (Rust `as` saturates), reinterpreted as u64.  Rounding mode: the x86 SIMD
 * path the reference runs on AVX2/AVX-512 hosts uses round-half-even (fft/x86.rs:864,
 * _MM_FROUND_NINT) while the scalar fallback uses f64::round (half away); ties have measure zero
 * for FFT outputs.  We use half-even (nearbyint), like the GPU's rint. */
inline uint64_t from_torus(double x) {
    double fract = x - std::nearbyint(x);
    fract *= 0x1p64;
    fract = std::nearbyint(fract);
    int64_t s;
    if (fract >= 0x1p63) s = INT64_MAX;
    else if (fract <= -0x1p63) s = INT64_MIN;
    else s = (int64_t)fract;
    return (uint64_t)s;
}

/* ------------------------------------------------------------------------------------ FFT */
struct FftPlan {
    uint32_t n = 0;               /* complex size = N/2 */
    std::vector<double> tw_re, tw_im; /* twisties: exp(i*pi*j/N), j < n  (fft/mod.rs:58-69) */
    std::vector<double> st_re, st_im; /* per-stage twiddles exp(-2*pi*i*t/len), packed */
    std::vector<uint32_t> brev;
    /* n == 1024 only: 32 x 32 four-step transform, vectorised across the 32 columns (see fft_forward_1024) */
    std::vector<double> w32_re, w32_im;   /* [stage 0..4][t < 16]: exp(-2 pi i t / (2 half)), half = 16 >> stage */
    std::vector<double> t4_re, t4_im;     /* [r < 32][j2 < 32]: exp(-2 pi i j2 brev5(r) / 1024) */
};

const FftPlan &get_plan(uint32_t N) {
    static std::mutex mu;
    static std::map<uint32_t, std::unique_ptr<FftPlan>> plans;
    std::lock_guard<std::mutex> lock(mu);
    auto it = plans.find(N);
    if (it != plans.end()) return *it->second;
    auto p = std::make_unique<FftPlan>();
    uint32_t n = N / 2;
    p->n = n;
    p->tw_re.resize(n); p->tw_im.resize(n);
    double unit = M_PI / (2.0 * (double)n); /* Twisties::new: unit = pi/(2n) */
    for (uint32_t i = 0; i < n; i++) {
        p->tw_re[i] = std::cos((double)i * unit);
        p->tw_im[i] = std::sin((double)i * unit);
    }
    for (uint32_t len = 2; len <= n; len <<= 1) {
        for (uint32_t t = 0; t < len / 2; t++) {
            /* use long double for the angle so table entries are correctly rounded doubles */
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)t / (long double)len;
            p->st_re.push_back((double)cosl(a));
            p->st_im.push_back((double)sinl(a));
        }
    }
    p->brev.resize(n);
    uint32_t lg = 0;
    while ((1u << lg) < n) lg++;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t r = 0;
        for (uint32_t b = 0; b < lg; b++) r |= ((i >> b) & 1u) << (lg - 1 - b);
        p->brev[i] = r;
    }
    if (n == 1024) {
        const long double pi = 3.14159265358979323846264338327950288L;
        p->w32_re.assign(5 * 16, 1.0); p->w32_im.assign(5 * 16, 0.0);
        for (int st = 0; st < 5; st++) {
            const int half = 16 >> st;
            for (int t = 0; t < half; t++) {
                long double a = -2.0L * pi * (long double)t / (long double)(2 * half);
                p->w32_re[st * 16 + t] = (double)cosl(a); p->w32_im[st * 16 + t] = (double)sinl(a);
            }
        }
        p->t4_re.resize(1024); p->t4_im.resize(1024);
        for (int r = 0; r < 32; r++) {
            int k1 = 0;
            for (int b = 0; b < 5; b++) k1 |= ((r >> b) & 1) << (4 - b);
            for (int j2 = 0; j2 < 32; j2++) {
                long double a = -2.0L * pi * (long double)(j2 * k1) / 1024.0L;
                p->t4_re[r * 32 + j2] = (double)cosl(a); p->t4_im[r * 32 + j2] = (double)sinl(a);
            }
        }
    }
    auto &ref = *p;
    plans[N] = std::move(p);
    return ref;
}

/* Forward transform: radix-2 decimation in frequency, natural order in, bit-reversed order out.
 * Inverse: radix-2 decimation in time with conjugated twiddles, bit-reversed in, natural out.
 * The Fourier-domain ordering is therefore bit-reversed; that is free (see file header): every
 * Fourier-domain operation on this path is pointwise. */
static inline size_t stage_off(uint32_t len) { return (size_t)len / 2 - 1; }

/* ---- n = 1024 (N = 2048, the headline parameter set): 32 x 32 four-step transform.  The reference gets its speed
 * from concrete-fft's SIMD kernels (fft/mod.rs:161: a plan measured at run time); this is the CPU baseline's
 * equivalent: every loop runs over the 32 contiguous columns of a row, so the compiler vectorises it (AVX2 /
 * AVX-512 clones selected at load time, see ORC_SIMD).  View the input as A[j1][j2] (j = 32 j1 + j2):
 *   pass 1  32-point DIF over j1 for all columns j2      -> row r holds k1 = brev5(r)
 *   twiddle times exp(-2 pi i j2 k1 / 1024), transpose   -> C[j2][r]
 *   pass 2  32-point DIF over j2 for all columns r       -> row s holds k2 = brev5(s): D[s][r] = X[k1 + 32 k2]
 * The inverse runs the same steps backwards (DIT, conjugated).  The Fourier-domain order is a fixed permutation,
 * which is free: every Fourier-domain operation on this path is pointwise. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define ORC_SIMD __attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
#else
#define ORC_SIMD
#endif

ORC_SIMD static void cols32_dif(double *__restrict re, double *__restrict im, const double *__restrict wr, const double *__restrict wi) {
    for (int st = 0, half = 16; half >= 1; st++, half >>= 1)
        for (int base = 0; base < 32; base += 2 * half)
            for (int t = 0; t < half; t++) {
                double *__restrict ar = re + (base + t) * 32, *__restrict ai = im + (base + t) * 32;
                double *__restrict br = ar + half * 32, *__restrict bi = ai + half * 32;
                const double cr = wr[st * 16 + t], ci = wi[st * 16 + t];
                for (int c = 0; c < 32; c++) {
                    const double dr = ar[c] - br[c], di = ai[c] - bi[c];
                    ar[c] += br[c]; ai[c] += bi[c];
                    br[c] = dr * cr - di * ci;
                    bi[c] = dr * ci + di * cr;
                }
            }
}
ORC_SIMD static void cols32_dit_conj(double *__restrict re, double *__restrict im, const double *__restrict wr, const double *__restrict wi) {
    for (int st = 4, half = 1; half <= 16; st--, half <<= 1)
        for (int base = 0; base < 32; base += 2 * half)
            for (int t = 0; t < half; t++) {
                double *__restrict ar = re + (base + t) * 32, *__restrict ai = im + (base + t) * 32;
                double *__restrict br = ar + half * 32, *__restrict bi = ai + half * 32;
                const double cr = wr[st * 16 + t], ci = wi[st * 16 + t];
                for (int c = 0; c < 32; c++) {
                    const double xr = br[c] * cr + bi[c] * ci, xi = bi[c] * cr - br[c] * ci;   /* b * conj(w) */
                    br[c] = ar[c] - xr; bi[c] = ai[c] - xi;
                    ar[c] += xr; ai[c] += xi;
                }
            }
}
/* out[j2][r] = in[r][j2] * (tr + i ti)[r][j2]   (CONJ: conj twiddle applied after the transpose: out[r][j2] = in[j2][r] * conj t[r][j2]) */
ORC_SIMD static void twiddle_transpose(const double *__restrict ir, const double *__restrict ii, double *__restrict orr, double *__restrict oi,
                                       const double *__restrict tr, const double *__restrict ti) {
    for (int r = 0; r < 32; r++)
        for (int c = 0; c < 32; c++) {
            const double a = ir[r * 32 + c], b = ii[r * 32 + c], x = tr[r * 32 + c], y = ti[r * 32 + c];
            orr[c * 32 + r] = a * x - b * y;
            oi[c * 32 + r] = a * y + b * x;
        }
}
ORC_SIMD static void transpose_twiddle_conj(const double *__restrict ir, const double *__restrict ii, double *__restrict orr, double *__restrict oi,
                                            const double *__restrict tr, const double *__restrict ti) {
    for (int r = 0; r < 32; r++)
        for (int c = 0; c < 32; c++) {
            const double a = ir[c * 32 + r], b = ii[c * 32 + r], x = tr[r * 32 + c], y = -ti[r * 32 + c];
            orr[r * 32 + c] = a * x - b * y;
            oi[r * 32 + c] = a * y + b * x;
        }
}
static void fft_forward_1024(const FftPlan &pl, double *re, double *im) {
    alignas(64) double tr[1024], ti[1024];
    cols32_dif(re, im, pl.w32_re.data(), pl.w32_im.data());
    twiddle_transpose(re, im, tr, ti, pl.t4_re.data(), pl.t4_im.data());
    cols32_dif(tr, ti, pl.w32_re.data(), pl.w32_im.data());
    std::memcpy(re, tr, sizeof(tr)); std::memcpy(im, ti, sizeof(ti));
}
static void fft_inverse_1024(const FftPlan &pl, double *re, double *im) {
    alignas(64) double tr[1024], ti[1024];
    cols32_dit_conj(re, im, pl.w32_re.data(), pl.w32_im.data());
    transpose_twiddle_conj(re, im, tr, ti, pl.t4_re.data(), pl.t4_im.data());
    cols32_dit_conj(tr, ti, pl.w32_re.data(), pl.w32_im.data());
    std::memcpy(re, tr, sizeof(tr)); std::memcpy(im, ti, sizeof(ti));
}

void fft_forward(const FftPlan &pl, double *__restrict re, double *__restrict im) {
    const uint32_t n = pl.n;
    if (n == 1024) return fft_forward_1024(pl, re, im);
    for (uint32_t len = n; len >= 2; len >>= 1) {
        const uint32_t half = len / 2;
        const double *__restrict wr = &pl.st_re[stage_off(len)], *__restrict wi = &pl.st_im[stage_off(len)];
        for (uint32_t base = 0; base < n; base += len) {
            double *__restrict ar = re + base, *__restrict ai = im + base;
            double *__restrict br = re + base + half, *__restrict bi = im + base + half;
            for (uint32_t t = 0; t < half; t++) {
                const double dr = ar[t] - br[t], di = ai[t] - bi[t];
                ar[t] += br[t]; ai[t] += bi[t];
                br[t] = dr * wr[t] - di * wi[t];
                bi[t] = dr * wi[t] + di * wr[t];
            }
        }
    }
}

void fft_inverse(const FftPlan &pl, double *__restrict re, double *__restrict im) {
    const uint32_t n = pl.n;
    if (n == 1024) return fft_inverse_1024(pl, re, im);
    for (uint32_t len = 2; len <= n; len <<= 1) {
        const uint32_t half = len / 2;
        const double *__restrict wr = &pl.st_re[stage_off(len)], *__restrict wi = &pl.st_im[stage_off(len)];
        for (uint32_t base = 0; base < n; base += len) {
            double *__restrict ar = re + base, *__restrict ai = im + base;
            double *__restrict br = re + base + half, *__restrict bi = im + base + half;
            for (uint32_t t = 0; t < half; t++) {
                const double xr = br[t] * wr[t] + bi[t] * wi[t];   /* b * conj(w) */
                const double xi = bi[t] * wr[t] - br[t] * wi[t];
                br[t] = ar[t] - xr; bi[t] = ai[t] - xi;
                ar[t] += xr; ai[t] += xi;
            }
        }
    }
}

struct FftScratch {
    std::vector<double> re, im;
    void ensure(uint32_t n) { if (re.size() < n) { re.resize(n); im.resize(n); } }
};
static thread_local FftScratch tl_fft;

/* The conversion loops around the transforms, as separate functions so that they get the SIMD clones too
 * (i64 <-> f64 conversions vectorise with AVX-512DQ; the reference does the same in fft/x86.rs:505+,823-874). */
ORC_SIMD static void twist_in(double *__restrict re, double *__restrict im, const uint64_t *__restrict poly, const double *__restrict twr,
                              const double *__restrict twi, uint32_t n, double scale) {
    for (uint32_t j = 0; j < n; j++) {
        const double a = (double)(int64_t)poly[j] * scale, b = (double)(int64_t)poly[j + n] * scale;
        re[j] = a * twr[j] - b * twi[j];
        im[j] = a * twi[j] + b * twr[j];
    }
}
ORC_SIMD static void interleave(double *__restrict out, const double *__restrict re, const double *__restrict im, uint32_t n) {
    for (uint32_t j = 0; j < n; j++) { out[2 * j] = re[j]; out[2 * j + 1] = im[j]; }
}
ORC_SIMD static void deinterleave(double *__restrict re, double *__restrict im, const double *__restrict in, uint32_t n) {
    for (uint32_t j = 0; j < n; j++) { re[j] = in[2 * j]; im[j] = in[2 * j + 1]; }
}
/* untwist, from_torus (same rounding and saturation as from_torus() above), wrapping add */
ORC_SIMD static void untwist_add(uint64_t *__restrict poly, const double *__restrict re, const double *__restrict im, const double *__restrict twr,
                                 const double *__restrict twi, uint32_t n, double norm) {
    for (uint32_t half = 0; half < 2; half++) {
        uint64_t *__restrict dst = poly + (size_t)half * n;
        for (uint32_t j = 0; j < n; j++) {
            const double wr = twr[j] * norm, wi = -twi[j] * norm;
            const double v = half == 0 ? re[j] * wr - im[j] * wi : re[j] * wi + im[j] * wr;
            double f = v - __builtin_nearbyint(v);
            f = __builtin_nearbyint(f * 0x1p64);
            const double c = f >= 0x1p63 ? 0.0 : f;                      /* keeps the conversion in range */
            const int64_t s = f >= 0x1p63 ? INT64_MAX : (int64_t)c;     /* Rust `as i64` saturates (-2^63 converts exactly) */
            dst[j] += (uint64_t)s;
        }
    }
}

/* Fourier-domain polynomials are kept SPLIT inside the oracle: n real parts, then n imaginary parts (the exported
 * orc_fft_* test entry points convert to / from interleaved complex).
 * fft/mod.rs:220-239 + 496-515: z_j = (i64(p_j) + i*i64(p_{j+N/2})) * w_j, then forward DFT */
void forward_integer(const FftPlan &pl, double *out /*split*/, const uint64_t *poly) {
    const uint32_t n = pl.n;
    twist_in(out, out + n, poly, pl.tw_re.data(), pl.tw_im.data(), n, 1.0);
    fft_forward(pl, out, out + n);
}

/* fft/mod.rs:197-218: same with inputs scaled by 2^-64 (key conversion) */
void forward_torus(const FftPlan &pl, double *out /*split*/, const uint64_t *poly) {
    const uint32_t n = pl.n;
    twist_in(out, out + n, poly, pl.tw_re.data(), pl.tw_im.data(), n, 0x1p-64);
    fft_forward(pl, out, out + n);
}

/* fft/mod.rs:285-304 + 539-557: inverse DFT, times conj(w_j)/n, from_torus, wrapping add.  Destroys `fourier`. */
void add_backward_torus(const FftPlan &pl, uint64_t *poly, double *fourier /*split*/) {
    const uint32_t n = pl.n;
    fft_inverse(pl, fourier, fourier + n);
    untwist_add(poly, fourier, fourier + n, pl.tw_re.data(), pl.tw_im.data(), n, 1.0 / (double)n);
}

/* out (+)= a * b, pointwise complex on split polynomials (update_with_fmadd, ggsw.rs:616-697) */
ORC_SIMD static void cmul_split(double *__restrict o, const double *__restrict a, const double *__restrict b, uint32_t n, bool accumulate) {
    double *__restrict orr = o, *__restrict oi = o + n;
    const double *__restrict ar = a, *__restrict ai = a + n, *__restrict br = b, *__restrict bi = b + n;
    if (accumulate)
        for (uint32_t q = 0; q < n; q++) { orr[q] += ar[q] * br[q] - ai[q] * bi[q]; oi[q] += ar[q] * bi[q] + ai[q] * br[q]; }
    else
        for (uint32_t q = 0; q < n; q++) { orr[q] = ar[q] * br[q] - ai[q] * bi[q]; oi[q] = ar[q] * bi[q] + ai[q] * br[q]; }
}

/* --------------------------------------------------------------------- integer primitives */
/* commons/math/decomposition/decomposer.rs:98-116 */
inline uint64_t closest_representable(uint64_t x, uint32_t base_log, uint32_t level) {
    uint32_t non_rep = 64 - base_log * level;
    uint32_t shift = non_rep - 1;
    uint64_t res = x >> shift;
    res += 1;
    res &= ~(uint64_t)1;
    return res << shift;
}

/* iter.rs:120-127 */
inline uint64_t decompose_one_level(uint32_t base_log, uint64_t &state, uint64_t mod_b_mask) {
    uint64_t res = state & mod_b_mask;
    state >>= base_log;
    uint64_t carry = ((res - 1) | state) & res;
    carry >>= (base_log - 1);
    state += carry;
    return res - (carry << base_log);
}

/* TensorSignedDecompositionLendingIter (fft64/math/decomposition.rs:26-86) over whole polynomials, vectorised */
ORC_SIMD static void init_states(uint64_t *__restrict st, const uint64_t *__restrict x, size_t len, uint32_t base_log, uint32_t level) {
    const uint32_t shift = 64 - base_log * level - 1;
    for (size_t j = 0; j < len; j++) st[j] = (((x[j] >> shift) + 1) & ~(uint64_t)1) << shift >> (shift + 1);
}
ORC_SIMD static void decompose_level(uint64_t *__restrict term, uint64_t *__restrict st, size_t len, uint32_t base_log, uint64_t mask) {
    for (size_t j = 0; j < len; j++) {
        const uint64_t res = st[j] & mask, hi = st[j] >> base_log;
        const uint64_t carry = (((res - 1) | hi) & res) >> (base_log - 1);
        st[j] = hi + carry;
        term[j] = res - (carry << base_log);
    }
}

/* fft_impl/common.rs:26-43 with offset 0, lut_count_log 0 */
inline uint64_t modulus_switch(uint64_t x, uint32_t log2N) {
    uint64_t out = x >> (64 - log2N - 2);
    out += 1;
    out >>= 1;
    return out;
}

/* algorithms/polynomial_algorithms.rs:315-354 */
void monomial_div(uint64_t *out, const uint64_t *in, size_t N, size_t degree) {
    size_t rem = degree % N, full = degree / N;
    bool neg_head = (full % 2) != 0;
    for (size_t j = 0; j < N - rem; j++) out[j] = neg_head ? (0 - in[rem + j]) : in[rem + j];
    for (size_t j = 0; j < rem; j++) out[N - rem + j] = neg_head ? in[j] : (0 - in[j]);
}

/* algorithms/polynomial_algorithms.rs:375-414 */
void monomial_mul(uint64_t *out, const uint64_t *in, size_t N, size_t degree) {
    size_t rem = degree % N, full = degree / N;
    bool flip = (full % 2) != 0;
    for (size_t j = 0; j < rem; j++) out[j] = flip ? in[N - rem + j] : (0 - in[N - rem + j]);
    for (size_t j = rem; j < N; j++) out[j] = flip ? (0 - in[j - rem]) : in[j - rem];
}

/* algorithms/polynomial_algorithms.rs:425-491 */
void monomial_mul_and_subtract(uint64_t *out, const uint64_t *in, size_t N, size_t degree) {
    size_t rem = degree % N, full = degree / N;
    bool flip = (full % 2) != 0;
    for (size_t j = 0; j < rem; j++) {
        uint64_t src = in[N - rem + j];
        out[j] = (flip ? src : (0 - src)) - in[j];
    }
    for (size_t j = rem; j < N; j++) {
        uint64_t src = in[j - rem];
        out[j] = (flip ? (0 - src) : src) - in[j];
    }
}

/* algorithms/glwe_sample_extraction.rs:91-147 with nth = 0 */
void sample_extract0(uint64_t *lwe, const uint64_t *glwe, uint32_t k, uint32_t N) {
    lwe[(size_t)k * N] = glwe[(size_t)k * N];
    for (uint32_t p = 0; p < k; p++) {
        const uint64_t *a = glwe + (size_t)p * N;
        uint64_t *m = lwe + (size_t)p * N;
        m[0] = a[0];
        for (uint32_t j = 1; j < N; j++) m[j] = 0 - a[N - j];
    }
}

/* exact negacyclic product by a binary polynomial (key generation only) */
void negacyclic_mul_binary_add(uint64_t *out, const uint64_t *a, const uint64_t *s_bits, uint32_t N) {
    for (uint32_t d = 0; d < N; d++) {
        if (!s_bits[d]) continue;
        for (uint32_t j = 0; j < N - d; j++) out[j + d] += a[j];
        for (uint32_t j = N - d; j < N; j++) out[j + d - N] -= a[j];
    }
}

} // namespace

/* ======================================================================= keyset ========== */
struct orc_keyset {
    orc_params p;
    std::vector<uint64_t> small_sk, big_sk, ksk, bsk_std;
    std::vector<double> bsk_fourier;
    uint64_t enc_counter = 0;
    uint64_t seed = 0;
};

extern "C" {

void orc_params_message_2_carry_2(orc_params *p) {
    p->lwe_dimension = 742; p->glwe_dimension = 1; p->polynomial_size = 2048;
    p->lwe_modular_std_dev = 0.000007069849454709433;
    p->glwe_modular_std_dev = 0.00000000000000029403601535432533;
    p->pbs_base_log = 23; p->pbs_level = 1; p->ks_base_log = 3; p->ks_level = 5;
    p->message_modulus = 4; p->carry_modulus = 4;
}

void orc_params_toy(orc_params *p, uint32_t n, uint32_t N) {
    orc_params_message_2_carry_2(p);
    p->lwe_dimension = n; p->polynomial_size = N;
}

static void lwe_encrypt_into(const uint64_t *sk, uint32_t n, uint64_t plaintext, double std, Rng &rng,
                             uint64_t *out) {
    /* algorithms/lwe_encryption.rs:61-105: mask uniform, body = <a,s> + noise + plaintext */
    uint64_t acc = 0;
    for (uint32_t i = 0; i < n; i++) { out[i] = rng.next(); acc += out[i] * sk[i]; }
    double g0, g1;
    rng.gaussian_pair(std, g0, g1);
    out[n] = acc + from_torus(g0) + plaintext;
}

orc_keyset *orc_keyset_create(const orc_params *pp, uint64_t seed, int n_threads) {
    auto *ks = new orc_keyset();
    ks->p = *pp; ks->seed = seed;
    const orc_params &p = ks->p;
    const uint32_t n = p.lwe_dimension, k = p.glwe_dimension, N = p.polynomial_size;
    const uint32_t big = k * N;
    if (n_threads < 1) n_threads = 1;
    {   /* binary secret keys (allocate_and_generate_new_binary_*_secret_key) */
        Rng r(derive_seed(seed, 1, 0));
        ks->small_sk.resize(n);
        for (auto &b : ks->small_sk) b = r.next() >> 63;
        ks->big_sk.resize(big);
        for (auto &b : ks->big_sk) b = r.next() >> 63;
    }
    /* KSK: algorithms/lwe_keyswitch_key_generation.rs:65-130. Block i = encryptions under the
     * small key of s_big[i] * 2^(64 - base_log*lvl) for lvl = level..1 (that storage order). */
    const uint32_t out_size = n + 1;
    ks->ksk.resize((size_t)big * p.ks_level * out_size);
    {
        auto work = [&](uint32_t t) {
            for (uint32_t i = t; i < big; i += (uint32_t)n_threads) {
                Rng r(derive_seed(seed, 2, i));
                for (uint32_t li = 0; li < p.ks_level; li++) {
                    uint32_t lvl = p.ks_level - li;
                    uint64_t pt = ks->big_sk[i] << (64 - p.ks_base_log * lvl);
                    lwe_encrypt_into(ks->small_sk.data(), n, pt, p.lwe_modular_std_dev, r,
                                     &ks->ksk[((size_t)i * p.ks_level + li) * out_size]);
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work, (uint32_t)t);
        for (auto &x : th) x.join();
    }
    /* BSK: algorithms/lwe_bootstrap_key_generation.rs:76-141 -> one constant GGSW per small-key
     * bit; algorithms/ggsw_encryption.rs:72-151,300-331: level matrices stored level 1..l,
     * factor = -m * 2^(64 - base_log*lvl); row r<k body = factor * S_r(X); last row body[0] =
     * -factor; then GLWE encryption in place (glwe_encryption.rs: body += sum_i A_i*S_i + e). */
    const size_t ggsw_len = (size_t)p.pbs_level * (k + 1) * (k + 1) * N;
    ks->bsk_std.assign((size_t)n * ggsw_len, 0);
    {
        auto work = [&](uint32_t t) {
            std::vector<uint64_t> body(N);
            for (uint32_t i = t; i < n; i += (uint32_t)n_threads) {
                Rng r(derive_seed(seed, 3, i));
                uint64_t m = ks->small_sk[i];
                for (uint32_t li = 0; li < p.pbs_level; li++) {
                    uint32_t lvl = li + 1;
                    uint64_t factor = (0 - m) * ((uint64_t)1 << (64 - p.pbs_base_log * lvl));
                    for (uint32_t row = 0; row <= k; row++) {
                        uint64_t *glwe = &ks->bsk_std[(size_t)i * ggsw_len + ((size_t)li * (k + 1) + row) * (k + 1) * N];
                        uint64_t *bd = glwe + (size_t)k * N;
                        if (row < k) for (uint32_t j = 0; j < N; j++) bd[j] = ks->big_sk[(size_t)row * N + j] * factor;
                        else { std::fill(bd, bd + N, 0); bd[0] = 0 - factor; }
                        for (uint32_t q = 0; q < k; q++) {
                            uint64_t *a = glwe + (size_t)q * N;
                            for (uint32_t j = 0; j < N; j++) a[j] = r.next();
                            negacyclic_mul_binary_add(bd, a, &ks->big_sk[(size_t)q * N], N);
                        }
                        for (uint32_t j = 0; j < N; j += 2) {
                            double g0, g1;
                            r.gaussian_pair(p.glwe_modular_std_dev, g0, g1);
                            bd[j] += from_torus(g0);
                            if (j + 1 < N) bd[j + 1] += from_torus(g1);
                        }
                    }
                }
            }
        };
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work, (uint32_t)t);
        for (auto &x : th) x.join();
    }
    /* std -> Fourier: fft/mod.rs:719-764 (forward_as_torus on every polynomial) */
    const FftPlan &pl = get_plan(N);
    const size_t n_polys = (size_t)n * p.pbs_level * (k + 1) * (k + 1);
    ks->bsk_fourier.resize(n_polys * N); /* N/2 complex = N doubles per poly */
    {
        auto work = [&](uint32_t t) {
            for (size_t q = t; q < n_polys; q += (size_t)n_threads)
                forward_torus(pl, &ks->bsk_fourier[q * N], &ks->bsk_std[q * N]);
        };
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work, (uint32_t)t);
        for (auto &x : th) x.join();
    }
    return ks;
}

void orc_keyset_destroy(orc_keyset *ks) { delete ks; }
const orc_params *orc_keyset_params(const orc_keyset *ks) { return &ks->p; }
const uint64_t *orc_keyset_small_sk(const orc_keyset *ks) { return ks->small_sk.data(); }
const uint64_t *orc_keyset_big_sk(const orc_keyset *ks) { return ks->big_sk.data(); }
const uint64_t *orc_keyset_ksk(const orc_keyset *ks) { return ks->ksk.data(); }
size_t orc_keyset_ksk_len(const orc_keyset *ks) { return ks->ksk.size(); }
const uint64_t *orc_keyset_bsk_standard(const orc_keyset *ks) { return ks->bsk_std.data(); }
size_t orc_keyset_bsk_len(const orc_keyset *ks) { return ks->bsk_std.size(); }
const double *orc_keyset_bsk_fourier(const orc_keyset *ks) { return ks->bsk_fourier.data(); }

/* ------------------------------------------------------------------ exported primitives */
uint64_t orc_closest_representable(uint64_t x, uint32_t base_log, uint32_t level) {
    return closest_representable(x, base_log, level);
}

void orc_decompose(uint64_t x, uint32_t base_log, uint32_t level, int64_t *digits) {
    /* decomposer.rs:144-152 + iter.rs:37-50 */
    uint64_t state = closest_representable(x, base_log, level) >> (64 - base_log * level);
    uint64_t mask = ((uint64_t)1 << base_log) - 1;
    for (uint32_t i = 0; i < level; i++) digits[i] = (int64_t)decompose_one_level(base_log, state, mask);
}

uint64_t orc_modulus_switch(uint64_t x, uint32_t log2N) { return modulus_switch(x, log2N); }
void orc_monomial_div(uint64_t *o, const uint64_t *i, size_t N, size_t d) { monomial_div(o, i, N, d); }
void orc_monomial_mul(uint64_t *o, const uint64_t *i, size_t N, size_t d) { monomial_mul(o, i, N, d); }
void orc_monomial_mul_and_subtract(uint64_t *o, const uint64_t *i, size_t N, size_t d) {
    monomial_mul_and_subtract(o, i, N, d);
}
void orc_sample_extract0(uint64_t *lwe, const uint64_t *glwe, uint32_t k, uint32_t N) {
    sample_extract0(lwe, glwe, k, N);
}
void orc_fft_forward_integer(double *f, const uint64_t *poly, uint32_t N) {
    std::vector<double> t(N);
    forward_integer(get_plan(N), t.data(), poly);
    interleave(f, t.data(), t.data() + N / 2, N / 2);
}
void orc_fft_forward_torus(double *f, const uint64_t *poly, uint32_t N) {
    std::vector<double> t(N);
    forward_torus(get_plan(N), t.data(), poly);
    interleave(f, t.data(), t.data() + N / 2, N / 2);
}
void orc_fft_add_backward_torus(uint64_t *poly, const double *f, uint32_t N) {
    std::vector<double> t(N);
    deinterleave(t.data(), t.data() + N / 2, f, N / 2);
    add_backward_torus(get_plan(N), poly, t.data());
}
uint64_t orc_from_torus(double x) { return from_torus(x); }

/* ---------------------------------------------------------------------------- keyswitch */
/* algorithms/lwe_keyswitch.rs:96-170 + slice_algorithms.rs:363-462 */
/* slice_wrapping_sub_scalar_mul_assign, algorithms/slice_algorithms.rs:363-462 */
ORC_SIMD static void sub_scalar_mul(uint64_t *__restrict out, const uint64_t *__restrict row, uint64_t d, uint32_t len) {
    for (uint32_t j = 0; j < len; j++) out[j] -= row[j] * d;
}

void orc_keyswitch_raw(const uint64_t *ksk, uint32_t in_dim, uint32_t out_dim, uint32_t base_log,
                       uint32_t level, const uint64_t *in, uint64_t *out) {
    const uint32_t out_size = out_dim + 1;
    std::fill(out, out + out_size, 0);
    out[out_dim] = in[in_dim];
    const uint64_t mask = ((uint64_t)1 << base_log) - 1;
    for (uint32_t i = 0; i < in_dim; i++) {
        uint64_t state = closest_representable(in[i], base_log, level) >> (64 - base_log * level);
        for (uint32_t li = 0; li < level; li++) {
            uint64_t d = decompose_one_level(base_log, state, mask);
            const uint64_t *row = ksk + ((size_t)i * level + li) * out_size;
            sub_scalar_mul(out, row, d, out_size);
        }
    }
}

void orc_keyswitch(const orc_keyset *ks, const uint64_t *in_big, uint64_t *out_small) {
    const orc_params &p = ks->p;
    orc_keyswitch_raw(ks->ksk.data(), p.glwe_dimension * p.polynomial_size, p.lwe_dimension, p.ks_base_log,
                      p.ks_level, in_big, out_small);
}

/* ------------------------------------------------------------------- external product */
/* fft_impl/fft64/crypto/ggsw.rs:477-598.  acc += ggsw (x) glwe.  Scratch passed in. */
struct Scratch {
    std::vector<uint64_t> ct1, states, term;
    std::vector<double> fourier, out_f;
    Scratch(uint32_t k, uint32_t N) : ct1((size_t)(k + 1) * N), states((size_t)(k + 1) * N), term(N), fourier(N), out_f((size_t)(k + 1) * N) {}
};

static void add_external_product(const orc_params &p, const FftPlan &pl, uint64_t *out, const double *ggsw_f,
                                 const uint64_t *glwe, Scratch &s) {
    const uint32_t k = p.glwe_dimension, N = p.polynomial_size, bl = p.pbs_base_log, lv = p.pbs_level;
    const size_t glwe_len = (size_t)(k + 1) * N;
    const uint64_t mask = ((uint64_t)1 << bl) - 1;
    /* TensorSignedDecompositionLendingIter::new (fft64/math/decomposition.rs:26-46) */
    init_states(s.states.data(), glwe, glwe_len, bl, lv);
    bool uninit = true;
    for (uint32_t li = 0; li < lv; li++) {
        /* levels come out l, l-1, ..., 1; GGSW level matrices are stored 1..l and iterated .rev() */
        uint32_t lvl = lv - li;
        const double *mat = ggsw_f + (size_t)(lvl - 1) * (k + 1) * (k + 1) * N; /* N doubles per poly */
        for (uint32_t row = 0; row <= k; row++) {
            decompose_level(s.term.data(), &s.states[(size_t)row * N], N, bl, mask);
            forward_integer(pl, s.fourier.data(), s.term.data());
            const double *rowp = mat + (size_t)row * (k + 1) * N;
            /* update_with_fmadd, ggsw.rs:616-697: first write is a mul, later ones fused mul-add */
            for (uint32_t col = 0; col <= k; col++)
                cmul_split(&s.out_f[(size_t)col * N], rowp + (size_t)col * N, s.fourier.data(), N / 2, !uninit);
            uninit = false;
        }
    }
    for (uint32_t col = 0; col <= k; col++) add_backward_torus(pl, out + (size_t)col * N, &s.out_f[(size_t)col * N]);
}

/* fft_impl/fft64/crypto/bootstrap.rs:242-331 */
/* n_steps < n stops the CMUX loop early (test hook: one external product on identical inputs pins the Fourier stage
 * without the decomposer random walk); the body is then read at lwe[n_steps]. */
static void blind_rotate(const orc_keyset *ks, const uint64_t *lwe, uint64_t *acc, Scratch &s, uint32_t n_steps = UINT32_MAX) {
    const orc_params &p = ks->p;
    const uint32_t n = std::min(p.lwe_dimension, n_steps), k = p.glwe_dimension, N = p.polynomial_size;
    uint32_t log2N = 0;
    while ((1u << log2N) < N) log2N++;
    const FftPlan &pl = get_plan(N);
    const size_t ggsw_f_len = (size_t)p.pbs_level * (k + 1) * (k + 1) * N; /* doubles */
    uint64_t b_hat = modulus_switch(lwe[n], log2N);
    std::vector<uint64_t> tmp(N);
    for (uint32_t q = 0; q <= k; q++) {
        std::copy(acc + (size_t)q * N, acc + (size_t)(q + 1) * N, tmp.begin());
        monomial_div(acc + (size_t)q * N, tmp.data(), N, b_hat);
    }
    for (uint32_t i = 0; i < n; i++) {
        if (lwe[i] == 0) continue;
        uint64_t a_hat = modulus_switch(lwe[i], log2N);
        for (uint32_t q = 0; q <= k; q++)
            monomial_mul_and_subtract(&s.ct1[(size_t)q * N], acc + (size_t)q * N, N, a_hat);
        add_external_product(p, pl, acc, &ks->bsk_fourier[(size_t)i * ggsw_f_len], s.ct1.data(), s);
    }
}

void orc_blind_rotate(const orc_keyset *ks, const uint64_t *lwe_small, uint64_t *acc) {
    Scratch s(ks->p.glwe_dimension, ks->p.polynomial_size);
    blind_rotate(ks, lwe_small, acc, s);
}

/* bootstrap.rs:333-364 */
static void bootstrap(const orc_keyset *ks, const uint64_t *lwe_small, const uint64_t *lut, uint64_t *out_big,
                      Scratch &s) {
    const uint32_t k = ks->p.glwe_dimension, N = ks->p.polynomial_size;
    std::vector<uint64_t> acc(lut, lut + (size_t)(k + 1) * N);
    blind_rotate(ks, lwe_small, acc.data(), s);
    sample_extract0(out_big, acc.data(), k, N);
}

void orc_bootstrap(const orc_keyset *ks, const uint64_t *lwe_small, const uint64_t *lut, uint64_t *out_big) {
    Scratch s(ks->p.glwe_dimension, ks->p.polynomial_size);
    bootstrap(ks, lwe_small, lut, out_big, s);
}

void orc_bootstrap_steps(const orc_keyset *ks, const uint64_t *lwe_prefix, uint32_t n_steps, const uint64_t *lut, uint64_t *out_big) {
    const uint32_t k = ks->p.glwe_dimension, N = ks->p.polynomial_size;
    Scratch s(k, N);
    std::vector<uint64_t> acc(lut, lut + (size_t)(k + 1) * N);
    blind_rotate(ks, lwe_prefix, acc.data(), s, n_steps);
    sample_extract0(out_big, acc.data(), k, N);
}

/* shortint/server_key/mod.rs:783-857 (classic branch, non trivial input) */
void orc_ks_pbs(const orc_keyset *ks, const uint64_t *in_big, const uint64_t *lut, uint64_t *out_big) {
    std::vector<uint64_t> small(ks->p.lwe_dimension + 1);
    orc_keyswitch(ks, in_big, small.data());
    orc_bootstrap(ks, small.data(), lut, out_big);
}

/* the reference fans independent ciphertexts over rayon workers (benches/core_crypto/
 * pbs_bench.rs:517-531); same here with std::thread */
void orc_ks_pbs_batch(const orc_keyset *ks, const uint64_t *in_big, const uint64_t *luts, const uint32_t *lut_idx,
                      uint64_t *out_big, size_t batch, int n_threads) {
    const orc_params &p = ks->p;
    const size_t big_size = (size_t)p.glwe_dimension * p.polynomial_size + 1;
    const size_t lut_len = (size_t)(p.glwe_dimension + 1) * p.polynomial_size;
    if (n_threads < 1) n_threads = 1;
    auto work = [&](size_t t) {
        Scratch s(p.glwe_dimension, p.polynomial_size);
        std::vector<uint64_t> small(p.lwe_dimension + 1);
        for (size_t b = t; b < batch; b += (size_t)n_threads) {
            orc_keyswitch(ks, in_big + b * big_size, small.data());
            bootstrap(ks, small.data(), luts + (size_t)(lut_idx ? lut_idx[b] : 0) * lut_len, out_big + b * big_size, s);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work, (size_t)t);
    for (auto &x : th) x.join();
}

/* ------------------------------------------------------------------------- shortint */
/* shortint/engine/mod.rs:72-128 */
uint64_t orc_fill_accumulator(const orc_params *p, const uint64_t *table, uint64_t *glwe_out) {
    const uint32_t k = p->glwe_dimension, N = p->polynomial_size;
    const uint64_t modulus_sup = (uint64_t)p->message_modulus * p->carry_modulus;
    const size_t box = N / modulus_sup;
    const uint64_t delta = ((uint64_t)1 << 63) / modulus_sup;
    std::fill(glwe_out, glwe_out + (size_t)k * N, 0);
    uint64_t *body = glwe_out + (size_t)k * N;
    uint64_t maxv = 0;
    for (uint64_t i = 0; i < modulus_sup; i++) {
        uint64_t f = table[i];
        maxv = std::max(maxv, f);
        for (size_t j = 0; j < box; j++) body[i * box + j] = f * delta;
    }
    const size_t half = box / 2;
    for (size_t j = 0; j < half; j++) body[j] = 0 - body[j];
    std::rotate(body, body + half, body + N);
    return maxv;
}

/* shortint/server_key/mod.rs:763-781 */
uint64_t orc_trivial_pbs(const orc_params *p, uint64_t body, const uint64_t *lut_glwe) {
    const uint64_t modulus_sup = (uint64_t)p->message_modulus * p->carry_modulus;
    const uint64_t delta = ((uint64_t)1 << 63) / modulus_sup;
    const size_t box = p->polynomial_size / modulus_sup;
    const uint64_t *lb = lut_glwe + (size_t)p->glwe_dimension * p->polynomial_size;
    uint64_t v = body / delta;
    if (v >= modulus_sup) return 0 - lb[(v % modulus_sup) * box];
    return lb[v * box];
}

/* shortint/engine/client_side.rs:58-85 with message_modulus = full message*carry space
 * (unchecked range, cf. client_side.rs:231), big key, glwe noise */
void orc_encrypt_seeded(const orc_keyset *ks, uint64_t m, uint64_t seed, uint64_t *out_big) {
    const orc_params &p = ks->p;
    const uint64_t modulus_sup = (uint64_t)p.message_modulus * p.carry_modulus;
    const uint64_t delta = ((uint64_t)1 << 63) / modulus_sup;
    Rng r(seed);
    lwe_encrypt_into(ks->big_sk.data(), p.glwe_dimension * p.polynomial_size, (m % (2 * modulus_sup)) * delta,
                     p.glwe_modular_std_dev, r, out_big);
}

void orc_encrypt(orc_keyset *ks, uint64_t m, uint64_t *out_big) {
    orc_encrypt_seeded(ks, m, derive_seed(ks->seed, 4, ks->enc_counter++), out_big);
}

void orc_encrypt_batch_seeded(const orc_keyset *ks, const uint64_t *messages, size_t batch, uint64_t seed,
                              uint64_t *out_big) {
    const size_t big_size = (size_t)ks->p.glwe_dimension * ks->p.polynomial_size + 1;
    for (size_t b = 0; b < batch; b++)
        orc_encrypt_seeded(ks, messages[b], derive_seed(seed, 5, b), out_big + b * big_size);
}

/* algorithms/lwe_encryption.rs:520-545 (decrypt_lwe_ciphertext): body - <mask, key> */
uint64_t orc_decrypt_phase(const orc_keyset *ks, const uint64_t *ct) {
    const size_t dim = (size_t)ks->p.glwe_dimension * ks->p.polynomial_size;
    uint64_t acc = 0;
    for (size_t i = 0; i < dim; i++) acc += ct[i] * ks->big_sk[i];
    return ct[dim] - acc;
}

uint64_t orc_decrypt_small_phase(const orc_keyset *ks, const uint64_t *ct) {
    const size_t dim = ks->p.lwe_dimension;
    uint64_t acc = 0;
    for (size_t i = 0; i < dim; i++) acc += ct[i] * ks->small_sk[i];
    return ct[dim] - acc;
}

/* shortint/client_key/mod.rs:281-302 */
uint64_t orc_decrypt_message_and_carry(const orc_keyset *ks, const uint64_t *ct) {
    const uint64_t modulus_sup = (uint64_t)ks->p.message_modulus * ks->p.carry_modulus;
    const uint64_t delta = ((uint64_t)1 << 63) / modulus_sup;
    uint64_t ph = orc_decrypt_phase(ks, ct);
    uint64_t rounding = (ph & (delta >> 1)) << 1;
    return (ph + rounding) / delta;
}

void orc_decrypt_batch(const orc_keyset *ks, const uint64_t *cts, size_t batch, uint64_t *out) {
    const size_t big_size = (size_t)ks->p.glwe_dimension * ks->p.polynomial_size + 1;
    for (size_t b = 0; b < batch; b++) out[b] = orc_decrypt_message_and_carry(ks, cts + b * big_size);
}

} // extern "C"
