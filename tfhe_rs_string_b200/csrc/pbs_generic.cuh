// pbs_generic.cuh -- programmable bootstrap for ANY classic parameter set (glwe_dimension k >= 1, polynomial
// size N = 256 .. 32768, any decomposition base / level count), templated on the torus word (u64 shortint,
// u32 boolean).  One CTA per ciphertext; the specialised kernels (pbs_kernel5.cuh, pbs_kernel_lat.cuh) take over
// for k = 1, N = 2048, one level of base 2^23 -- this file is the coverage path for the other sets of
// shortint/parameters/mod.rs (PARAM_MESSAGE_1_CARRY_1 k=3 N=512 :613-627, 3_3 N=8192 l=2 :853-867, 4_4 N=32768
// :1063-1077) and boolean/parameters/mod.rs:123-192.
//
// Per ciphertext it restates FourierLweBootstrapKeyView::bootstrap (core_crypto/fft_impl/fft64/crypto/
// bootstrap.rs:242-364) with the multi-level external product of ggsw.rs:477-598:
//   acc = LUT * X^-b~;  for i < n:  ct1 = acc * X^a~_i - acc;
//       for every GLWE polynomial r and level l: Out_c += BSK[i][l][r][c] * FFT(digit_l(ct1_r))   (c <= k)
//       acc_c += from_torus(IFFT(Out_c))
// then sample extraction.  The accumulator and the k+1 Fourier accumulators live in a per-ciphertext global
// workspace (L2 resident); the FFT runs in shared memory when N/2 complex doubles fit (N <= 16384), else in the
// workspace.  As many of the (k+1) * level forward transforms of a CMUX step as fit in shared memory run side by side (one
// block barrier per butterfly stage serves all of them; the k+1 inverse transforms likewise): a single bootstrap is bound
// by the number of block barriers per step, not by arithmetic.  Transform: in-place radix-2, forward decimation in frequency (natural in, bit-reversed out), inverse
// decimation in time; the Fourier key is produced by the same forward transform (bsk_to_fourier_generic_kernel),
// so the frequency order is private to this file, and pre-scaled by 2/N.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

struct GenPbsArgs {
    const void *lwe_small;     // [batch][n + 1] torus words
    const uint32_t *lut_idx;   // [batch] or nullptr
    const void *luts;          // [n_luts][k + 1][N]
    const double2 *bsk;        // [n][level][k + 1][k + 1][N / 2], level index 0 = level 1 (ggsw_encryption.rs:116-119)
    const double2 *roots;      // [N / 4] exp(-2 pi i t / (N / 2))
    const double2 *twist;      // [N / 2] exp(i pi j / N)   (fft/mod.rs:58-69)
    void *acc_ws;              // [batch][k + 1][N] torus words
    double2 *fourier_ws;       // [batch][k + 1 (+1 when the FFT buffer does not fit shared memory)][N / 2]
    void *out;                 // [batch][k N + 1]
    int batch, n, k, log2N, base_log, level;
    int fft_in_smem;
    int polys_in_smem;         // FFT buffers (N / 2 complex doubles each) the launch reserved in shared memory, >= 1 when fft_in_smem
    uint32_t n_luts;           // ids >= n_luts are rejected in the kernel (table 0 used, *err_flag set)
    uint32_t *err_flag;
};

// Launch shape of pbs_generic_kernel: as many FFT buffers as the (k+1) * level forward transforms of a step can use and
// shared memory holds (next to the N/4 roots of unity), one butterfly per thread and stage of a group up to 1024 threads
// (a block barrier over 4 warps costs a third of one over 16, so small shapes stay small).
struct GenLaunch { int polys_in_smem; size_t smem; unsigned threads; };
inline GenLaunch gen_launch_shape(const uint32_t N, const uint32_t k, const uint32_t level, const bool fft_in_smem, const size_t smem_limit) {
    GenLaunch gl{1, 0, 0};
    const uint32_t jobs = (k + 1) * level;
    if (fft_in_smem) {
        const size_t roots = (size_t)(N / 4) * sizeof(double2), per = (size_t)(N / 2) * sizeof(double2);
        const size_t fit = (smem_limit - 1024 - roots) / per;
        gl.polys_in_smem = (int)(fit < 1 ? 1 : fit > jobs ? jobs : fit);
        gl.smem = roots + (size_t)gl.polys_in_smem * per;
    }
    const uint32_t rounds = (jobs + gl.polys_in_smem - 1) / gl.polys_in_smem, grp = (jobs + rounds - 1) / rounds;
    const uint32_t want = grp * (N / 4);
    gl.threads = want < 128u ? 128u : want > 1024u ? 1024u : want;
    return gl;
}

template <typename Torus> struct TorusTraits;
template <> struct TorusTraits<uint64_t> { static constexpr int bits = 64; typedef int64_t Signed; };
template <> struct TorusTraits<uint32_t> { static constexpr int bits = 32; typedef int32_t Signed; };

// closest_representable (decomposer.rs:98-116) then the balanced digit of level `lvl` (1 = most significant),
// by running decompose_one_level (iter.rs:120-127) from level `level` up to `lvl`.
template <typename Torus>
__device__ __forceinline__ double gen_digit(Torus v, const int base_log, const int level, const int lvl) {
    constexpr int bits = TorusTraits<Torus>::bits;
    const int non_rep = bits - base_log * level;
    Torus state = v >> (non_rep - 1);
    state = (Torus)(state + 1) >> 1;                        // rounded, already shifted down by non_rep
    if (non_rep + base_log * level < bits) {}               // (never: kept for symmetry with the reference)
    const Torus mask = ((Torus)1 << base_log) - 1;
    Torus digit = 0;
    for (int li = level; li >= lvl; li--) {
        Torus res = state & mask;
        state >>= base_log;
        const Torus carry = ((Torus)((res - 1) | state) & res) >> (base_log - 1);
        state += carry;
        digit = res - (carry << base_log);
    }
    return (double)(typename TorusTraits<Torus>::Signed)digit;
}

// acc * X^deg coefficient j for deg in [0, 2N]: polynomial_wrapping_monic_monomial_mul (polynomial_algorithms.rs:375-414)
template <typename Torus>
__device__ __forceinline__ Torus gen_rot(const Torus *poly, const int j, const int deg, const int N) {
    const int idx = (j - deg) & (2 * N - 1);              // source index in the 2N-periodic negacyclic extension
    const Torus v = poly[idx & (N - 1)];
    return (idx & N) ? (Torus)(0 - v) : v;
}

__device__ __forceinline__ double2 gen_cmul(const double2 a, const double2 b) {
    return make_double2(fma(-a.y, b.y, a.x * b.x), fma(a.y, b.x, a.x * b.y));
}

// in-place transforms over `cnt` contiguous buffers of n = N/2 complex points each, all threads of the CTA cooperate
// (butterfly b of the concatenated array belongs to transform b / (n/2): the index arithmetic below needs no change)
__device__ __forceinline__ void gen_fft_forward(double2 *buf, const double2 *__restrict__ roots, const int n, const int cnt = 1) {
    for (int half = n >> 1, stride = 1; half >= 1; half >>= 1, stride <<= 1) {
        __syncthreads();
        for (int b = threadIdx.x; b < cnt * (n >> 1); b += blockDim.x) {
            const int t = b & (half - 1), base = (b - t) << 1;
            const double2 x = buf[base + t], y = buf[base + t + half];
            const double2 w = roots[t * stride];
            buf[base + t] = make_double2(x.x + y.x, x.y + y.y);
            buf[base + t + half] = gen_cmul(make_double2(x.x - y.x, x.y - y.y), w);
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void gen_fft_inverse(double2 *buf, const double2 *__restrict__ roots, const int n, const int cnt = 1) {
    for (int half = 1, stride = n >> 1; half < n; half <<= 1, stride >>= 1) {
        __syncthreads();
        for (int b = threadIdx.x; b < cnt * (n >> 1); b += blockDim.x) {
            const int t = b & (half - 1), base = (b - t) << 1;
            const double2 w = roots[t * stride];
            const double2 x = buf[base + t], y = gen_cmul(buf[base + t + half], make_double2(w.x, -w.y));
            buf[base + t] = make_double2(x.x + y.x, x.y + y.y);
            buf[base + t + half] = make_double2(x.x - y.x, x.y - y.y);
        }
    }
    __syncthreads();
}

template <typename Torus>
__device__ __forceinline__ Torus gen_from_torus(const double x) {      // torus/mod.rs:72-78, half-even like x86.rs:864
    constexpr int bits = TorusTraits<Torus>::bits;
    const double f = x - rint(x);
    return (Torus)(unsigned long long)__double2ll_rn(f * (bits == 64 ? 18446744073709551616.0 : 4294967296.0));
}

template <typename Torus>
__global__ void __launch_bounds__(1024, 1) pbs_generic_kernel(const GenPbsArgs a) {
    extern __shared__ __align__(16) unsigned char gen_smem[];
    constexpr int bits = TorusTraits<Torus>::bits;
    const int ct = blockIdx.x;
    if (ct >= a.batch) return;
    const int N = 1 << a.log2N, half = N >> 1, k1 = a.k + 1;
    const Torus *lwe = (const Torus *)a.lwe_small + (size_t)ct * (a.n + 1);
    uint32_t lid = a.lut_idx ? a.lut_idx[ct] : 0u;
    if (lid >= a.n_luts) {
        if (a.err_flag && threadIdx.x == 0) atomicOr(a.err_flag, 1u);
        lid = 0;
    }
    const Torus *lut = (const Torus *)a.luts + (size_t)lid * k1 * N;
    Torus *acc = (Torus *)a.acc_ws + (size_t)ct * k1 * N;
    const int n_f = k1 + (a.fft_in_smem ? 0 : 1);
    double2 *outf = a.fourier_ws + (size_t)ct * n_f * half;
    double2 *buf = a.fft_in_smem ? reinterpret_cast<double2 *>(gen_smem) : outf + (size_t)k1 * half;
    const int ms_shift = bits - a.log2N - 2;                             // fast_pbs_modulus_switch, common.rs:26-43
    // the N/4 roots of unity are read with power-of-two strides by every butterfly: keep them next to the buffer in shared
    // memory (a strided global load per butterfly was most of the time of a CMUX step for N = 8192)
    const double2 *roots = a.roots;
    if (a.fft_in_smem) {
        double2 *rs = reinterpret_cast<double2 *>(gen_smem) + (size_t)max(1, a.polys_in_smem) * half;
        for (int t = threadIdx.x; t < (N >> 2); t += blockDim.x) rs[t] = a.roots[t];
        roots = rs;
        __syncthreads();
    }

    // acc = LUT * X^-b~ (polynomial_wrapping_monic_monomial_div, polynomial_algorithms.rs:315-354)
    const int bhat = (int)((((lwe[a.n] >> ms_shift) + 1) >> 1));
    for (int r = 0; r < k1; r++)
        for (int j = threadIdx.x; j < N; j += blockDim.x) acc[(size_t)r * N + j] = gen_rot(lut + (size_t)r * N, j, 2 * N - bhat, N);
    __syncthreads();

    // forward transforms run in groups of `grp` (r, level) pairs, inverse transforms in groups of `grp_inv` polynomials
    const int jobs = k1 * a.level, slots = a.fft_in_smem ? max(1, a.polys_in_smem) : 1;
    const int grp = (jobs + (jobs + slots - 1) / slots - 1) / ((jobs + slots - 1) / slots);
    const int grp_inv = (k1 + (k1 + slots - 1) / slots - 1) / ((k1 + slots - 1) / slots);
    const int log2half = a.log2N - 1;

    for (int i = 0; i < a.n; i++) {
        const Torus ai = lwe[i];
        if (ai == 0) continue;                                           // bootstrap.rs:281
        const int ahat = (int)((((ai >> ms_shift) + 1) >> 1));
        bool first = true;
        // job = r * level + (level - lvl): polynomial r outermost, levels from `level` down to 1 (ggsw.rs:524), the order
        // in which the reference adds the products into the Fourier accumulators
        for (int j0 = 0; j0 < jobs; j0 += grp) {
            const int cnt = min(grp, jobs - j0);
            for (int x = threadIdx.x; x < (cnt << log2half); x += blockDim.x) {
                const int job = j0 + (x >> log2half), j = x & (half - 1);
                const int r = job / a.level, lvl = a.level - job % a.level;
                const Torus *ar = acc + (size_t)r * N;
                const Torus v0 = gen_rot(ar, j, ahat, N) - ar[j];
                const Torus v1 = gen_rot(ar, j + half, ahat, N) - ar[j + half];
                const double d0 = gen_digit<Torus>(v0, a.base_log, a.level, lvl);
                const double d1 = gen_digit<Torus>(v1, a.base_log, a.level, lvl);
                buf[x] = gen_cmul(make_double2(d0, d1), a.twist[j]);
            }
            gen_fft_forward(buf, roots, half, cnt);
            for (int x = threadIdx.x; x < (k1 << log2half); x += blockDim.x) {
                const int c = x >> log2half, f = x & (half - 1);
                double2 sum = first ? make_double2(0.0, 0.0) : outf[x];
                for (int g = 0; g < cnt; g++) {
                    const int job = j0 + g, r = job / a.level, lvl = a.level - job % a.level;
                    const double2 *b = a.bsk + (((((size_t)i * a.level + (lvl - 1)) * k1 + r) * k1) + c) * half;
                    const double2 p = gen_cmul(b[f], buf[((size_t)g << log2half) + f]);
                    sum = (first && g == 0) ? p : make_double2(sum.x + p.x, sum.y + p.y);   // is_output_uninit, ggsw.rs:652-676
                }
                outf[x] = sum;
            }
            first = false;
            __syncthreads();
        }
        for (int c0 = 0; c0 < k1; c0 += grp_inv) {
            const int cnt = min(grp_inv, k1 - c0);
            const double2 *o = outf + ((size_t)c0 << log2half);
            for (int x = threadIdx.x; x < (cnt << log2half); x += blockDim.x) buf[x] = o[x];
            gen_fft_inverse(buf, roots, half, cnt);
            for (int x = threadIdx.x; x < (cnt << log2half); x += blockDim.x) {
                const int j = x & (half - 1);
                Torus *ac = acc + (size_t)(c0 + (x >> log2half)) * N;
                const double2 tw = a.twist[j];
                const double2 y = gen_cmul(buf[x], make_double2(tw.x, -tw.y));
                ac[j] += gen_from_torus<Torus>(y.x);
                ac[j + half] += gen_from_torus<Torus>(y.y);
            }
            __syncthreads();
        }
    }

    // extract_lwe_sample_from_glwe_ciphertext, nth = 0 (glwe_sample_extraction.rs:91-147)
    Torus *o = (Torus *)a.out + (size_t)ct * ((size_t)a.k * N + 1);
    for (int r = 0; r < a.k; r++)
        for (int j = threadIdx.x; j < N; j += blockDim.x)
            o[(size_t)r * N + j] = j == 0 ? acc[(size_t)r * N] : (Torus)(0 - acc[(size_t)r * N + N - j]);
    if (threadIdx.x == 0) o[(size_t)a.k * N] = acc[(size_t)a.k * N];
}

// standard-domain polynomial -> Fourier (torus scaled to [-1/2, 1/2), twist, forward transform), times 2/N so the
// inverse transform needs no normalisation (fft/mod.rs:197-218).  One CTA per polynomial.
template <typename Torus>
__global__ void __launch_bounds__(512, 1) bsk_to_fourier_generic_kernel(const Torus *__restrict__ bsk_std, double2 *__restrict__ bsk_f,
                                                                        const double2 *__restrict__ roots, const double2 *__restrict__ twist,
                                                                        double2 *__restrict__ scratch, const int log2N, const int fft_in_smem) {
    extern __shared__ __align__(16) unsigned char gen_smem[];
    constexpr int bits = TorusTraits<Torus>::bits;
    const int N = 1 << log2N, half = N >> 1;
    const size_t poly = blockIdx.x;
    const Torus *src = bsk_std + poly * N;
    double2 *buf = fft_in_smem ? reinterpret_cast<double2 *>(gen_smem) : scratch + poly * half;
    const double scale = bits == 64 ? 0x1p-64 : 0x1p-32;
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const double re = (double)(typename TorusTraits<Torus>::Signed)src[j] * scale;
        const double im = (double)(typename TorusTraits<Torus>::Signed)src[j + half] * scale;
        buf[j] = gen_cmul(make_double2(re, im), twist[j]);
    }
    gen_fft_forward(buf, roots, half);
    const double norm = 1.0 / (double)half;
    for (int f = threadIdx.x; f < half; f += blockDim.x) bsk_f[poly * half + f] = make_double2(buf[f].x * norm, buf[f].y * norm);
}

// Plain CUDA-core keyswitch for torus words the tensor-core path does not cover (u32 boolean keys):
// keyswitch_lwe_ciphertext, core_crypto/algorithms/lwe_keyswitch.rs:96-170.  One CTA per ciphertext, each thread owns
// output words; the decomposed digits of the input mask are staged in shared memory chunk by chunk.
template <typename Torus>
__global__ void __launch_bounds__(256) ks_generic_kernel(const Torus *__restrict__ in, const Torus *__restrict__ ksk, Torus *__restrict__ out,
                                                         const int batch, const int n_in, const int out_size, const int base_log, const int level) {
    __shared__ int s_digit[256 * 8];
    constexpr int bits = TorusTraits<Torus>::bits;
    const int ct = blockIdx.x;
    if (ct >= batch) return;
    const Torus *x = in + (size_t)ct * (n_in + 1);
    Torus accv[4] = {0, 0, 0, 0};                                        // out_size <= 1024 with 256 threads
    const Torus mask = ((Torus)1 << base_log) - 1;
    const int non_rep = bits - base_log * level;
    for (int i0 = 0; i0 < n_in; i0 += 256) {
        __syncthreads();
        if (i0 + (int)threadIdx.x < n_in) {
            Torus state = x[i0 + threadIdx.x] >> (non_rep - 1);
            state = (Torus)(state + 1) >> 1;
            for (int li = 0; li < level; li++) {                         // storage order of the KSK: level l first (lwe_keyswitch_key_generation.rs:109-111)
                Torus res = state & mask;
                state >>= base_log;
                const Torus carry = ((Torus)((res - 1) | state) & res) >> (base_log - 1);
                state += carry;
                s_digit[threadIdx.x * 8 + li] = (int)(typename TorusTraits<Torus>::Signed)(res - (carry << base_log));
            }
        }
        __syncthreads();
        const int cnt = min(256, n_in - i0);
        for (int ii = 0; ii < cnt; ii++)
            for (int li = 0; li < level; li++) {
                const Torus d = (Torus)(typename TorusTraits<Torus>::Signed)s_digit[ii * 8 + li];
                const Torus *row = ksk + ((size_t)(i0 + ii) * level + li) * out_size;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int j = threadIdx.x + 256 * q;
                    if (j < out_size) accv[q] += d * row[j];
                }
            }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int j = threadIdx.x + 256 * q;
        if (j < out_size) out[(size_t)ct * out_size + j] = (j == out_size - 1 ? x[n_in] : (Torus)0) - accv[q];
    }
}

}  // namespace b200
