// pbs_kernel.cuh -- batched programmable bootstrap (blind rotation + sample extraction) for
// PARAM_MESSAGE_2_CARRY_2_KS_PBS shapes: k = 1, N = 2048, one decomposition level.
//
// Replaces, per ciphertext, FourierLweBootstrapKeyView::bootstrap
// (core_crypto/fft_impl/fft64/crypto/bootstrap.rs:333-364): accumulator setup acc = LUT * X^-b~
// (:254-271), the n-step CMUX loop acc += BSK_i (x) (acc * X^a~_i - acc) (:279-316, external
// product ggsw.rs:477-598) and extract_lwe_sample_from_glwe_ciphertext
// (algorithms/glwe_sample_extraction.rs:91-147), all fused in one kernel.
//
// Mapping: one warp per GLWE polynomial, two warps per ciphertext, CTS ciphertexts per CTA.
//   * lane l owns coefficients l + 32 m and l + 32 m + 1024 (home layout, see fft.cuh);
//   * the accumulator's home copy lives in TMEM (or shared memory when USE_TMEM = false);
//   * a rotation copy lives in the warp's shared transposition buffer between two CMUX steps
//     (the rotation acc * X^a~ is a data-dependent gather, hence shared memory);
//   * the Fourier BSK slice of step i (64 KiB) is read straight from L2 with 512 B coalesced
//     loads, in the same [q][lane] order the forward transform leaves its output in.
#pragma once
#include "fft.cuh"
#include "tmem.cuh"

namespace b200 {

struct PbsArgs {
    const uint64_t *lwe_small;  // [batch][n + 1]
    const uint32_t *lut_idx;    // [batch] or nullptr (LUT 0)
    const uint64_t *luts;       // [n_luts][2][2048] GLWE accumulators (mask poly, body poly)
    const double2 *bsk;         // [n][row r][col c][q][lane], scaled by 1/1024
    const double2 *twid;        // T'[k1][l]
    uint64_t *out;              // [batch][2049]
    int batch;
    int n;
};

constexpr int kMaxSmallDim = 1024;  // capacity of the per-ciphertext a~ table (u16 entries)

template <bool USE_TMEM>
__host__ __device__ constexpr size_t pbs_ct_smem_bytes() {
    return (size_t)2 * kTBufElems * sizeof(double2) + kMaxSmallDim * sizeof(uint16_t) +
           (USE_TMEM ? 0 : (size_t)2 * kN * sizeof(uint64_t));
}
template <int CTS, bool USE_TMEM>
__host__ __device__ constexpr size_t pbs_smem_bytes() {
    return 16 + CTS * pbs_ct_smem_bytes<USE_TMEM>();
}
template <int CTS>
__host__ __device__ constexpr uint32_t pbs_tmem_cols() {
    // CTS*2 warps, 4 TMEM quadrants, 128 columns per warp; allocation must be a power of two >= 32
    return (CTS * 2 / 4) * 128 <= 128 ? 128 : ((CTS * 2 / 4) * 128 <= 256 ? 256 : 512);
}

// fast_pbs_modulus_switch, core_crypto/fft_impl/common.rs:26-43 (log2 N = 11): result in [0, 2N]
__device__ __forceinline__ uint32_t modswitch2048(uint64_t x) { return (uint32_t)(((x >> 51) + 1) >> 1); }

// closest_representable + the single balanced digit of level 1 (base 2^23):
// commons/math/decomposition/decomposer.rs:98-116, iter.rs:120-127.  Only bits 63..40 of d matter.
// Result as a double, converted exactly with the 2^52 + 2^31 biased-mantissa trick.
__device__ __forceinline__ double digit23_as_double(uint64_t d) {
    const uint32_t hi = (uint32_t)(d >> 32);
    const uint32_t t = (((hi >> 8) + 1u) >> 1) & 0x7FFFFFu;
    const int32_t dig = (int32_t)t - ((t > 0x400000u) ? 0x800000 : 0);
    return __hiloint2double(0x43300000, (int)((uint32_t)dig ^ 0x80000000u)) - 4503601774854144.0;
}

// from_torus, core_crypto/commons/math/torus/mod.rs:72-78 (round-half-even like the x86 SIMD path)
__device__ __forceinline__ uint64_t from_torus_dev(double x) {
    const double f = x - rint(x);
    return (uint64_t)__double2ll_rn(f * 18446744073709551616.0);
}

template <bool USE_TMEM>
__device__ __forceinline__ void home_load(uint32_t (&h)[32], const int c, const uint32_t taddr,
                                          const uint64_t *home_s, const int lane) {
    if (USE_TMEM) {
        tmem_ld32(taddr + c * 32, h);
        tmem_wait_ld();
    } else {
#pragma unroll
        for (int mm = 0; mm < 8; mm++) {
            const int j = lane + 32 * (c * 8 + mm);
            const uint64_t v0 = home_s[j], v1 = home_s[j + kHalf];
            h[4 * mm] = (uint32_t)v0; h[4 * mm + 1] = (uint32_t)(v0 >> 32);
            h[4 * mm + 2] = (uint32_t)v1; h[4 * mm + 3] = (uint32_t)(v1 >> 32);
        }
    }
}
template <bool USE_TMEM>
__device__ __forceinline__ void home_store(const uint32_t (&h)[32], const int c, const uint32_t taddr,
                                           uint64_t *home_s, const int lane) {
    if (USE_TMEM) {
        tmem_st32(taddr + c * 32, h);
    } else {
#pragma unroll
        for (int mm = 0; mm < 8; mm++) {
            const int j = lane + 32 * (c * 8 + mm);
            home_s[j] = ((uint64_t)h[4 * mm + 1] << 32) | h[4 * mm];
            home_s[j + kHalf] = ((uint64_t)h[4 * mm + 3] << 32) | h[4 * mm + 2];
        }
    }
}

__device__ __forceinline__ void ct_barrier(const int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

template <int CTS, bool USE_TMEM>
__global__ void __launch_bounds__(CTS * 64, 1) pbs_kernel(const PbsArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctl = warp >> 1, p = warp & 1;
    const int ct = blockIdx.x * CTS + ctl;
    const bool active = ct < a.batch;

    unsigned char *ctbase = smem + 16 + (size_t)ctl * pbs_ct_smem_bytes<USE_TMEM>();
    double2 *tb_own = reinterpret_cast<double2 *>(ctbase) + p * kTBufElems;
    const double2 *tb_oth = reinterpret_cast<double2 *>(ctbase) + (1 - p) * kTBufElems;
    uint16_t *ahat = reinterpret_cast<uint16_t *>(ctbase + (size_t)2 * kTBufElems * sizeof(double2));
    uint64_t *home_s = reinterpret_cast<uint64_t *>(ctbase + (size_t)2 * kTBufElems * sizeof(double2) +
                                                    kMaxSmallDim * sizeof(uint16_t)) + p * kN;
    uint64_t *rot = reinterpret_cast<uint64_t *>(tb_own);  // rotation copy aliases the transposition buffer

    uint32_t tbase = 0;
    if (USE_TMEM) {
        uint32_t *slot = reinterpret_cast<uint32_t *>(smem);
        if (warp == 0) tmem_alloc(slot, pbs_tmem_cols<CTS>());
        tmem_fence_before();
        __syncthreads();
        tmem_fence_after();
        tbase = *slot;
    }
    const uint32_t taddr = tbase + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 128u;

    if (active) {
        // ---------------------------------------------------------------- prologue
        const uint64_t *lwe = a.lwe_small + (size_t)ct * (a.n + 1);
        for (int i = p * 32 + lane; i < a.n; i += 64) ahat[i] = (uint16_t)modswitch2048(lwe[i]);
        const uint32_t bhat = modswitch2048(lwe[a.n]);
        const uint64_t *lut = a.luts + ((size_t)(a.lut_idx ? a.lut_idx[ct] : 0u) * 2 + p) * kN;
        // polynomial_wrapping_monic_monomial_div (algorithms/polynomial_algorithms.rs:315-354)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t h[32];
#pragma unroll
            for (int mm = 0; mm < 8; mm++) {
                const int j = lane + 32 * (c * 8 + mm);
                const uint32_t i0 = (uint32_t)(j + bhat) & 4095u, i1 = (i0 + 1024u) & 4095u;
                uint64_t v0 = lut[i0 & 2047u], v1 = lut[i1 & 2047u];
                if (i0 & 2048u) v0 = 0 - v0;
                if (i1 & 2048u) v1 = 0 - v1;
                rot[j] = v0; rot[j + kHalf] = v1;
                h[4 * mm] = (uint32_t)v0; h[4 * mm + 1] = (uint32_t)(v0 >> 32);
                h[4 * mm + 2] = (uint32_t)v1; h[4 * mm + 3] = (uint32_t)(v1 >> 32);
            }
            home_store<USE_TMEM>(h, c, taddr, home_s, lane);
        }
        if (USE_TMEM) tmem_wait_st();
        ct_barrier(1 + ctl);  // a~ table visible to both warps; rot copy visible within the warp

        // ---------------------------------------------------------------- CMUX loop
        for (int i = 0; i < a.n; i++) {
            const uint32_t ah = ahat[i];
            // X^0 / X^{2N}: ct1 = acc*X^a - acc = 0 and the external product adds exactly zero
            // (the reference skips on the un-switched element, bootstrap.rs:281; same result).
            if ((ah & 4095u) == 0u) continue;

            double xr[32], xi[32];
            // phase A: ct1 = acc * X^a~ - acc (polynomial_algorithms.rs:425-491), round + digit
            // (ggsw.rs:514-521), exact int -> double, twist by C_m (fft/mod.rs:220-239)
            const uint32_t idx0 = (uint32_t)(lane + 4096 - (int)ah);
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t h[32];
                home_load<USE_TMEM>(h, c, taddr, home_s, lane);
#pragma unroll
                for (int mm = 0; mm < 8; mm++) {
                    const int m = c * 8 + mm;
                    const uint32_t i0 = (idx0 + 32u * m) & 4095u, i1 = (i0 + 1024u) & 4095u;
                    uint64_t v0 = rot[i0 & 2047u], v1 = rot[i1 & 2047u];
                    if (i0 & 2048u) v0 = 0 - v0;
                    if (i1 & 2048u) v1 = 0 - v1;
                    const uint64_t a0 = ((uint64_t)h[4 * mm + 1] << 32) | h[4 * mm];
                    const uint64_t a1 = ((uint64_t)h[4 * mm + 3] << 32) | h[4 * mm + 2];
                    double fr = digit23_as_double(v0 - a0), fi = digit23_as_double(v1 - a1);
                    twist_m(fr, fi, m);
                    xr[brev5(m)] = fr; xi[brev5(m)] = fi;
                }
            }
            __syncwarp();  // all rotation reads done before the buffer is reused for the transposition

            fwd1024(xr, xi, tb_own, a.twid, lane);

            // Fourier-domain multiply (update_with_fmadd, ggsw.rs:616-697).  This warp holds
            // F_p; it keeps BSK[p][p]*F_p and hands BSK[p][1-p]*F_p to the sibling warp.
            double zr[32], zi[32];
            {
                const double2 *bk = a.bsk + ((size_t)i * 4 + (size_t)p * 2) * kHalf + lane;
                const double2 *b_own = bk + (size_t)p * kHalf;
                const double2 *b_oth = bk + (size_t)(1 - p) * kHalf;
#pragma unroll
                for (int q = 0; q < 32; q++) {
                    const double2 bo = __ldg(b_own + q * 32);
                    const double2 bx = __ldg(b_oth + q * 32);
                    const double fr = xr[q], fi = xi[q];
                    zr[brev5(q)] = fma(-bo.y, fi, bo.x * fr);
                    zi[brev5(q)] = fma(bo.y, fr, bo.x * fi);
                    double2 o;
                    o.x = fma(-bx.y, fi, bx.x * fr);
                    o.y = fma(bx.y, fr, bx.x * fi);
                    tb_own[q * 32 + lane] = o;
                }
            }
            ct_barrier(1 + ctl);
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const double2 o = tb_oth[q * 32 + lane];
                zr[brev5(q)] += o.x; zi[brev5(q)] += o.y;
            }
            ct_barrier(1 + ctl);

            inv1024(zr, zi, tb_own, a.twid, lane);

            // phase D: untwist, from_torus, wrapping add (fft/mod.rs:285-304), refresh both copies
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t h[32];
                home_load<USE_TMEM>(h, c, taddr, home_s, lane);
#pragma unroll
                for (int mm = 0; mm < 8; mm++) {
                    const int m = c * 8 + mm;
                    const int j = lane + 32 * m;
                    double yr = zr[m], yi = zi[m];
                    untwist_m(yr, yi, m);
                    const uint64_t a0 = (((uint64_t)h[4 * mm + 1] << 32) | h[4 * mm]) + from_torus_dev(yr);
                    const uint64_t a1 = (((uint64_t)h[4 * mm + 3] << 32) | h[4 * mm + 2]) + from_torus_dev(yi);
                    rot[j] = a0; rot[j + kHalf] = a1;
                    h[4 * mm] = (uint32_t)a0; h[4 * mm + 1] = (uint32_t)(a0 >> 32);
                    h[4 * mm + 2] = (uint32_t)a1; h[4 * mm + 3] = (uint32_t)(a1 >> 32);
                }
                home_store<USE_TMEM>(h, c, taddr, home_s, lane);
            }
            if (USE_TMEM) tmem_wait_st();
            __syncwarp();  // rotation copy complete before the next step's gather
        }

        // ---------------------------------------------------------------- sample extraction
        uint64_t *o = a.out + (size_t)ct * (kN + 1);
        if (p == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t h[32];
                home_load<USE_TMEM>(h, c, taddr, home_s, lane);
#pragma unroll
                for (int mm = 0; mm < 8; mm++) {
                    const int j = lane + 32 * (c * 8 + mm);
                    const uint64_t a0 = ((uint64_t)h[4 * mm + 1] << 32) | h[4 * mm];
                    const uint64_t a1 = ((uint64_t)h[4 * mm + 3] << 32) | h[4 * mm + 2];
                    if (j == 0) o[0] = a0; else o[kN - j] = 0 - a0;
                    o[kHalf - j] = 0 - a1;  // coefficient j + 1024 -> index N - (j + 1024)
                }
            }
        } else {
            uint32_t h[32];
            home_load<USE_TMEM>(h, 0, taddr, home_s, lane);
            if (lane == 0) o[kN] = ((uint64_t)h[1] << 32) | h[0];
        }
    }

    if (USE_TMEM) {
        tmem_fence_before();
        __syncthreads();
        if (warp == 0) tmem_dealloc(tbase, pbs_tmem_cols<CTS>());
    }
}

// ---------------------------------------------------------------------------------------------
// Standard-domain BSK -> Fourier BSK in this library's [q][lane] order, scaled by 1/1024.
// Replaces par_convert_standard_lwe_bootstrap_key_to_fourier
// (core_crypto/algorithms/lwe_bootstrap_key_conversion.rs:99+, fft/mod.rs:197-218,719-764).
// One warp per polynomial; grid-stride over n_polys.
__global__ void __launch_bounds__(64) bsk_to_fourier_kernel(const uint64_t *__restrict__ bsk_std,
                                                            double2 *__restrict__ bsk_f,
                                                            const double2 *__restrict__ twid, const int n_polys) {
    __shared__ __align__(16) double2 tbuf[2][kTBufElems];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int poly = blockIdx.x * 2 + warp; poly < n_polys; poly += gridDim.x * 2) {
        const uint64_t *src = bsk_std + (size_t)poly * kN;
        double xr[32], xi[32];
#pragma unroll
        for (int m = 0; m < 32; m++) {
            const int j = lane + 32 * m;
            // convert_forward_torus: i64 -> f64, times 2^-64
            double fr = (double)(long long)src[j] * 0x1p-64;
            double fi = (double)(long long)src[j + kHalf] * 0x1p-64;
            twist_m(fr, fi, m);
            xr[brev5(m)] = fr; xi[brev5(m)] = fi;
        }
        fwd1024(xr, xi, tbuf[warp], twid, lane);
        double2 *dst = bsk_f + (size_t)poly * kHalf + lane;
#pragma unroll
        for (int q = 0; q < 32; q++) dst[q * 32] = make_double2(xr[q] * 0x1p-10, xi[q] * 0x1p-10);
    }
}

// Debug/unit-test kernel: out += a (x) b over Z[X]/(X^N + 1) with a taken as integer digits
// (|a| < 2^31) and b as torus elements, through exactly the transforms the PBS uses.
// Mirrors the reference's FFT product test (fft/tests.rs:82-222).  One warp per product.
__global__ void __launch_bounds__(32) negacyclic_mul_test_kernel(const uint64_t *__restrict__ a_int,
                                                                 const uint64_t *__restrict__ b_torus,
                                                                 uint64_t *__restrict__ out,
                                                                 const double2 *__restrict__ twid, const int count) {
    __shared__ __align__(16) double2 tbuf[kTBufElems];
    const int lane = threadIdx.x;
    const int idx = blockIdx.x;
    if (idx >= count) return;
    const uint64_t *pa = a_int + (size_t)idx * kN, *pb = b_torus + (size_t)idx * kN;
    double ar[32], ai[32], br[32], bi[32];
#pragma unroll
    for (int m = 0; m < 32; m++) {
        const int j = lane + 32 * m;
        double fr = (double)(long long)pa[j], fi = (double)(long long)pa[j + kHalf];
        twist_m(fr, fi, m);
        ar[brev5(m)] = fr; ai[brev5(m)] = fi;
        double gr = (double)(long long)pb[j] * 0x1p-64, gi = (double)(long long)pb[j + kHalf] * 0x1p-64;
        twist_m(gr, gi, m);
        br[brev5(m)] = gr; bi[brev5(m)] = gi;
    }
    fwd1024(ar, ai, tbuf, twid, lane);
    fwd1024(br, bi, tbuf, twid, lane);
    double zr[32], zi[32];
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const double sr = br[q] * 0x1p-10, si = bi[q] * 0x1p-10;
        zr[brev5(q)] = fma(-si, ai[q], sr * ar[q]);
        zi[brev5(q)] = fma(si, ar[q], sr * ai[q]);
    }
    inv1024(zr, zi, tbuf, twid, lane);
    uint64_t *po = out + (size_t)idx * kN;
#pragma unroll
    for (int m = 0; m < 32; m++) {
        const int j = lane + 32 * m;
        double yr = zr[m], yi = zi[m];
        untwist_m(yr, yi, m);
        po[j] += from_torus_dev(yr);
        po[j + kHalf] += from_torus_dev(yi);
    }
}

}  // namespace b200
