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
//   * TMEM holds, per lane, the accumulator's home copy (128 columns per warp) and the lane's
//     column of the inter-pass twiddle table (128 columns per TMEM quadrant);
//   * a rotation copy of the accumulator lives in the warp's shared transposition buffer between
//     two CMUX steps (acc * X^a~ is a data-dependent gather, hence shared memory);
//   * BSK_SMEM: the 64 KiB Fourier BSK slice of step i is fetched ONCE per CTA by a bulk async
//     copy (TMA, cp.async.bulk + mbarrier) into shared memory and reused by all CTS ciphertexts;
//     the warp that consumes it last issues the copy for step i + 1.  Otherwise every warp reads
//     its rows straight from L2 with 512 B coalesced loads.
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
    long long *dbg;             // development: per-warp phase timestamps of CTA 0 (nullptr = off)
};

constexpr int kMaxSmallDim = 1024;       // capacity of the per-ciphertext a~ table (u16 entries)
constexpr int kPbsHeaderBytes = 128;     // tmem slot, mbarrier, consumer counter
constexpr int kBskSliceBytes = 4 * kHalf * (int)sizeof(double2);   // 65,536
constexpr uint32_t kTmemTwCols = 128;    // twiddle columns per quadrant

__host__ __device__ constexpr size_t pbs_ct_smem_bytes() {
    return (size_t)2 * kTBufElems * sizeof(double2) + kMaxSmallDim * sizeof(uint16_t);
}
template <int CTS, bool BSK_SMEM>
__host__ __device__ constexpr size_t pbs_smem_bytes() {
    return kPbsHeaderBytes + (BSK_SMEM ? kBskSliceBytes : 0) + CTS * pbs_ct_smem_bytes();
}
template <int CTS>
__host__ __device__ constexpr uint32_t pbs_tmem_cols() {
    // twiddles + (CTS*2 warps / 4 quadrants) x 128 accumulator columns; power of two >= 32
    return (kTmemTwCols + (CTS * 2 / 4) * 128) <= 256 ? 256 : 512;
}

// fast_pbs_modulus_switch, core_crypto/fft_impl/common.rs:26-43 (log2 N = 11): result in [0, 2N]
__device__ __forceinline__ uint32_t modswitch2048(uint64_t x) { return (uint32_t)(((x >> 51) + 1) >> 1); }

// closest_representable + the single balanced digit of level 1 (base 2^23):
// commons/math/decomposition/decomposer.rs:98-116, iter.rs:120-127.  Only bits 63..40 of d matter.
// Result as a double (exact: |digit| <= 2^22).  The int -> double conversion is issued to the XU
// pipe (I2F.F64.S32) on purpose: the FP64 pipe is the binding resource of this kernel.
__device__ __forceinline__ double digit23_as_double(uint64_t d) {
    const uint32_t hi = (uint32_t)(d >> 32);
    const uint32_t t = (((hi >> 8) + 1u) >> 1) & 0x7FFFFFu;
    const int32_t dig = (int32_t)t - ((t > 0x400000u) ? 0x800000 : 0);
    return __int2double_rn(dig);
}

// from_torus, core_crypto/commons/math/torus/mod.rs:72-78 (round-half-even like the x86 SIMD path)
__device__ __forceinline__ uint64_t from_torus_dev(double x) {
    const double f = x - rint(x);
    return (uint64_t)__double2ll_rn(f * 18446744073709551616.0);
}

__device__ __forceinline__ void ct_barrier(const int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// ---- mbarrier / bulk async copy (TMA 1-D) ----------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
                     "r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes),
                 "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
// one thread: fetch the Fourier BSK slice of CMUX step i into shared memory
__device__ __forceinline__ void issue_bsk_slice(double2 *bsk_s, const double2 *bsk_g, const int i, uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the buffer before async writes
    mbar_arrive_expect_tx(bar, (uint32_t)kBskSliceBytes);
    const double2 *src = bsk_g + (size_t)i * 4 * kHalf;
#pragma unroll
    for (int part = 0; part < 4; part++)
        bulk_g2s(bsk_s + part * kHalf, src + part * kHalf, kBskSliceBytes / 4, bar);
}

// TMEM-resident twiddles: lane-private column of T', 4 twiddles (16 words) per load
struct TmemTwiddles {
    uint32_t taddr;
    __device__ __forceinline__ void issue(const int chunk, uint32_t (&r)[16]) const { tmem_ld16(taddr + chunk * 16, r); }
    __device__ __forceinline__ void wait() const { tmem_wait_ld(); }
};

__device__ __forceinline__ uint64_t pack64(const uint32_t lo, const uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

template <int CTS, bool BSK_SMEM>
__global__ void __launch_bounds__(CTS * 64, 1) pbs_kernel(const PbsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctl = warp >> 1, p = warp & 1;
    const int ct = blockIdx.x * CTS + ctl;
    const bool active = ct < a.batch;

    uint32_t *slot = reinterpret_cast<uint32_t *>(smem);
    uint64_t *bsk_bar = reinterpret_cast<uint64_t *>(smem + 8);
    unsigned int *consumed = reinterpret_cast<unsigned int *>(smem + 16);
    double2 *bsk_s = reinterpret_cast<double2 *>(smem + kPbsHeaderBytes);
    unsigned char *ctbase = smem + kPbsHeaderBytes + (BSK_SMEM ? kBskSliceBytes : 0) + (size_t)ctl * pbs_ct_smem_bytes();
    double2 *tb_own = reinterpret_cast<double2 *>(ctbase) + p * kTBufElems;
    const double2 *tb_oth = reinterpret_cast<double2 *>(ctbase) + (1 - p) * kTBufElems;
    uint16_t *ahat = reinterpret_cast<uint16_t *>(ctbase + (size_t)2 * kTBufElems * sizeof(double2));
    uint64_t *rot = reinterpret_cast<uint64_t *>(tb_own);  // rotation copy aliases the transposition buffer

    // ---------------------------------------------------------------- CTA setup
    if (warp == 0) tmem_alloc(slot, pbs_tmem_cols<CTS>());
    if (BSK_SMEM && threadIdx.x == 0) {
        mbar_init(bsk_bar, 1);
        *consumed = 0;
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *slot;
    const uint32_t tquad = tbase + (((uint32_t)(warp & 3) * 32u) << 16);
    const uint32_t t_acc = tquad + kTmemTwCols + (uint32_t)(warp >> 2) * 128u;
    const TmemTwiddles tw{tquad};
    if (warp < 4) {   // one warp per TMEM quadrant stores its lanes' twiddle columns
#pragma unroll
        for (int c = 0; c < 8; c++) {
            uint32_t r[16];
            GlobalTwiddles{a.twid, lane}.issue(c, r);
            tmem_st16(tquad + c * 16, r);
        }
        tmem_wait_st();
    }
    const int n_act_cts = min(CTS, a.batch - (int)blockIdx.x * CTS);
    const unsigned int n_act_warps = 2u * (unsigned int)n_act_cts;
    if (BSK_SMEM && threadIdx.x == 0) issue_bsk_slice(bsk_s, a.bsk, 0, bsk_bar);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();

    if (active) {
        // ---------------------------------------------------------------- prologue
        const uint64_t *lwe = a.lwe_small + (size_t)ct * (a.n + 1);
        for (int i = p * 32 + lane; i < a.n; i += 64) ahat[i] = (uint16_t)modswitch2048(lwe[i]);
        const uint32_t bhat = modswitch2048(lwe[a.n]);
        const uint64_t *lut = a.luts + ((size_t)(a.lut_idx ? a.lut_idx[ct] : 0u) * 2 + p) * kN;
        // acc = LUT * X^-b~: polynomial_wrapping_monic_monomial_div (polynomial_algorithms.rs:315-354)
#pragma unroll
        for (int c = 0; c < 8; c++) {
            uint32_t h[16];
#pragma unroll
            for (int mm = 0; mm < 4; mm++) {
                const int j = lane + 32 * (c * 4 + mm);
                const uint32_t i0 = (uint32_t)(j + bhat) & 4095u, i1 = (i0 + 1024u) & 4095u;
                uint64_t v0 = lut[i0 & 2047u], v1 = lut[i1 & 2047u];
                if (i0 & 2048u) v0 = 0 - v0;
                if (i1 & 2048u) v1 = 0 - v1;
                rot[j] = v0; rot[j + kHalf] = v1;
                h[4 * mm] = (uint32_t)v0; h[4 * mm + 1] = (uint32_t)(v0 >> 32);
                h[4 * mm + 2] = (uint32_t)v1; h[4 * mm + 3] = (uint32_t)(v1 >> 32);
            }
            tmem_st16(t_acc + c * 16, h);
        }
        tmem_wait_st();
        ct_barrier(1 + ctl);  // a~ table visible to both warps; rot copy visible within the warp


        // ---------------------------------------------------------------- CMUX loop
        // Steps with a~ = 0 (mod 2N) are not skipped as the reference does (bootstrap.rs:281): the
        // rotation is then the identity, every digit is 0 and the step adds exactly zero.
        for (int i = 0; i < a.n; i++) {
            const uint32_t ah = ahat[i];
            double xr[32], xi[32];
            // phase A: ct1 = acc * X^a~ - acc (polynomial_algorithms.rs:425-491), round + digit
            // (ggsw.rs:514-521), exact int -> double, twist by C_m (fft/mod.rs:220-239)
            {
                const uint32_t idx0 = (uint32_t)(lane + 4096 - (int)ah);
                uint32_t h0[16], h1[16];
                tmem_ld16(t_acc, h0);
#pragma unroll
                for (int c2 = 0; c2 < 4; c2++) {
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int c = 2 * c2 + half;
                        tmem_wait_ld();
                        uint32_t(&h)[16] = half ? h1 : h0;
                        if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, half ? h0 : h1);
#pragma unroll
                        for (int mm = 0; mm < 4; mm++) {
                            const int m = c * 4 + mm;
                            const uint32_t i0 = (idx0 + 32u * m) & 4095u, i1 = (i0 + 1024u) & 4095u;
                            uint64_t v0 = rot[i0 & 2047u], v1 = rot[i1 & 2047u];
                            if (i0 & 2048u) v0 = 0 - v0;
                            if (i1 & 2048u) v1 = 0 - v1;
                            double fr = digit23_as_double(v0 - pack64(h[4 * mm], h[4 * mm + 1]));
                            double fi = digit23_as_double(v1 - pack64(h[4 * mm + 2], h[4 * mm + 3]));
                            twist_m(fr, fi, m);
                            xr[brev5(m)] = fr; xi[brev5(m)] = fi;
                        }
                    }
                }
            }
            __syncwarp();  // all rotation reads done before the buffer is reused for the transposition

            fwd1024(xr, xi, tb_own, tw, lane);

            // Fourier-domain multiply (update_with_fmadd, ggsw.rs:616-697).  This warp holds F_p;
            // it keeps BSK[p][p]*F_p (written into the bit-reversed slot the inverse transform
            // wants) and hands BSK[p][1-p]*F_p to the sibling warp through shared memory.
            double zr[32], zi[32];
            {
                const double2 *bk;
                if (BSK_SMEM) {
                    mbar_wait(bsk_bar, (uint32_t)(i & 1));
                    bk = bsk_s + (size_t)p * 2 * kHalf + lane;
                } else {
                    bk = a.bsk + ((size_t)i * 4 + (size_t)p * 2) * kHalf + lane;
                }
                const double2 *b_own = bk + (size_t)p * kHalf;
                const double2 *b_oth = bk + (size_t)(1 - p) * kHalf;
#pragma unroll
                for (int q = 0; q < 32; q++) {
                    if (q > brev5(q)) continue;           // handled together with its mirror
#pragma unroll
                    for (int s = 0; s < 2; s++) {
                        const int qq = s ? brev5(q) : q;
                        if (s && qq == q) continue;
                        const double2 bo = BSK_SMEM ? b_own[qq * 32] : __ldg(b_own + qq * 32);
                        const double2 bx = BSK_SMEM ? b_oth[qq * 32] : __ldg(b_oth + qq * 32);
                        const double fr = xr[qq], fi = xi[qq];
                        zr[brev5(qq)] = fma(-bo.y, fi, bo.x * fr);
                        zi[brev5(qq)] = fma(bo.y, fr, bo.x * fi);
                        double2 o;
                        o.x = fma(-bx.y, fi, bx.x * fr);
                        o.y = fma(bx.y, fr, bx.x * fi);
                        tb_own[qq * 32 + lane] = o;
                    }
                }
            }
            if (BSK_SMEM) {
                // this warp is done with the slice; the last of the CTA's warps refills the buffer
                __syncwarp();
                if (lane == 0) {
                    const unsigned int old = atomicAdd(consumed, 1u);
                    if (old == (unsigned int)(i + 1) * n_act_warps - 1u && i + 1 < a.n)
                        issue_bsk_slice(bsk_s, a.bsk, i + 1, bsk_bar);
                }
            }
            ct_barrier(1 + ctl);
#pragma unroll
            for (int q = 0; q < 32; q++) {
                const double2 o = tb_oth[q * 32 + lane];
                zr[brev5(q)] += o.x; zi[brev5(q)] += o.y;
            }
            ct_barrier(1 + ctl);

            inv1024(zr, zi, tb_own, tw, lane);

            // phase D: untwist, from_torus, wrapping add (fft/mod.rs:285-304), refresh both copies
            {
                uint32_t h0[16], h1[16];
                tmem_ld16(t_acc, h0);
#pragma unroll
                for (int c2 = 0; c2 < 4; c2++) {
#pragma unroll
                    for (int half = 0; half < 2; half++) {
                        const int c = 2 * c2 + half;
                        tmem_wait_ld();
                        uint32_t(&h)[16] = half ? h1 : h0;
                        if (c < 7) tmem_ld16(t_acc + (c + 1) * 16, half ? h0 : h1);
#pragma unroll
                        for (int mm = 0; mm < 4; mm++) {
                            const int m = c * 4 + mm;
                            const int j = lane + 32 * m;
                            double yr = zr[m], yi = zi[m];
                            untwist_m(yr, yi, m);
                            const uint64_t a0 = pack64(h[4 * mm], h[4 * mm + 1]) + from_torus_dev(yr);
                            const uint64_t a1 = pack64(h[4 * mm + 2], h[4 * mm + 3]) + from_torus_dev(yi);
                            rot[j] = a0; rot[j + kHalf] = a1;
                            h[4 * mm] = (uint32_t)a0; h[4 * mm + 1] = (uint32_t)(a0 >> 32);
                            h[4 * mm + 2] = (uint32_t)a1; h[4 * mm + 3] = (uint32_t)(a1 >> 32);
                        }
                        tmem_st16(t_acc + c * 16, h);
                    }
                }
                tmem_wait_st();
            }
            __syncwarp();  // rotation copy complete before the next step's gather
        }

        // ---------------------------------------------------------------- sample extraction
        uint64_t *o = a.out + (size_t)ct * (kN + 1);
        if (p == 0) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t h[16];
                tmem_ld16(t_acc + c * 16, h);
                tmem_wait_ld();
#pragma unroll
                for (int mm = 0; mm < 4; mm++) {
                    const int j = lane + 32 * (c * 4 + mm);
                    const uint64_t a0 = pack64(h[4 * mm], h[4 * mm + 1]);
                    const uint64_t a1 = pack64(h[4 * mm + 2], h[4 * mm + 3]);
                    if (j == 0) o[0] = a0; else o[kN - j] = 0 - a0;
                    o[kHalf - j] = 0 - a1;  // coefficient j + 1024 -> index N - (j + 1024)
                }
            }
        } else {
            uint32_t h[16];
            tmem_ld16(t_acc, h);
            tmem_wait_ld();
            if (lane == 0) o[kN] = pack64(h[0], h[1]);
        }
    }

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, pbs_tmem_cols<CTS>());
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
    const GlobalTwiddles tw{twid, lane};
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
        fwd1024(xr, xi, tbuf[warp], tw, lane);
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
    __shared__ __align__(16) double2 fa[kHalf];
    const int lane = threadIdx.x;
    const int idx = blockIdx.x;
    if (idx >= count) return;
    const GlobalTwiddles tw{twid, lane};
    const uint64_t *pa = a_int + (size_t)idx * kN, *pb = b_torus + (size_t)idx * kN;
    double xr[32], xi[32];
#pragma unroll
    for (int m = 0; m < 32; m++) {
        const int j = lane + 32 * m;
        double fr = (double)(long long)pa[j], fi = (double)(long long)pa[j + kHalf];
        twist_m(fr, fi, m);
        xr[brev5(m)] = fr; xi[brev5(m)] = fi;
    }
    fwd1024(xr, xi, tbuf, tw, lane);
#pragma unroll
    for (int q = 0; q < 32; q++) fa[q * 32 + lane] = make_double2(xr[q], xi[q]);
#pragma unroll
    for (int m = 0; m < 32; m++) {
        const int j = lane + 32 * m;
        double gr = (double)(long long)pb[j] * 0x1p-64, gi = (double)(long long)pb[j + kHalf] * 0x1p-64;
        twist_m(gr, gi, m);
        xr[brev5(m)] = gr; xi[brev5(m)] = gi;
    }
    fwd1024(xr, xi, tbuf, tw, lane);
    double zr[32], zi[32];
#pragma unroll
    for (int q = 0; q < 32; q++) {
        const double2 f = fa[q * 32 + lane];
        const double sr = xr[q] * 0x1p-10, si = xi[q] * 0x1p-10;
        zr[brev5(q)] = fma(-si, f.y, sr * f.x);
        zi[brev5(q)] = fma(si, f.x, sr * f.y);
    }
    inv1024(zr, zi, tbuf, tw, lane);
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
