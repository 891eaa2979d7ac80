// pbs_kernel5.cuh -- the batched programmable bootstrap for k = 1, N = 2048, one level of base 2^23
// (PARAM_MESSAGE_2_CARRY_2 and the other N = 2048 classic sets): one warp per GLWE polynomial, kCts ciphertexts per CTA,
// one CTA per SM.  Replaces FourierLweBootstrapKeyView::bootstrap (core_crypto/fft_impl/fft64/crypto/bootstrap.rs:333-364):
// mod-switch (fft_impl/common.rs:26-43), accumulator setup LUT * X^-b~ (polynomial_algorithms.rs:315-354), lwe_dimension
// CMUX steps acc += GGSW_i (x) (acc * X^a~_i - acc) (bootstrap.rs:267-331, ggsw.rs:477-598: decompose, forward negacyclic
// FFT, multiply-accumulate with the Fourier BSK, inverse FFT, from_torus, wrapping add), sample extraction
// (lwe_sample_extraction, glwe_sample_extraction.rs).  Data layouts: DESIGN.md section 3.
//
// Against its predecessor pbs_kernel3 (tools/lab/pbs_kernel3.cuh, kept as the bit-for-bit regression reference; measured
// there: every phase of a CMUX step is bound by ONE unit that both warps of an SM sub-partition want at the same time --
// gather: instruction issue, transforms: FP64 pipe, exchange: shared memory + pair barriers, from_torus: the conversion unit):
//
//   * Accumulator convention G = -acc, the same 64-bit word in TMEM (home layout) and in the shared-memory rotation
//     copy; the rounding/bias constant C = 2^63 - 2^40 of the exact digit (decomposer.rs:98-116, iter.rs:120-127) is an
//     immediate of the high-word add:  digit + (2^22 - 1) = hi32(G_own +- G_src + C) >> 9.
//   * Rotated gather with NO per-element wrap or sign test.  The rotation copy carries an overflow zone of 8 slots
//     (256 words holding the negated first 256 coefficients).  A lane's 64 slots are visited in 8 groups of 8; for a
//     group the source indices of all 32 lanes lie in a window of 256 + 31 consecutive coefficients of the
//     4096-periodic negacyclic extension, so ONE warp-uniform (segment, offset) pair serves the whole group: the
//     address is base + immediate and the sign is one mask per group (polynomial_algorithms.rs:425-491).
//   * from_torus without the conversion unit (FRND / F2I are 13-cycle instructions here): two exponent-aligned
//     additions split x into its top 14 fractional bits and an exact remainder (torus/mod.rs:72-78; round-half-even
//     like fft/x86.rs:864; bit-identical to rint + cvt.rni except that a fractional part of exactly +1/2 gives 2^63
//     instead of the saturated 2^63 - 1).
//   * The two warps of a ciphertext write their transforms INTO EACH OTHER'S buffer, so no warp ever waits for the
//     sibling to finish reading: the two pair barriers of a step are split arrive / wait mbarriers whose arrive sits
//     a full 32-point transform before the matching wait.
#pragma once
#include "pbs_common.cuh"

namespace b200 {

constexpr int kZone5 = 256;                                   // overflow zone of the rotation copy (words)
constexpr int kBuf5Bytes = (kN + kZone5) * 8;                 // 18,432 B per warp: rotation copy / transposition / sibling's transform
constexpr int kHdr5Bytes = 128;                               // tmem slot @0, BSK mbarrier @8, consumer counter @16, pair mbarriers @32
static_assert(kBuf5Bytes >= kTBufElems * (int)sizeof(double2), "transposition buffer must fit");

template <int kCts>
__host__ __device__ constexpr size_t pbs5_smem_bytes() {
    return kHdr5Bytes + kBskSliceBytes + (size_t)kCts * (2 * kBuf5Bytes + kMaxSmallDim * sizeof(uint16_t));
}

// from_torus in three FP64 instructions instead of four (bit-identical; derivation at from_torus_fp in pbs_common.cuh)
#ifndef PBS5_FT3
#define PBS5_FT3 1
#endif
// register part C_m of the forward twist folded into the first pass (fft32_dit_twisted, fft.cuh)
#ifndef PBS5_TWFOLD
#define PBS5_TWFOLD 0
#endif

// The BSK slice is refilled in two halves, each with its own mbarrier and consumer counter: the diagonal polynomials B[p][p] as
// soon as every warp of the CTA has done its own product, the off-diagonal ones after the sibling products (kSplit = 1) or
// already before the inverse first pass (kSplit = 2).  The four ciphertexts of a CTA drift up to 0.7 of a step apart and are
// coupled only through this buffer: a refill that waits for the slowest warp's LAST use of the whole slice lands late for the
// fastest one.  Measured per 4096 bootstraps: one refill 46.4 ms, two halves 44.7 ms (4 per SM); 3 per SM: 6.34 / 6.25 / 6.09 ms
// per 444 for kSplit = 0 / 1 / 2.  -1 = that choice per kCts; 0, 1, 2 force one (tools/lab).
#ifndef PBS5_SPLIT_BSK
#define PBS5_SPLIT_BSK -1
#endif
template <int kCts> __host__ __device__ constexpr int pbs5_split() { return PBS5_SPLIT_BSK >= 0 ? PBS5_SPLIT_BSK : (kCts == 3 ? 2 : 1); }
// one thread: fetch half `which` (0: polynomials (0,0) and (1,1), 1: (0,1) and (1,0)) of the Fourier BSK slice of CMUX step i
__device__ __forceinline__ void issue_bsk_half(double2 *bsk_s, const double2 *bsk_g, const int i, const int which, uint64_t *bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive_expect_tx(bar, (uint32_t)kBskSliceBytes / 2);
    const double2 *src = bsk_g + (size_t)i * 4 * kHalf;
    const int p0 = which ? 1 : 0, p1 = which ? 2 : 3;
    bulk_g2s(bsk_s + p0 * kHalf, src + p0 * kHalf, kBskSliceBytes / 4, bar);   // (two copies of 8 KB per polynomial: -0.6 %, four: +1.2 %)
    bulk_g2s(bsk_s + p1 * kHalf, src + p1 * kHalf, kBskSliceBytes / 4, bar);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}

// G -= from_torus(x) on the two 32-bit halves of G, in three integer instructions (subtract with borrow; the bias of u
// and the 14 top bits from t go into the high word with one multiply-add), and the rotation copy written straight from
// the register pair the TMEM store uses.
__device__ __forceinline__ void acc_sub_from_torus(uint32_t &glo, uint32_t &ghi, const double x) {
    // the magic constant is lowered by 0x10CE ulps (0x10CE << 18 = 0x43380000, the high word of 1.5 * 2^52), so that the
    // multiply-add below removes the bias of u together with placing t's 14 bits: -(t.lo - 0x10CE) << 18
#if PBS5_FT3
    // three-instruction form (from_torus_fp, pbs_common.cuh) with the same lowered constant, 64 binades up
    const double t = fma(x, 18446744073709551616.0, __longlong_as_double(0x4657FFFFFFFFEF32ll));   // (1.5 * 2^52 - 0x10CE) * 2^50
    const double u = fma(x, 18446744073709551616.0, __longlong_as_double(0x4657FFFFFFFFEF38ll) - t);
#else
    const double m1 = __longlong_as_double(0x4257FFFFFFFFEF32ll);   // 1.5 * 2^38 - 0x10CE * 2^-14
    const double t = x + m1;
    const double l = x - (t - m1);
    const double u = fma(l, 18446744073709551616.0, 6755399441055744.0);
#endif
    uint32_t hi;
    asm("sub.cc.u32 %0, %0, %2;\n\tsubc.u32 %1, %3, %4;" : "+r"(glo), "=r"(hi) : "r"((uint32_t)__double2loint(u)), "r"(ghi), "r"((uint32_t)__double2hiint(u)));
    ghi = (uint32_t)__double2loint(t) * 0xFFFC0000u + hi;
}
__device__ __forceinline__ void sts_v2(uint64_t *p, const uint32_t lo, const uint32_t hi) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(lo), "r"(hi) : "memory");
}

// Development switches (tools/lab): bit 0 = last stage of the forward transform fused with the hand-over and the own
// product; bit 1 = sibling product at the leaves of a depth-first inverse first pass.
#ifndef PBS5_FUSE
#define PBS5_FUSE 0
#endif
// bit 2 = the accumulator's TMEM chunks hold the slots in the bit-reversed order in which the forward transform's first
// pass consumes its registers (chunk c, element k <-> slot brev5(4c + k)), so that butterflies become ready while the
// gather is still running.
#ifndef PBS5_ORDER
#define PBS5_ORDER 0
#endif
// 4 ciphertexts per CTA: the two warps of a ciphertext are warps w and w + 4, i.e. they share an SM sub-partition and a TMEM
// quadrant (measured: 1.2 % faster than warps 2c and 2c + 1 -- the pair then runs in lock step on one scheduler).  They
// could then hand the first kXtQ frequency blocks of their transforms over through the 128 tensor-memory columns the
// quadrant has left (64 per direction) instead of shared memory; measured slower (kXtQ = 8: +2 %, 16: +8 % time: the
// TMEM store / fence / load chain sits on the critical path of the hand-over), so kXtQ = 0.
#ifndef PBS5_XT
#define PBS5_XT 1
#endif
// digit -> double: 0 = exponent trick + DADD (FP64 pipe), 1 = I2F.F64.S32 (conversion unit), 2 = half and half
#ifndef PBS5_I2F
#define PBS5_I2F 0
#endif
#ifndef PBS5_XTQ
#define PBS5_XTQ 0
#endif
constexpr int kXtQ = PBS5_XTQ;             // frequency blocks q < kXtQ travel through TMEM
constexpr uint32_t kTmemXt0 = 384;         // column offset of the two 64-column exchange areas
__host__ __device__ constexpr int slot5(const int c, const int k) { return PBS5_ORDER ? brev5(4 * c + k) : 4 * c + k; }

// stages [kFirst, kLast] (butterfly distance 2^stage) of the in-register 32-point DIT transform of fft.cuh
template <bool INV, int kFirst, int kLast>
__device__ __forceinline__ void fft32_dit_stages(double (&xr)[32], double (&xi)[32]) {
#pragma unroll
    for (int half = 1 << kFirst; half <= (1 << kLast); half <<= 1) {
#pragma unroll
        for (int base = 0; base < 32; base += 2 * half) {
#pragma unroll
            for (int t = 0; t < half; t++)
                bfly<INV>(xr[base + t], xi[base + t], xr[base + t + half], xi[base + t + half], t * (16 / half));
        }
    }
}
// the same butterflies in depth-first order; leaf(r, xr, xi) produces register r (bit-reversed input order) just before
// its first use.  (The arrays are passed down explicitly: a lambda capturing them by reference sends them to local memory.)
template <bool INV, int kLo, int kLen, class Leaf>
__device__ __forceinline__ void fft32_dit_df(double (&xr)[32], double (&xi)[32], const Leaf &leaf) {
    if constexpr (kLen == 1) {
        leaf.template run<kLo>(xr, xi);
    } else {
        fft32_dit_df<INV, kLo, kLen / 2>(xr, xi, leaf);
        fft32_dit_df<INV, kLo + kLen / 2, kLen / 2>(xr, xi, leaf);
#pragma unroll
        for (int t = 0; t < kLen / 2; t++)
            bfly<INV>(xr[kLo + t], xi[kLo + t], xr[kLo + t + kLen / 2], xi[kLo + t + kLen / 2], t * (32 / kLen));
    }
}
// z[r] += B[1-p][p][q] * F_sibling[q] for the frequency block q = brev5(r) that register r of the inverse transform holds
template <bool kXT>
struct SiblingProduct {
    const double2 *b_oth, *f_oth;   // both already offset by the lane
    const uint32_t (*xt)[16];       // kXT: the sibling's blocks q < kXtQ as read from TMEM (4 blocks per 16 words)
    template <int kR>
    __device__ __forceinline__ void run(double (&zr)[32], double (&zi)[32]) const {
        constexpr int q = brev5(kR);
        const double2 bx = b_oth[q * 32];
        double2 g;
        if constexpr (kXT && q < kXtQ) g = make_double2(dbl(xt[q / 4][4 * (q % 4)], xt[q / 4][4 * (q % 4) + 1]), dbl(xt[q / 4][4 * (q % 4) + 2], xt[q / 4][4 * (q % 4) + 3]));
        else g = f_oth[q * 32];
        const double o_r = fma(bx.x, g.x, zr[kR]), o_i = fma(bx.x, g.y, zi[kR]);
        zr[kR] = fma(-bx.y, g.y, o_r); zi[kR] = fma(bx.y, g.x, o_i);
    }
};
template <int kR, bool kXT>
__device__ __forceinline__ void sibling_products(double (&zr)[32], double (&zi)[32], const SiblingProduct<kXT> &oth) {
    if constexpr (kR < 32) {
        oth.template run<kR>(zr, zi);
        sibling_products<kR + 1, kXT>(zr, zi, oth);
    }
}
template <int kQ, bool kXT = false>
__device__ __forceinline__ void own_product(double (&zr)[32], double (&zi)[32], const double (&xr)[32], const double (&xi)[32],
                                            const double2 *b_own, double2 *f_dst) {
    if constexpr (!(kXT && kQ < kXtQ)) f_dst[kQ * 32] = make_double2(xr[kQ], xi[kQ]);
    const double2 bo = b_own[kQ * 32];
    zr[brev5(kQ)] = fma(-bo.y, xi[kQ], bo.x * xr[kQ]);
    zi[brev5(kQ)] = fma(bo.y, xr[kQ], bo.x * xi[kQ]);
}
template <int kQ, bool kFused, bool kXT = false>
__device__ __forceinline__ void own_products(double (&zr)[32], double (&zi)[32], double (&xr)[32], double (&xi)[32],
                                             const double2 *b_own, double2 *f_dst) {
    if constexpr (kFused) {
        if constexpr (kQ < 16) {
            bfly<false>(xr[kQ], xi[kQ], xr[kQ + 16], xi[kQ + 16], kQ);   // last stage: outputs kQ and kQ + 16 are final
            own_product<kQ, kXT>(zr, zi, xr, xi, b_own, f_dst);
            own_product<kQ + 16, kXT>(zr, zi, xr, xi, b_own, f_dst);
            own_products<kQ + 1, kFused, kXT>(zr, zi, xr, xi, b_own, f_dst);
        }
    } else {
        if constexpr (kQ < 32) {
            own_product<kQ, kXT>(zr, zi, xr, xi, b_own, f_dst);
            own_products<kQ + 1, kFused, kXT>(zr, zi, xr, xi, b_own, f_dst);
        }
    }
}

#ifdef B200TFHE_TIMELINE
#define PBS5_TS(k) do { if (a.dbg && blockIdx.x == 0 && lane == 0 && i >= 100 && i < 108) a.dbg[((i - 100) * 8 + warp) * 16 + (k)] = clock64(); } while (0)
#else
#define PBS5_TS(k) do { } while (0)
#endif

#ifdef B200TFHE_LAB_DELAY
__device__ int g_lab_delay;   // development (tools/lab): the warps of ciphertext c enter the CMUX loop c * g_lab_delay cycles late
#endif

// kPhase = 1 (4 full ciphertexts per CTA only): the two halves of the CTA run HALF A STEP apart.  Warps 0-3 ("early") and
// warps 4-7 ("late") share their SM sub-partitions pairwise; a step is cut at the hand-over of the transforms into
// [gather, forward transform, own product] and [sibling product, inverse transform, from_torus], about 7k cycles each for
// a warp running alone, and the late half starts its first part when the early half starts its second (named barriers 5
// and 6, bar.arrive / bar.sync over 256 threads).  The two warps of a sub-partition then want different units most of
// the time (FP64 pipe vs issue slots / shared memory) instead of the same unit at the same time.
template <int kCts, int kPhase = 0>
__global__ void __launch_bounds__(kCts * 64, 1) pbs_kernel5(const PbsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr bool kXT = PBS5_XT && kCts == 4 && kPhase == 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctl = kXT ? (warp & 3) : (warp >> 1), p = kXT ? (warp >> 2) : (warp & 1);
    const int ct = blockIdx.x * kCts + ctl;
    const bool active = ct < a.batch;

    uint32_t *slot = reinterpret_cast<uint32_t *>(smem);
    uint64_t *bsk_bar = reinterpret_cast<uint64_t *>(smem + 8);
    unsigned int *consumed = reinterpret_cast<unsigned int *>(smem + 16);
    uint64_t *bsk_bar_o = reinterpret_cast<uint64_t *>(smem + 96);            // PBS5_SPLIT_BSK: off-diagonal half
    unsigned int *consumed_o = reinterpret_cast<unsigned int *>(smem + 104);
    uint64_t *bar_free = reinterpret_cast<uint64_t *>(smem + 32) + 2 * ctl;   // both warps are done reading their own buffer
    uint64_t *bar_full = bar_free + 1;                                        // both warps have written their transform
    double2 *bsk_s = reinterpret_cast<double2 *>(smem + kHdr5Bytes);
    unsigned char *ctbase = smem + kHdr5Bytes + kBskSliceBytes + (size_t)ctl * (2 * kBuf5Bytes + kMaxSmallDim * sizeof(uint16_t));
    double2 *tb_own = reinterpret_cast<double2 *>(ctbase + (size_t)p * kBuf5Bytes);
    double2 *tb_sib = reinterpret_cast<double2 *>(ctbase + (size_t)(1 - p) * kBuf5Bytes);
    uint16_t *ahat = reinterpret_cast<uint16_t *>(ctbase + (size_t)2 * kBuf5Bytes);
    uint64_t *rot = reinterpret_cast<uint64_t *>(tb_own);   // rotation copy (G = -acc, + overflow zone) aliases the transposition buffer

    // ---------------------------------------------------------------- CTA setup
    constexpr uint32_t kTmemCols = kCts <= 2 ? 256u : 512u;   // twiddles + 128 accumulator columns per warp of a quadrant
    if (warp == 0) tmem_alloc(slot, kTmemCols);
    constexpr int kSplit = pbs5_split<kCts>();
    if (threadIdx.x == 0) {
        mbar_init(bsk_bar, 1);
        *consumed = 0;
        if (kSplit) {
            mbar_init(bsk_bar_o, 1);
            *consumed_o = 0;
        }
    }
    if (lane == 0 && p == 0) {
        mbar_init(bar_free, 2);
        mbar_init(bar_full, 2);
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *slot;
    const uint32_t tquad = tbase + (((uint32_t)(warp & 3) * 32u) << 16);
    const uint32_t t_acc = tquad + kTmemAcc0 + (uint32_t)(warp >> 2) * 128u;
    const uint32_t t_xt_out = tquad + kTmemXt0 + (uint32_t)p * 64u, t_xt_in = tquad + kTmemXt0 + (uint32_t)(1 - p) * 64u;   // kXT only
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
    const int n_act_cts = min(kCts, a.batch - (int)blockIdx.x * kCts);
    const unsigned int n_act_warps = 2u * (unsigned int)n_act_cts;
    const bool dephase = kPhase > 0 && kCts == 4 && n_act_cts == 4;   // CTA-uniform
    const bool late = warp >= 4;
    if (threadIdx.x == 0) {
        if (kSplit) { issue_bsk_half(bsk_s, a.bsk, 0, 0, bsk_bar); issue_bsk_half(bsk_s, a.bsk, 0, 1, bsk_bar_o); }
        else issue_bsk_slice(bsk_s, a.bsk, 0, bsk_bar);
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();

    if (active) {
        // ---------------------------------------------------------------- prologue
        const uint64_t *lwe = a.lwe_small + (size_t)ct * (a.n + 1);
        for (int i = p * 32 + lane; i < a.n; i += 64) ahat[i] = (uint16_t)modswitch2048(lwe[i]);
        const uint32_t bhat = modswitch2048(lwe[a.n]);
        const uint64_t *lut = a.luts + ((size_t)pbs_lut_id(a, ct) * 2 + p) * kN;
        // acc = LUT * X^-b~: polynomial_wrapping_monic_monomial_div (polynomial_algorithms.rs:315-354); stored as G = -acc
#pragma unroll
        for (int c = 0; c < 8; c++) {
            uint32_t h[16];
#pragma unroll
            for (int mm = 0; mm < 4; mm++) {
                const int m = slot5(c, mm);
                const int j = lane + 32 * m;
                const uint32_t i0 = (uint32_t)(j + bhat) & 4095u, i1 = (i0 + 1024u) & 4095u;
                uint64_t v0 = lut[i0 & 2047u], v1 = lut[i1 & 2047u];
                if (i0 & 2048u) v0 = 0 - v0;
                if (i1 & 2048u) v1 = 0 - v1;
                const uint64_t g0 = 0 - v0, g1 = 0 - v1;
                rot[j] = g0; rot[j + kHalf] = g1;
                if (m < 8) rot[kN + j] = v0;
                h[4 * mm] = (uint32_t)g0; h[4 * mm + 1] = (uint32_t)(g0 >> 32);
                h[4 * mm + 2] = (uint32_t)g1; h[4 * mm + 3] = (uint32_t)(g1 >> 32);
            }
            tmem_st16(t_acc + c * 16, h);
        }
        tmem_wait_st();
        ct_barrier(1 + ctl);  // a~ table visible to both warps; rotation copy visible within the warp

#ifdef B200TFHE_LAB_DELAY
        { const long long t_in = clock64(); while (clock64() - t_in < (long long)g_lab_delay * ctl) { } }   // ciphertext c starts c * delay cycles late
#endif
        // ---------------------------------------------------------------- CMUX loop
        // Steps with a~ = 0 (mod 2N) are not skipped as the reference does (bootstrap.rs:281): the rotation is then the
        // identity, every digit is 0 and the step adds exactly zero.
        for (int i = 0; i < a.n; i++) {
            const uint32_t par = (uint32_t)(i & 1);
            if (dephase) {
                if (late) bar_sync_n(5, 256);           // the early half has reached the middle of step i
                else if (i > 0) bar_sync_n(6, 256);     // the late half has reached the middle of step i - 1
            }
            PBS5_TS(0);
            double xr[32], xi[32];
            // phase A: ct1 = acc * X^a~ - acc, round + digit (ggsw.rs:514-521), exact int -> double, twist by C_m
            {
                const uint32_t q0 = (4096u - (uint32_t)ahat[i]) & 4095u;   // source index (in the 4096-periodic extension) of coefficient 0
                const uint64_t *lanebase = rot + lane;
                // group g = slots 8g .. 8g+7 (g < 4: coefficients < 1024, g >= 4: the second half): one segment / offset for
                // all lanes and all 8 slots.  Stored word is R = -acc_src: segment 0 -> V = +acc_src = ~R + 1 (mask ~0),
                // segment 1 -> V = -acc_src = R (mask 0).
                const uint64_t *gp[8];
                uint32_t gt[8];
#pragma unroll
                for (int g = 0; g < 8; g++) {
                    const uint32_t qg = (q0 + 256u * (uint32_t)g) & 4095u;
                    gp[g] = lanebase + (qg & 2047u);
                    gt[g] = (qg >> 11) - 1u;
                }
                uint32_t h0[16], h1[16];
                tmem_ld16_nc(t_acc, h0);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t(&h)[16] = (c & 1) ? h1 : h0;
                    tmem_wait_ld16(h);
                    if (c < 7) tmem_ld16_nc(t_acc + (c + 1) * 16, (c & 1) ? h0 : h1);
#pragma unroll
                    for (int mm = 0; mm < 4; mm++) {
                        const int m = slot5(c, mm);
                        const int g = m >> 3, sl = m & 7;
                        const uint32_t tr = gt[g], ti = gt[4 + g];
                        const uint64_t r0 = gp[g][32 * sl], r1 = gp[4 + g][32 * sl];
                        const uint64_t e0 = pack64(h[4 * mm], h[4 * mm + 1]) + (r0 ^ pack64(tr, tr)) + pack64(tr & 1u, 0x7FFFFF00u);
                        const uint64_t e1 = pack64(h[4 * mm + 2], h[4 * mm + 3]) + (r1 ^ pack64(ti, ti)) + pack64(ti & 1u, 0x7FFFFF00u);
                        const uint32_t d0 = (uint32_t)(e0 >> 41), d1 = (uint32_t)(e1 >> 41);
                        // digit + (2^22 - 1) in [0, 2^23) -> double by exponent trick (exact)
#if PBS5_I2F
                        double fr = (double)((int)d0 - 4194303), fi = (double)((int)d1 - 4194303);   // conversion unit instead of the FP64 pipe
#else
                        double fr = dbl(d0, 0x43300000u) - 4503599631564799.0;
                        double fi = dbl(d1, 0x43300000u) - 4503599631564799.0;
#endif
                        PBS_DUMP(0, d0); PBS_DUMP(1, d1);
#if !PBS5_TWFOLD
                        twist_m(fr, fi, m);
#endif
                        xr[brev5(m)] = fr; xi[brev5(m)] = fi;
                    }
                }
            }
            PBS5_TS(1);

            // forward transform (fft.cuh: fwd1024), with the hand-over points of the pair protocol
            {
                uint32_t t0[16], t1[16];
#if PBS5_TWFOLD
                fft32_dit_twisted(xr, xi);
#else
                fft32_dit<false>(xr, xi);
#endif
                __syncwarp();  // all rotation reads done before the buffer is reused for the transposition
                tw.issue(0, t0);
#pragma unroll
                for (int c2 = 0; c2 < 4; c2++) {
                    tw.wait();
                    tw.issue(2 * c2 + 1, t1);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const int k1 = 8 * c2 + kk;
                        double2 y;
                        cmul_tw<false>(y.x, y.y, xr[k1], xi[k1], t0, kk);
                        tb_own[lane * kTStride + k1] = y;
                    }
                    tw.wait();
                    if (c2 < 3) tw.issue(2 * c2 + 2, t0);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const int k1 = 8 * c2 + 4 + kk;
                        double2 y;
                        cmul_tw<false>(y.x, y.y, xr[k1], xi[k1], t1, kk);
                        tb_own[lane * kTStride + k1] = y;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int l = 0; l < 32; l++) {
                    const double2 v = tb_own[l * kTStride + lane];
                    xr[brev5(l)] = v.x; xi[brev5(l)] = v.y;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_free);   // this warp's buffer may now receive the sibling's transform
                fft32_dit_stages<false, 0, (PBS5_FUSE & 1) ? 3 : 4>(xr, xi);
            }
            PBS5_TS(2);

            // hand the transform to the sibling (written into ITS buffer, layout [q][lane]); Out_p = B[p][p] F_p +
            // B[1-p][p] F_{1-p} (update_with_fmadd, ggsw.rs:616-697), written into the bit-reversed slot the inverse
            // transform wants.  The own product needs nothing from the sibling and hides the hand-over.
            mbar_wait(bar_free, par);
            mbar_wait(bsk_bar, par);
            PBS5_TS(3);
            PBS5_TS(4);
            double zr[32], zi[32];
            {
                const double2 *b_own = bsk_s + (size_t)(p * 2 + p) * kHalf + lane;         // row p, column p
                if (kXT) {   // blocks q < kXtQ of the transform go to the sibling through tensor memory
#pragma unroll
                    for (int c = 0; c < kXtQ / 4; c++) {
                        uint32_t sw[16];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            undbl(xr[4 * c + k], sw[4 * k], sw[4 * k + 1]);
                            undbl(xi[4 * c + k], sw[4 * k + 2], sw[4 * k + 3]);
                        }
                        tmem_st16_nc(t_xt_out + c * 16, sw);
                    }
                }
                own_products<0, (PBS5_FUSE & 1) != 0, kXT>(zr, zi, xr, xi, b_own, tb_sib + lane);
                if (kXT) {
                    tmem_wait_st();
                    tmem_fence_before();
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full);
            if (kSplit && lane == 0) {   // this warp is done with the diagonal half; the last of the CTA's warps refills it
                const unsigned int old = atomicAdd(consumed, 1u);
                if (old == (unsigned int)(i + 1) * n_act_warps - 1u && i + 1 < a.n) {
                    issue_bsk_half(bsk_s, a.bsk, i + 1, 0, bsk_bar);
#ifdef B200TFHE_LAB_BSKLAT   // development: latency of the refill, measured by spinning on the barrier (perturbs the kernel)
                    if (a.dbg && blockIdx.x < 8 && i >= 100 && i < 116) {
                        const long long t0 = clock64();
                        mbar_wait(bsk_bar, (uint32_t)((i + 1) & 1));
                        a.dbg[blockIdx.x * 16 + (i - 100)] = clock64() - t0;
                    }
#endif
                }
            }
            if (dephase) {
                if (!late) bar_arrive(5, 256);
                else if (i + 1 < a.n) bar_arrive(6, 256);
            }
            PBS5_TS(5);
            mbar_wait(bar_full, par);
            if (kSplit) mbar_wait(bsk_bar_o, par);
            PBS5_TS(6);
            {
                const double2 *b_oth = bsk_s + (size_t)((1 - p) * 2 + p) * kHalf + lane;   // row 1-p, column p
                uint32_t xt[kXtQ / 4 + 1][16];
                if (kXT) {
                    tmem_fence_after();
#pragma unroll
                    for (int c = 0; c < kXtQ / 4; c++) tmem_ld16_nc(t_xt_in + c * 16, xt[c]);
#pragma unroll
                    for (int c = 0; c < kXtQ / 4; c++) tmem_wait_ld16(xt[c]);
                }
                const SiblingProduct<kXT> oth{b_oth, tb_own + lane, xt};
                if (PBS5_FUSE & 2) {
                    fft32_dit_df<true, 0, 32>(zr, zi, oth);   // inverse first pass (registers only), products at the leaves
                } else {
                    sibling_products<0, kXT>(zr, zi, oth);
                    if (kSplit == 2) {   // off-diagonal half released BEFORE the inverse first pass
                        __syncwarp();
                        if (lane == 0) {
                            const unsigned int old = atomicAdd(consumed_o, 1u);
                            if (old == (unsigned int)(i + 1) * n_act_warps - 1u && i + 1 < a.n)
                                issue_bsk_half(bsk_s, a.bsk, i + 1, 1, bsk_bar_o);
                        }
                    }
                    inv1024_pass1(zr, zi);
                }
            }
            PBS5_TS(7);
            // this warp is done with the slice (and with the sibling's transform); the last of the CTA's warps refills the slice buffer
            __syncwarp();
            if (kSplit != 2 && lane == 0) {
                if (kSplit) {
                    const unsigned int old = atomicAdd(consumed_o, 1u);
                    if (old == (unsigned int)(i + 1) * n_act_warps - 1u && i + 1 < a.n)
                        issue_bsk_half(bsk_s, a.bsk, i + 1, 1, bsk_bar_o);
                } else {
                    const unsigned int old = atomicAdd(consumed, 1u);
                    if (old == (unsigned int)(i + 1) * n_act_warps - 1u && i + 1 < a.n)
                        issue_bsk_slice(bsk_s, a.bsk, i + 1, bsk_bar);
                }
            }
            PBS5_TS(8);
            inv1024_rest(zr, zi, tb_own, tw, lane);   // (its transposition reads are complete before its second pass starts)
            PBS5_TS(9);

            // phase D: untwist, from_torus, wrapping add (fft/mod.rs:285-304): acc += delta <=> G -= delta; refresh
            // TMEM, the rotation copy and its overflow zone (negated first 256 coefficients)
            {
                uint32_t h0[16], h1[16];
                tmem_ld16_nc(t_acc, h0);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t(&h)[16] = (c & 1) ? h1 : h0;
                    tmem_wait_ld16(h);
                    if (c < 7) tmem_ld16_nc(t_acc + (c + 1) * 16, (c & 1) ? h0 : h1);
#pragma unroll
                    for (int mm = 0; mm < 4; mm++) {
                        const int m = slot5(c, mm);
                        const int j = lane + 32 * m;
                        double yr = zr[m], yi = zi[m];
                        untwist_m(yr, yi, m);
                        PBS_DUMP(2, __double_as_longlong(yr)); PBS_DUMP(3, __double_as_longlong(yi));
                        PBS_DUMP(4, from_torus_fp(yr) - kFtBias); PBS_DUMP(5, from_torus_fp(yi) - kFtBias);
                        acc_sub_from_torus(h[4 * mm], h[4 * mm + 1], yr);
                        acc_sub_from_torus(h[4 * mm + 2], h[4 * mm + 3], yi);
                        sts_v2(rot + j, h[4 * mm], h[4 * mm + 1]);
                        sts_v2(rot + j + kHalf, h[4 * mm + 2], h[4 * mm + 3]);
                        if (m < 8) rot[kN + j] = 0 - pack64(h[4 * mm], h[4 * mm + 1]);
                    }
                    tmem_st16_nc(t_acc + c * 16, h);
                }
                tmem_wait_st();
            }
            __syncwarp();  // rotation copy complete before the next step's gather
            PBS5_TS(10);
        }

        // ---------------------------------------------------------------- sample extraction (acc = -G)
        uint64_t *o = a.out + (size_t)ct * (kN + 1);
        if (p == 0) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                uint32_t h[16];
                tmem_ld16(t_acc + c * 16, h);
                tmem_wait_ld();
#pragma unroll
                for (int mm = 0; mm < 4; mm++) {
                    const int j = lane + 32 * slot5(c, mm);
                    const uint64_t g0 = pack64(h[4 * mm], h[4 * mm + 1]);       // = -acc[j]
                    const uint64_t g1 = pack64(h[4 * mm + 2], h[4 * mm + 3]);   // = -acc[j + 1024]
                    if (j == 0) o[0] = 0 - g0; else o[kN - j] = g0;
                    o[kHalf - j] = g1;  // coefficient j + 1024 -> index N - (j + 1024), negated
                }
            }
        } else {
            uint32_t h[16];
            tmem_ld16(t_acc, h);
            tmem_wait_ld();
            if (lane == 0) o[kN] = 0 - pack64(h[0], h[1]);
        }
    }

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, kTmemCols);
}

}  // namespace b200
