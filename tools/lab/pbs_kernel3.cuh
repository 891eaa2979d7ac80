// pbs_kernel3.cuh -- third generation of the batched programmable bootstrap (same contract and
// reference citations as pbs_kernel.cuh; 4 ciphertexts per CTA, one warp per GLWE polynomial).
//
// What changed against pbs_kernel<4, true>:
//   * The two warps of a ciphertext exchange their TRANSFORMS F_p (not partial products) through
//     shared memory; warp p then forms its own output polynomial Out_p = B[p][p] F_p + B[1-p][p] F_{1-p}
//     (update_with_fmadd, ggsw.rs:616-697) in place: no separate accumulation adds and half the
//     live registers in the multiply.
//   * The accumulator is kept as G = C - acc with C = 2^63 - 2^40 (TMEM, home layout) and as
//     r = -acc (shared memory, rotation copy).  Then for x = rot - acc the reference's
//     closest_representable + balanced digit (decomposer.rs:98-116, iter.rs:120-127) is exactly
//         digit = (hi32(G + (+-rot)) >> 9) - (2^22 - 1)
//     i.e. one 64-bit add, one shift and one exact int -> double conversion by exponent trick; the
//     tie (x = 2^63 - 2^40 mod 2^64 -> +2^22, not -2^22) comes out right by construction.
//   * The rotated gather uses the fact that, for one lane, the 32 indices (idx0 + 32 m) mod 2N cross
//     a multiple of N at most once: one compare per element selects both the wrapped address and the
//     negacyclic sign (polynomial_algorithms.rs:425-491).
//   * from_torus scales by 2^64 with an exponent add instead of a DMUL (the FP64 pipe is the
//     binding resource).
#pragma once
#include "../../tfhe_rs_string_b200/csrc/pbs_common.cuh"

namespace b200 {

// kCts3 ciphertexts per CTA: 4 for throughput; 1..3 for small batches (fewer warps per SM sub-partition
// shorten the per-step critical path, see launch_pbs3), always one CTA per SM.
template <int kCts3>
__host__ __device__ constexpr size_t pbs3_smem_bytes() {
    return kPbsHeaderBytes + kBskSliceBytes + (size_t)kCts3 * pbs_ct_smem_bytes();
}

// kPhase: 0 = all warps in lock step; p > 0 (4 ciphertexts per CTA only) = the two halves of the CTA run one phase
// apart: warps 4-7 (which share their SM sub-partitions with warps 0-3) start the gather of step i when warps 0-3
// have finished theirs, and warps 0-3 start step i+1 when warps 4-7 have finished the inverse transform of step i.
// Every phase of a CMUX step is bound by a different unit (gather: integer ALU + instruction fetch, transforms: FP64
// pipe, exchange/multiply: shared memory, from_torus: conversion unit), so the two warps of a sub-partition then
// always ask for different units.  Hand-over by named barriers 5 and 6 (bar.arrive / bar.sync, 256 threads).
template <int kCts3, int kPhase = 0>
__global__ void __launch_bounds__(kCts3 * 64, 1) pbs_kernel3(const PbsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctl = warp >> 1, p = warp & 1;
    const int ct = blockIdx.x * kCts3 + ctl;
    const bool active = ct < a.batch;

    uint32_t *slot = reinterpret_cast<uint32_t *>(smem);
    uint64_t *bsk_bar = reinterpret_cast<uint64_t *>(smem + 8);
    unsigned int *consumed = reinterpret_cast<unsigned int *>(smem + 16);
    double2 *bsk_s = reinterpret_cast<double2 *>(smem + kPbsHeaderBytes);
    unsigned char *ctbase = smem + kPbsHeaderBytes + kBskSliceBytes + (size_t)ctl * pbs_ct_smem_bytes();
    double2 *tb_own = reinterpret_cast<double2 *>(ctbase) + p * kTBufElems;
    const double2 *tb_oth = reinterpret_cast<double2 *>(ctbase) + (1 - p) * kTBufElems;
    uint16_t *ahat = reinterpret_cast<uint16_t *>(ctbase + (size_t)2 * kTBufElems * sizeof(double2));
    uint64_t *rot = reinterpret_cast<uint64_t *>(tb_own);  // rotation copy (r = -acc) aliases the transposition buffer

    // ---------------------------------------------------------------- CTA setup
    constexpr uint32_t kTmemCols = kCts3 <= 2 ? 256u : 512u;   // twiddles + 128 accumulator columns per warp of a quadrant
    if (warp == 0) tmem_alloc(slot, kTmemCols);
    if (threadIdx.x == 0) {
        mbar_init(bsk_bar, 1);
        *consumed = 0;
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *slot;
    const uint32_t tquad = tbase + (((uint32_t)(warp & 3) * 32u) << 16);
    const uint32_t t_acc = tquad + kTmemAcc0 + (uint32_t)(warp >> 2) * 128u;
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
    const int n_act_cts = min(kCts3, a.batch - (int)blockIdx.x * kCts3);
    const bool dephase = kPhase > 0 && kCts3 == 4 && n_act_cts == 4;   // CTA-uniform
    const bool late = warp >= 4;
    const unsigned int n_act_warps = 2u * (unsigned int)n_act_cts;
    if (threadIdx.x == 0) issue_bsk_slice(bsk_s, a.bsk, 0, bsk_bar);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();

    if (active) {
        // ---------------------------------------------------------------- prologue
        const uint64_t *lwe = a.lwe_small + (size_t)ct * (a.n + 1);
        for (int i = p * 32 + lane; i < a.n; i += 64) ahat[i] = (uint16_t)modswitch2048(lwe[i]);
        const uint32_t bhat = modswitch2048(lwe[a.n]);
        const uint64_t *lut = a.luts + ((size_t)pbs_lut_id(a, ct) * 2 + p) * kN;
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
                rot[j] = 0 - v0; rot[j + kHalf] = 0 - v1;
                const uint64_t g0 = kAccC - v0, g1 = kAccC - v1;
                h[4 * mm] = (uint32_t)g0; h[4 * mm + 1] = (uint32_t)(g0 >> 32);
                h[4 * mm + 2] = (uint32_t)g1; h[4 * mm + 3] = (uint32_t)(g1 >> 32);
            }
            tmem_st16(t_acc + c * 16, h);
        }
        tmem_wait_st();
        ct_barrier(1 + ctl);  // a~ table visible to both warps; rot copy visible within the warp

        // ---------------------------------------------------------------- CMUX loop
        // Steps with a~ = 0 (mod 2N) are not skipped as the reference does (bootstrap.rs:281): the
        // rotation is then the identity, every digit is 0 and the step adds exactly zero.
#ifdef B200TFHE_LAB_DUP_LOOP
        // development experiment (tools/lab): the upper half of the CTA runs a second copy of the same loop, so the two
        // warps of a sub-partition fetch different addresses while executing the same phases
#if B200TFHE_LAB_DUP_LOOP == 2
        if (p) {     // copies split by sub-partition parity: the two warps of a sub-partition share one copy
#else
        if (late) {  // copies split inside every sub-partition
#endif
#include "pbs_kernel3_loop.inc"
        } else {
#include "pbs_kernel3_loop.inc"
        }
#else
#include "pbs_kernel3_loop.inc"
#endif

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
                    const uint64_t a0 = kAccC - pack64(h[4 * mm], h[4 * mm + 1]);
                    const uint64_t a1 = kAccC - pack64(h[4 * mm + 2], h[4 * mm + 3]);
                    if (j == 0) o[0] = a0; else o[kN - j] = 0 - a0;
                    o[kHalf - j] = 0 - a1;  // coefficient j + 1024 -> index N - (j + 1024)
                }
            }
        } else {
            uint32_t h[16];
            tmem_ld16(t_acc, h);
            tmem_wait_ld();
            if (lane == 0) o[kN] = kAccC - pack64(h[0], h[1]);
        }
    }

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, kTmemCols);
}

}  // namespace b200
