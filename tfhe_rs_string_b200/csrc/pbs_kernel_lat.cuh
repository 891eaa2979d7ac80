// pbs_kernel_lat.cuh -- latency variant of the programmable bootstrap for small batches (same contract
// and reference citations as pbs_kernel5.cuh).
//
// A CMUX step is a serial dependency chain per polynomial; with one warp per polynomial that chain
// is 14 k cycles.  Here a polynomial gets TWO warps (warps q and q + 4, same TMEM quadrant q): thread
// (lane l, half h) owns the 16 folded points l + 32 m, m in [16 h, 16 h + 16).  The 32-point passes of
// the 32 x 32 FFT become a cross-warp radix-2 stage plus a 16-point transform in registers:
//   forward (DIF first):  t = own + s * recv (s = +1 / -1 for h = 0 / 1); h = 1 multiplies by -W32^mm;
//                         16-point DIT -> outputs with index 2 kappa + h
//   inverse (DIT last):   16-point DIT on the inputs with index 2 kappa + h; h = 1 multiplies by
//                         conj(W32^mm); out = recv + s * own
// The 16 values cross between the two warps through TMEM (lane private: thread (l, 0) <-> (l, 1)), the
// lane <-> register transposition between the passes stays in shared memory (shared by the two
// warps).  Frequency layout is unchanged: thread (lane k1, h) holds F[k1 + 32 (2 kappa + h)], so the
// Fourier BSK produced by bsk_to_fourier_kernel is used as is.  tools/proto_fft16x2.py is the numpy
// model of this data flow.
//
// Mapping: 8 warps per CTA, quadrant q = warp & 3 hosts one polynomial: ciphertext q >> 1, polynomial
// q & 1, half h = warp >> 2; 1 or 2 ciphertexts per CTA (quadrants 2, 3 idle for 1).
#pragma once
#include "pbs_kernel5.cuh"

namespace b200 {

constexpr uint32_t kLatAcc = 128, kLatX = 256;   // TMEM columns: [0,128) twiddles (64 per half), acc + 64 h, exchange + 64 h
__host__ __device__ constexpr int brev4(int x) { return ((x & 1) << 3) | ((x & 2) << 1) | ((x & 4) >> 1) | ((x & 8) >> 3); }

// In-register 16-point DFT, decimation in time: input in bit-reversed register order, output natural.
template <bool INV>
__device__ __forceinline__ void fft16_dit(double (&xr)[16], double (&xi)[16]) {
#pragma unroll
    for (int half = 1; half < 16; half <<= 1) {
#pragma unroll
        for (int base = 0; base < 16; base += 2 * half) {
#pragma unroll
            for (int t = 0; t < half; t++)
                bfly<INV>(xr[base + t], xi[base + t], xr[base + t + half], xi[base + t + half], t * (16 / half));   // W16^(t*8/half) = W32^(t*16/half)
        }
    }
}

__device__ __forceinline__ void pair_barrier(const int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void quad_barrier(const int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// send 16 complex values to the sibling warp (64 TMEM columns at xout), receive its 16 (xin)
__device__ __forceinline__ void lat_send(const double (&vr)[16], const double (&vi)[16], const uint32_t xout) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t s[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            undbl(vr[4 * k + j], s[4 * j], s[4 * j + 1]);
            undbl(vi[4 * k + j], s[4 * j + 2], s[4 * j + 3]);
        }
        tmem_st16_nc(xout + k * 16, s);
    }
    tmem_wait_st();
    tmem_fence_before();
}
__device__ __forceinline__ void lat_recv(double (&vr)[16], double (&vi)[16], const uint32_t xin) {
    tmem_fence_after();
    uint32_t g0[16], g1[16];
    tmem_ld16_nc(xin, g0);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t(&g)[16] = (k & 1) ? g1 : g0;
        tmem_wait_ld16(g);
        if (k < 3) tmem_ld16_nc(xin + (k + 1) * 16, (k & 1) ? g0 : g1);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            vr[4 * k + j] = dbl(g[4 * j], g[4 * j + 1]);
            vi[4 * k + j] = dbl(g[4 * j + 2], g[4 * j + 3]);
        }
    }
}

// Two-round exchange for the 16-warp configuration, where only 32 TMEM columns per warp are left: 8
// values per round; round B writes into the buffer this warp has just read (the sibling does the
// same), so no write-after-read barrier is needed between the rounds.
__device__ __forceinline__ void lat_put8(const double (&vr)[16], const double (&vi)[16], const int off, const uint32_t x) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
        uint32_t s[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            undbl(vr[off + 4 * k + j], s[4 * j], s[4 * j + 1]);
            undbl(vi[off + 4 * k + j], s[4 * j + 2], s[4 * j + 3]);
        }
        tmem_st16_nc(x + k * 16, s);
    }
    tmem_wait_st();
    tmem_fence_before();
}
__device__ __forceinline__ void lat_get8(double (&vr)[16], double (&vi)[16], const int off, const uint32_t x) {
    tmem_fence_after();
    uint32_t g0[16], g1[16];
    tmem_ld16_nc(x, g0);
    tmem_ld16_nc(x + 16, g1);
    tmem_wait_ld16(g0);
    tmem_wait_ld16(g1);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        vr[off + j] = dbl(g0[4 * j], g0[4 * j + 1]); vi[off + j] = dbl(g0[4 * j + 2], g0[4 * j + 3]);
        vr[off + 4 + j] = dbl(g1[4 * j], g1[4 * j + 1]); vi[off + 4 + j] = dbl(g1[4 * j + 2], g1[4 * j + 3]);
    }
}
// swap 16 complex values with the sibling warp: send v, receive r.  MODE 0: one round through TMEM (siblings share a TMEM
// quadrant), 1: two rounds through TMEM (16-warp configuration), 2: through shared memory (siblings on different SM
// sub-partitions, see kSpread below; sx_own / sx_oth alternate between two buffers from call to call, so that a buffer is
// rewritten only after the pair barrier of the following swap).
template <int MODE>
__device__ __forceinline__ void lat_swap(const double (&vr)[16], const double (&vi)[16], double (&rr)[16], double (&ri)[16],
                                         const uint32_t xout, const uint32_t xin, double2 *sx_own, const double2 *sx_oth, const int bar) {
    if (MODE == 0) {
        lat_send(vr, vi, xout);
        pair_barrier(bar);
        lat_recv(rr, ri, xin);
    } else if (MODE == 1) {
        lat_put8(vr, vi, 0, xout);
        pair_barrier(bar);
        lat_get8(rr, ri, 0, xin);
        lat_put8(vr, vi, 8, xin);     // the buffer just read; the sibling fills xout
        pair_barrier(bar);
        lat_get8(rr, ri, 8, xout);
    } else {
#pragma unroll
        for (int k = 0; k < 16; k++) sx_own[k * 32] = make_double2(vr[k], vi[k]);
        pair_barrier(bar);
#pragma unroll
        for (int k = 0; k < 16; k++) { const double2 v = sx_oth[k * 32]; rr[k] = v.x; ri[k] = v.y; }
    }
}

// forward cross-warp stage + 16-point transform: x natural (own half) -> y[kappa] = output 2 kappa + h
template <int MODE>
__device__ __forceinline__ void lat_fwd_pass(const double (&xr)[16], const double (&xi)[16], double (&yr)[16], double (&yi)[16],
                                             const int h, const double sgn, const uint32_t xout, const uint32_t xin,
                                             double2 *sx_own, const double2 *sx_oth, const int bar) {
    double rr[16], ri[16];
    lat_swap<MODE>(xr, xi, rr, ri, xout, xin, sx_own, sx_oth, bar);
#pragma unroll
    for (int mm = 0; mm < 16; mm++) {
        double tr = fma(sgn, rr[mm], xr[mm]), ti = fma(sgn, ri[mm], xi[mm]);
        if (h && mm) {   // times -W32^mm (warp-uniform branch)
            const double wr = -w32_re(mm), wi = -w32_im(mm);
            const double nr = fma(-ti, wi, tr * wr);
            ti = fma(ti, wr, tr * wi);
            tr = nr;
        } else if (h) {
            tr = -tr; ti = -ti;
        }
        yr[brev4(mm)] = tr; yi[brev4(mm)] = ti;
    }
    fft16_dit<false>(yr, yi);
}
// inverse: x[brev4(kappa)] = inputs 2 kappa + h -> y[mm] = result for index mm + 16 h
template <int MODE>
__device__ __forceinline__ void lat_inv_pass(double (&xr)[16], double (&xi)[16], double (&yr)[16], double (&yi)[16],
                                             const int h, const double sgn, const uint32_t xout, const uint32_t xin,
                                             double2 *sx_own, const double2 *sx_oth, const int bar) {
    fft16_dit<true>(xr, xi);
    if (h) {
#pragma unroll
        for (int mm = 1; mm < 16; mm++) {   // times conj(W32^mm)
            const double wr = w32_re(mm), wi = -w32_im(mm);
            const double nr = fma(-xi[mm], wi, xr[mm] * wr);
            xi[mm] = fma(xi[mm], wr, xr[mm] * wi);
            xr[mm] = nr;
        }
    }
    lat_swap<MODE>(xr, xi, yr, yi, xout, xin, sx_own, sx_oth, bar);
#pragma unroll
    for (int mm = 0; mm < 16; mm++) {
        yr[mm] = fma(sgn, xr[mm], yr[mm]);
        yi[mm] = fma(sgn, xi[mm], yi[mm]);
    }
}

// per ciphertext: two polynomial buffers (rotation copy with its overflow zone / transposition / transform, see
// pbs_kernel5.cuh) and the a~ table
__host__ __device__ constexpr size_t pbs_lat_ct_bytes() { return (size_t)2 * kBuf5Bytes + kMaxSmallDim * sizeof(uint16_t); }
constexpr int kLatSxBytes = 2 * 2 * 16 * 32 * (int)sizeof(double2);   // per polynomial: two buffers x two halves x 16 values x 32 lanes
template <int CTS, bool kSpread = false>
__host__ __device__ constexpr size_t pbs_lat_smem_bytes() {
    return kPbsHeaderBytes + kBskSliceBytes + (size_t)CTS * pbs_lat_ct_bytes() + (kSpread ? 2 * kLatSxBytes : 0);
}

// CTS = 1, 2: 8 warps (latency); CTS = 4: 16 warps, 128 registers per thread, two polynomial pairs per
// TMEM quadrant (ciphertexts c and c + 2 share the sub-partitions), two-round exchanges (throughput).
// kSpread (one ciphertext per CTA, 4 warps): the two halves of a polynomial run on DIFFERENT SM sub-partitions (warp = p + 2 h,
// one warp per sub-partition and TMEM quadrant) and swap through shared memory.  With both halves on one sub-partition
// (the TMEM exchange needs a common quadrant) the 2,860 FP64 instructions of a polynomial's CMUX step share one FP64 pipe
// and two of the SM's four pipes idle; spread out, every pipe carries 1,430.
template <int CTS, bool kSpread = false>
__global__ void __launch_bounds__(kSpread ? 128 : CTS == 4 ? 512 : 256, 1) pbs_lat_kernel(const PbsArgs a) {
    static_assert(CTS == 1 || CTS == 2 || CTS == 4, "1, 2 or 4 ciphertexts per CTA");
    static_assert(!kSpread || CTS == 1, "the spread mapping is for one ciphertext per CTA");
    constexpr bool kTwoRounds = CTS == 4;
    constexpr int kX = kSpread ? 2 : kTwoRounds ? 1 : 0;
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qd = warp & 3, h = kSpread ? (warp >> 1) : (warp >> 2) & 1, pr = kSpread ? 0 : warp >> 3;
    const int ctl = kSpread ? 0 : 2 * pr + (qd >> 1), p = kSpread ? (warp & 1) : (qd & 1);
    const int ct = blockIdx.x * CTS + ctl;
    const bool active = ctl < CTS && ct < a.batch;   // (kSpread: all four warps belong to the CTA's one ciphertext)
    const double sgn = h ? -1.0 : 1.0;

    uint32_t *slot = reinterpret_cast<uint32_t *>(smem);
    uint64_t *bsk_bar = reinterpret_cast<uint64_t *>(smem + 8);
    unsigned int *consumed = reinterpret_cast<unsigned int *>(smem + 16);
    double2 *bsk_s = reinterpret_cast<double2 *>(smem + kPbsHeaderBytes);
    unsigned char *ctbase = smem + kPbsHeaderBytes + kBskSliceBytes + (size_t)(ctl < CTS ? ctl : 0) * pbs_lat_ct_bytes();
    double2 *tb_own = reinterpret_cast<double2 *>(ctbase + (size_t)p * kBuf5Bytes);            // shared by the two warps of the polynomial
    const double2 *tb_oth = reinterpret_cast<const double2 *>(ctbase + (size_t)(1 - p) * kBuf5Bytes);
    uint16_t *ahat = reinterpret_cast<uint16_t *>(ctbase + (size_t)2 * kBuf5Bytes);
    uint64_t *rot = reinterpret_cast<uint64_t *>(tb_own);   // rotation copy (G = -acc, + overflow zone, pbs_kernel5.cuh) aliases the transposition buffer

    if (warp == 0) tmem_alloc(slot, 512);
    if (threadIdx.x == 0) {
        mbar_init(bsk_bar, 1);
        *consumed = 0;
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *slot;
    const uint32_t tquad = tbase + (((uint32_t)qd * 32u) << 16);
    const uint32_t t_tw = tquad + (uint32_t)h * 64u;
    const uint32_t t_acc = tquad + kLatAcc + (uint32_t)(pr * 2 + h) * 64u;
    const uint32_t t_x_own = kTwoRounds ? tquad + 384u + (uint32_t)(pr * 2 + h) * 32u : tquad + kLatX + (uint32_t)h * 64u;
    const uint32_t t_x_oth = kTwoRounds ? tquad + 384u + (uint32_t)(pr * 2 + 1 - h) * 32u : tquad + kLatX + (uint32_t)(1 - h) * 64u;
    const int bar_pair = kSpread ? 1 + p : 1 + qd + 4 * pr, bar_ct = 9 + ctl;
    // kSpread: shared-memory swap buffers of this polynomial, [buffer b][half][16][32 lanes]
    double2 *sx = reinterpret_cast<double2 *>(smem + kPbsHeaderBytes + kBskSliceBytes + (size_t)CTS * pbs_lat_ct_bytes() + (size_t)p * kLatSxBytes);
    double2 *sx_own0 = sx + (0 * 2 + h) * 512 + lane, *sx_own1 = sx + (1 * 2 + h) * 512 + lane;
    const double2 *sx_oth0 = sx + (0 * 2 + 1 - h) * 512 + lane, *sx_oth1 = sx + (1 * 2 + 1 - h) * 512 + lane;
    if (pr == 0) {   // the quadrant's inter-pass twiddles T'[2 kappa + h][lane], one table per half
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t r[16];
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const double2 t = __ldg(a.twid + (2 * (4 * c + kk) + h) * 32 + lane);
                undbl(t.x, r[4 * kk], r[4 * kk + 1]);
                undbl(t.y, r[4 * kk + 2], r[4 * kk + 3]);
            }
            tmem_st16(t_tw + c * 16, r);
        }
        tmem_wait_st();
    }
    const int n_act_cts = min(CTS, a.batch - (int)blockIdx.x * CTS);
    const unsigned int n_act_warps = 4u * (unsigned int)n_act_cts;
    if (threadIdx.x == 0) issue_bsk_slice(bsk_s, a.bsk, 0, bsk_bar);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();

    if (active) {
        // ---------------------------------------------------------------- prologue
        const uint64_t *lwe = a.lwe_small + (size_t)ct * (a.n + 1);
        for (int i = (p * 2 + h) * 32 + lane; i < a.n; i += 128) ahat[i] = (uint16_t)modswitch2048(lwe[i]);
        const uint32_t bhat = modswitch2048(lwe[a.n]);
        const uint64_t *lut = a.luts + ((size_t)pbs_lut_id(a, ct) * 2 + p) * kN;
        // acc = LUT * X^-b~ (polynomial_algorithms.rs:315-354); TMEM and the rotation copy hold G = -acc
#pragma unroll
        for (int c = 0; c < 4; c++) {
            uint32_t hh[16];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int j = lane + 32 * (16 * h + 4 * c + k);
                const uint32_t i0 = (uint32_t)(j + bhat) & 4095u, i1 = (i0 + 1024u) & 4095u;
                uint64_t v0 = lut[i0 & 2047u], v1 = lut[i1 & 2047u];
                if (i0 & 2048u) v0 = 0 - v0;
                if (i1 & 2048u) v1 = 0 - v1;
                const uint64_t g0 = 0 - v0, g1 = 0 - v1;
                rot[j] = g0; rot[j + kHalf] = g1;
                if (j < kZone5) rot[kN + j] = v0;
                hh[4 * k] = (uint32_t)g0; hh[4 * k + 1] = (uint32_t)(g0 >> 32);
                hh[4 * k + 2] = (uint32_t)g1; hh[4 * k + 3] = (uint32_t)(g1 >> 32);
            }
            tmem_st16(t_acc + c * 16, hh);
        }
        tmem_wait_st();
        quad_barrier(bar_ct);   // a~ table and both rotation copies complete

        // ---------------------------------------------------------------- CMUX loop
        for (int i = 0; i < a.n; i++) {
            PBS3_TS(0);
            double xr[16], xi[16];
            // phase A: ct1 = acc * X^a~ - acc, round + digit, exact int -> double, twist by C_m (see pbs_kernel5.cuh)
            {
                // group-uniform gather of pbs_kernel5.cuh: this thread's slots 16 h .. 16 h + 15 of either half are two groups of 8
                const uint32_t q0 = (4096u - (uint32_t)ahat[i]) & 4095u;
                const uint64_t *lanebase = rot + lane;
                const uint64_t *gp[4];
                uint32_t gt[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {   // groups 2h, 2h + 1 (first half) and 4 + 2h, 5 + 2h (second half)
                    const uint32_t qg = (q0 + 256u * (uint32_t)(2 * h + (g & 1) + 4 * (g >> 1))) & 4095u;
                    gp[g] = lanebase + (qg & 2047u);
                    gt[g] = (qg >> 11) - 1u;
                }
                uint32_t h0[16], h1[16];
                tmem_ld16_nc(t_acc, h0);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t(&hh)[16] = (c & 1) ? h1 : h0;
                    tmem_wait_ld16(hh);
                    if (c < 3) tmem_ld16_nc(t_acc + (c + 1) * 16, (c & 1) ? h0 : h1);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int mm = 4 * c + k;
                        const int g = mm >> 3, sl = mm & 7;
                        const uint32_t tr = gt[g], ti = gt[2 + g];
                        const uint64_t r0 = gp[g][32 * sl], r1 = gp[2 + g][32 * sl];
                        const uint64_t e0 = pack64(hh[4 * k], hh[4 * k + 1]) + (r0 ^ pack64(tr, tr)) + pack64(tr & 1u, 0x7FFFFF00u);
                        const uint64_t e1 = pack64(hh[4 * k + 2], hh[4 * k + 3]) + (r1 ^ pack64(ti, ti)) + pack64(ti & 1u, 0x7FFFFF00u);
                        const double fr = dbl((uint32_t)(e0 >> 41), 0x43300000u) - 4503599631564799.0;
                        const double fi = dbl((uint32_t)(e1 >> 41), 0x43300000u) - 4503599631564799.0;
                        const double2 cm = c_twm[16 * h + mm];
                        xr[mm] = fma(-fi, cm.y, fr * cm.x);
                        xi[mm] = fma(fi, cm.x, fr * cm.y);
                    }
                }
            }
            // (the barrier inside the first pass also orders both warps' rotation reads before the buffer is reused)

            PBS3_TS(1);
            // ---- forward transform
            double yr[16], yi[16];
            lat_fwd_pass<kX>(xr, xi, yr, yi, h, sgn, t_x_own, t_x_oth, sx_own0, sx_oth0, bar_pair);
            {
                uint32_t t0[16], t1[16];
                tmem_ld16_nc(t_tw, t0);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t(&t)[16] = (c & 1) ? t1 : t0;
                    tmem_wait_ld16(t);
                    if (c < 3) tmem_ld16_nc(t_tw + (c + 1) * 16, (c & 1) ? t0 : t1);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const int kap = 4 * c + kk;
                        double2 y;
                        cmul_tw<false>(y.x, y.y, yr[kap], yi[kap], t, kk);
                        tb_own[lane * kTStride + 2 * kap + h] = y;
                    }
                }
            }
            PBS3_TS(2);
            pair_barrier(bar_pair);
#pragma unroll
            for (int ll = 0; ll < 16; ll++) {
                const double2 v = tb_own[(ll + 16 * h) * kTStride + lane];
                xr[ll] = v.x; xi[ll] = v.y;
            }
            lat_fwd_pass<kX>(xr, xi, yr, yi, h, sgn, t_x_own, t_x_oth, sx_own1, sx_oth1, bar_pair);   // its barrier also orders the transposition reads before the writes below

            PBS3_TS(3);
            // ---- exchange the transforms between the two polynomials, Out_p = B[p][p] F_p + B[1-p][p] F_{1-p}
#pragma unroll
            for (int kap = 0; kap < 16; kap++) tb_own[(2 * kap + h) * 32 + lane] = make_double2(yr[kap], yi[kap]);
            mbar_wait(bsk_bar, (uint32_t)(i & 1));
            double zr[16], zi[16];
            {
                const double2 *b_own = bsk_s + (size_t)(p * 2 + p) * kHalf + lane;
#pragma unroll
                for (int kap = 0; kap < 16; kap++) {
                    const double2 bo = b_own[(2 * kap + h) * 32];
                    zr[brev4(kap)] = fma(-bo.y, yi[kap], bo.x * yr[kap]);
                    zi[brev4(kap)] = fma(bo.y, yr[kap], bo.x * yi[kap]);
                }
            }
            PBS3_TS(4);
            quad_barrier(bar_ct);
            PBS3_TS(5);
            {
                const double2 *b_oth = bsk_s + (size_t)((1 - p) * 2 + p) * kHalf + lane;
#pragma unroll
                for (int kap = 0; kap < 16; kap++) {
                    const int q = 2 * kap + h;
                    const double2 bx = b_oth[q * 32], g = tb_oth[q * 32 + lane];
                    const double o_r = fma(bx.x, g.x, zr[brev4(kap)]), o_i = fma(bx.x, g.y, zi[brev4(kap)]);
                    zr[brev4(kap)] = fma(-bx.y, g.y, o_r); zi[brev4(kap)] = fma(bx.y, g.x, o_i);
                }
            }
            __syncwarp();
            if (lane == 0) {   // this warp is done with the slice; the last of the CTA's warps refills the buffer
                const unsigned int old = atomicAdd(consumed, 1u);
                if (old == (unsigned int)(i + 1) * n_act_warps - 1u && i + 1 < a.n)
                    issue_bsk_slice(bsk_s, a.bsk, i + 1, bsk_bar);
            }
            PBS3_TS(6);
            quad_barrier(bar_ct);   // the sibling polynomial has read this one's transform before the buffer is reused

            PBS3_TS(7);
            // ---- inverse transform
            lat_inv_pass<kX>(zr, zi, yr, yi, h, sgn, t_x_own, t_x_oth, sx_own0, sx_oth0, bar_pair);   // y[ll]: index l = ll + 16 h of lane k1
#pragma unroll
            for (int ll = 0; ll < 16; ll++) tb_own[lane * kTStride + ll + 16 * h] = make_double2(yr[ll], yi[ll]);
            pair_barrier(bar_pair);
            {
                uint32_t t0[16], t1[16];
                tmem_ld16_nc(t_tw, t0);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t(&t)[16] = (c & 1) ? t1 : t0;
                    tmem_wait_ld16(t);
                    if (c < 3) tmem_ld16_nc(t_tw + (c + 1) * 16, (c & 1) ? t0 : t1);
#pragma unroll
                    for (int kk = 0; kk < 4; kk++) {
                        const int kap = 4 * c + kk;
                        const double2 v = tb_own[(2 * kap + h) * kTStride + lane];
                        cmul_tw<true>(zr[brev4(kap)], zi[brev4(kap)], v.x, v.y, t, kk);
                    }
                }
            }
            lat_inv_pass<kX>(zr, zi, yr, yi, h, sgn, t_x_own, t_x_oth, sx_own1, sx_oth1, bar_pair);   // its barrier orders the transposition reads before the rotation copy below

            PBS3_TS(8);
            // ---- phase D: untwist, from_torus, G -= delta, refresh both copies
            {
                uint32_t h0[16], h1[16];
                tmem_ld16_nc(t_acc, h0);
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    uint32_t(&hh)[16] = (c & 1) ? h1 : h0;
                    tmem_wait_ld16(hh);
                    if (c < 3) tmem_ld16_nc(t_acc + (c + 1) * 16, (c & 1) ? h0 : h1);
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int mm = 4 * c + k;
                        const int j = lane + 32 * (mm + 16 * h);
                        const double2 cm = c_twm[16 * h + mm];
                        const double ur = fma(yi[mm], cm.y, yr[mm] * cm.x);
                        const double ui = fma(yi[mm], cm.x, -(yr[mm] * cm.y));
                        // acc += delta <=> G -= delta (from_torus on the FP64 pipe, pbs_kernel5.cuh)
                        const uint64_t g0 = pack64(hh[4 * k], hh[4 * k + 1]) + kFtBias - from_torus_fp(ur);
                        const uint64_t g1 = pack64(hh[4 * k + 2], hh[4 * k + 3]) + kFtBias - from_torus_fp(ui);
                        rot[j] = g0; rot[j + kHalf] = g1;
                        if (j < kZone5) rot[kN + j] = 0 - g0;   // (warp-uniform: h == 0, mm < 8)
                        hh[4 * k] = (uint32_t)g0; hh[4 * k + 1] = (uint32_t)(g0 >> 32);
                        hh[4 * k + 2] = (uint32_t)g1; hh[4 * k + 3] = (uint32_t)(g1 >> 32);
                    }
                    tmem_st16_nc(t_acc + c * 16, hh);
                }
                tmem_wait_st();
            }
            PBS3_TS(9);
            pair_barrier(bar_pair);   // rotation copy complete (both halves) before the next step's gather
            PBS3_TS(10);
        }

        // ---------------------------------------------------------------- sample extraction
        uint64_t *o = a.out + (size_t)ct * (kN + 1);
        if (p == 0) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t hh[16];
                tmem_ld16(t_acc + c * 16, hh);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int j = lane + 32 * (16 * h + 4 * c + k);
                    const uint64_t g0 = pack64(hh[4 * k], hh[4 * k + 1]);       // = -acc[j]
                    const uint64_t g1 = pack64(hh[4 * k + 2], hh[4 * k + 3]);   // = -acc[j + 1024]
                    if (j == 0) o[0] = 0 - g0; else o[kN - j] = g0;
                    o[kHalf - j] = g1;
                }
            }
        } else if (h == 0) {
            uint32_t hh[16];
            tmem_ld16(t_acc, hh);
            tmem_wait_ld();
            if (lane == 0) o[kN] = 0 - pack64(hh[0], hh[1]);
        }
    }

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
}

}  // namespace b200
