// pbs_kernel_lat4.cuh -- lowest-latency variant of the programmable bootstrap: ONE ciphertext per SM, FOUR warps per GLWE
// polynomial (same contract and reference citations as pbs_kernel5.cuh; used for batches of at most one ciphertext per SM).
//
// A CMUX step is a serial dependency chain.  pbs_lat_kernel<1, true> gives a polynomial two warps (16 points per thread) and
// needs 9.2k cycles per step of which 3k are FP64 issue: the rest is the latency of the chain itself.  Here a thread owns 8 folded
// points, so every link of the chain is half as long.  The accumulator (G = -acc, pbs_kernel5.cuh) and the inter-pass twiddles
// stay in REGISTERS for the whole bootstrap.  A 32-point pass is a cross-warp radix-4 stage, done as two radix-2 levels, plus an
// 8-point transform in registers; the lane <-> register transposition between the passes stays as in fft.cuh and the frequency
// layout is unchanged (thread (lane k1, residue r) holds F[k1 + 32 (4 kappa + r)]), so the Fourier BSK of bsk_to_fourier_kernel is
// used as is.  tools/proto_fft8x4.py is the numpy model of this data flow.
//
// Mapping: warp w = 4 e + 2 p + c holds the input quarter q = c + 2 e (folded points lane + 32 m, m in [8 q, 8 q + 8)) of
// polynomial p and produces the output residue r = 2 c + e.  The two warps of a tensor-memory quadrant (w and w + 4, which also
// share an SM sub-partition) hold quarters q and q + 2 of the SAME polynomial, so the level with butterfly distance 16 is a
// lane-private exchange through 32 TMEM columns and only the level with distance 8 (c <-> 1 - c) goes through shared memory:
//   forward:  level 1 (TMEM, e = 0 / 1):  a = x_q + x_(q+2)   /   b = (x_q - x_(q+2)) W32^(mm + 8 c)
//             level 2 (smem, c = 0 / 1):  s0 + s1              /   (s0 - s1) W32^(2 mm);   8-point transform -> outputs 4 kappa + 2 c + e
//   inverse:  8-point transform on the inputs 4 kappa + 2 c + e;  level 2: c = 1 publishes u conj(W32^(2 mm)), result own +- other;
//             level 1: e = 1 publishes s conj(W32^(mm + 8 c)), result own +- other  -> index mm + 8 (c + 2 e)
// (The first version of this kernel did the radix-4 stage in one go through shared memory, every thread reading the three foreign
// quarters: 832 shared-memory wavefronts per warp and step, 74 % of the SM's shared-memory pipe, 3.68 ms per bootstrap level.  With
// one level in tensor memory it is 576 wavefronts and 3.43 ms; refilling the BSK slice by one thread right after the CTA barrier
// that already follows the last read of the slice -- instead of a consumer counter whose atomic returns after ~250 cycles -- 3.24.)
#pragma once
#include "pbs_kernel5.cuh"

namespace b200 {

__host__ __device__ constexpr int brev3(int x) { return ((x & 1) << 2) | (x & 2) | ((x & 4) >> 2); }

// In-register 8-point DFT, decimation in time: input in bit-reversed register order, output natural.
template <bool INV>
__device__ __forceinline__ void fft8_dit(double (&xr)[8], double (&xi)[8]) {
#pragma unroll
    for (int half = 1; half < 8; half <<= 1) {
#pragma unroll
        for (int base = 0; base < 8; base += 2 * half) {
#pragma unroll
            for (int t = 0; t < half; t++)
                bfly<INV>(xr[base + t], xi[base + t], xr[base + t + half], xi[base + t + half], t * (16 / half));   // W_(2 half)^t = W32^(t*16/half)
        }
    }
}

constexpr int kLat4ExBytes = 4 * 8 * 32 * (int)sizeof(double2);   // one exchange buffer of a polynomial: [quarter][mm][lane]
constexpr int kLat4PolyBytes = kBuf5Bytes + 2 * kLat4ExBytes;     // rotation copy / transposition / transform + two exchange buffers
__host__ __device__ constexpr size_t pbs_lat4_smem_bytes() {
    return kPbsHeaderBytes + kBskSliceBytes + (size_t)2 * kLat4PolyBytes + kMaxSmallDim * sizeof(uint16_t);
}

__device__ __forceinline__ void poly_barrier(const int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void cta_barrier256(const int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }

__device__ __forceinline__ void pair_barrier64(const int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// swap 8 complex values with the warp that shares this warp's tensor-memory quadrant (lane-private)
__device__ __forceinline__ void lat4_tmem_swap(const double (&vr)[8], const double (&vi)[8], double (&rr)[8], double (&ri)[8],
                                                const uint32_t xout, const uint32_t xin, const int bar) {
#pragma unroll
    for (int k = 0; k < 2; k++) {
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
    pair_barrier64(bar);
    tmem_fence_after();
    uint32_t g0[16], g1[16];
    tmem_ld16_nc(xin, g0);
    tmem_ld16_nc(xin + 16, g1);
    tmem_wait_ld16(g0);
    tmem_wait_ld16(g1);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        rr[j] = dbl(g0[4 * j], g0[4 * j + 1]); ri[j] = dbl(g0[4 * j + 2], g0[4 * j + 3]);
        rr[4 + j] = dbl(g1[4 * j], g1[4 * j + 1]); ri[4 + j] = dbl(g1[4 * j + 2], g1[4 * j + 3]);
    }
}

// forward pass: x = own quarter (natural mm) -> y[kappa] = output 4 kappa + 2 c + e
__device__ __forceinline__ void lat4_fwd_pass(const double (&xr)[8], const double (&xi)[8], double (&yr)[8], double (&yi)[8],
                                               const int c, const int e, const uint32_t xout, const uint32_t xin, const int bar_pair,
                                               double2 *ex_own, const double2 *ex_oth, const int bar_poly) {
    double rr[8], ri[8], sr[8], si[8];
    lat4_tmem_swap(xr, xi, rr, ri, xout, xin, bar_pair);
    const double sg1 = e ? -1.0 : 1.0;   // e = 0: own + other;  e = 1: other - own = x_q - x_(q+2) (own is quarter q + 2)
#pragma unroll
    for (int mm = 0; mm < 8; mm++) {
        double tr = fma(sg1, xr[mm], rr[mm]), ti = fma(sg1, xi[mm], ri[mm]);
        if (e) {   // times W32^(mm + 8 c) (warp-uniform branch, table index uniform)
            const double wr = w32_re(mm + 8 * c), wi = w32_im(mm + 8 * c);
            const double nr = fma(-ti, wi, tr * wr);
            ti = fma(ti, wr, tr * wi);
            tr = nr;
        }
        sr[mm] = tr; si[mm] = ti;
        ex_own[mm * 32] = make_double2(tr, ti);
    }
    poly_barrier(bar_poly);
    const double sg2 = c ? -1.0 : 1.0;   // c = 0: own + other;  c = 1: other - own = s0 - s1
#pragma unroll
    for (int mm = 0; mm < 8; mm++) {
        const double2 v = ex_oth[mm * 32];
        double tr = fma(sg2, sr[mm], v.x), ti = fma(sg2, si[mm], v.y);
        if (c && mm) {   // times W32^(2 mm)
            if (mm == 4) { const double t = tr; tr = ti; ti = -t; }   // -i
            else {
                const double wr = w32_re(2 * mm), wi = w32_im(2 * mm);
                const double nr = fma(-ti, wi, tr * wr);
                ti = fma(ti, wr, tr * wi);
                tr = nr;
            }
        }
        yr[brev3(mm)] = tr; yi[brev3(mm)] = ti;
    }
    fft8_dit<false>(yr, yi);
}
// inverse pass: x[brev3(kappa)] = inputs 4 kappa + 2 c + e -> y[mm] = result for index mm + 8 (c + 2 e)
__device__ __forceinline__ void lat4_inv_pass(double (&xr)[8], double (&xi)[8], double (&yr)[8], double (&yi)[8],
                                               const int c, const int e, const uint32_t xout, const uint32_t xin, const int bar_pair,
                                               double2 *ex_own, const double2 *ex_oth, const int bar_poly) {
    fft8_dit<true>(xr, xi);
#pragma unroll
    for (int mm = 0; mm < 8; mm++) {
        if (c && mm) {   // times conj(W32^(2 mm))
            if (mm == 4) { const double t = xr[mm]; xr[mm] = -xi[mm]; xi[mm] = t; }   // +i
            else {
                const double wr = w32_re(2 * mm), wi = -w32_im(2 * mm);
                const double nr = fma(-xi[mm], wi, xr[mm] * wr);
                xi[mm] = fma(xi[mm], wr, xr[mm] * wi);
                xr[mm] = nr;
            }
        }
        ex_own[mm * 32] = make_double2(xr[mm], xi[mm]);
    }
    poly_barrier(bar_poly);
    const double sg2 = c ? -1.0 : 1.0;   // c = 0: own + other';  c = 1: other - own'
    double sr[8], si[8];
#pragma unroll
    for (int mm = 0; mm < 8; mm++) {
        const double2 v = ex_oth[mm * 32];
        double tr = fma(sg2, xr[mm], v.x), ti = fma(sg2, xi[mm], v.y);
        if (e) {   // times conj(W32^(mm + 8 c)) before it goes to the sibling
            const double wr = w32_re(mm + 8 * c), wi = -w32_im(mm + 8 * c);
            const double nr = fma(-ti, wi, tr * wr);
            ti = fma(ti, wr, tr * wi);
            tr = nr;
        }
        sr[mm] = tr; si[mm] = ti;
    }
    double rr[8], ri[8];
    lat4_tmem_swap(sr, si, rr, ri, xout, xin, bar_pair);
    const double sg1 = e ? -1.0 : 1.0;   // e = 0: own + other';  e = 1: other - own'
#pragma unroll
    for (int mm = 0; mm < 8; mm++) {
        yr[mm] = fma(sg1, sr[mm], rr[mm]);
        yi[mm] = fma(sg1, si[mm], ri[mm]);
    }
}

__global__ void __launch_bounds__(256, 1) pbs_lat4_kernel(const PbsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = warp & 1, p = (warp >> 1) & 1, e = warp >> 2;
    const int q = c + 2 * e;        // input quarter: folded points lane + 32 m, m in [8 q, 8 q + 8)
    const int r = 2 * c + e;        // output residue: frequencies 4 kappa + r
    const int ct = blockIdx.x;
    if (ct >= a.batch) return;   // (CTA-uniform)

    uint32_t *slot = reinterpret_cast<uint32_t *>(smem);
    uint64_t *bsk_bar = reinterpret_cast<uint64_t *>(smem + 8);
    double2 *bsk_s = reinterpret_cast<double2 *>(smem + kPbsHeaderBytes);
    unsigned char *pbase = smem + kPbsHeaderBytes + kBskSliceBytes + (size_t)p * kLat4PolyBytes;
    unsigned char *obase = smem + kPbsHeaderBytes + kBskSliceBytes + (size_t)(1 - p) * kLat4PolyBytes;
    double2 *tb_own = reinterpret_cast<double2 *>(pbase);                      // shared by the four warps of the polynomial
    const double2 *tb_oth = reinterpret_cast<const double2 *>(obase);
    // level-2 exchange buffers [buffer][c][e][mm][lane]: this warp writes (c, e) and reads (1 - c, e)
    double2 *exb = reinterpret_cast<double2 *>(pbase + kBuf5Bytes);
    double2 *ex0_own = exb + ((c * 2 + e) * 8) * 32 + lane, *ex1_own = ex0_own + kLat4ExBytes / (int)sizeof(double2);
    const double2 *ex0_oth = exb + (((1 - c) * 2 + e) * 8) * 32 + lane, *ex1_oth = ex0_oth + kLat4ExBytes / (int)sizeof(double2);
    uint16_t *ahat = reinterpret_cast<uint16_t *>(smem + kPbsHeaderBytes + kBskSliceBytes + (size_t)2 * kLat4PolyBytes);
    uint64_t *rot = reinterpret_cast<uint64_t *>(tb_own);   // rotation copy (G = -acc, + overflow zone, pbs_kernel5.cuh) aliases the transposition buffer
    const int bar_pair = 1 + (warp & 3), bar_poly = 5 + p, bar_ct = 7;

    if (warp == 0) tmem_alloc(slot, 128);
    if (threadIdx.x == 0) mbar_init(bsk_bar, 1);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *slot;
    const uint32_t tquad = tbase + (((uint32_t)(warp & 3) * 32u) << 16);
    // four 32-column exchange areas per quadrant: [buffer][half e]; this warp writes (buffer, e) and reads (buffer, 1 - e)
    const uint32_t tx0_out = tquad + (uint32_t)e * 32u, tx0_in = tquad + (uint32_t)(1 - e) * 32u;
    const uint32_t tx1_out = tx0_out + 64u, tx1_in = tx0_in + 64u;
    if (threadIdx.x == 0) issue_bsk_slice(bsk_s, a.bsk, 0, bsk_bar);

    // this thread's inter-pass twiddles T'[4 kappa + r][lane]
    double twr[8], twi[8];
#pragma unroll
    for (int kap = 0; kap < 8; kap++) {
        const double2 t = __ldg(a.twid + (4 * kap + r) * 32 + lane);
        twr[kap] = t.x; twi[kap] = t.y;
    }

    // ---------------------------------------------------------------- prologue
    const uint64_t *lwe = a.lwe_small + (size_t)ct * (a.n + 1);
    for (int i = threadIdx.x; i < a.n; i += 256) ahat[i] = (uint16_t)modswitch2048(lwe[i]);
    const uint32_t bhat = modswitch2048(lwe[a.n]);
    const uint64_t *lut = a.luts + ((size_t)pbs_lut_id(a, ct) * 2 + p) * kN;
    uint64_t g0[8], g1[8];   // G = -acc of this thread's 8 slots (coefficients j and j + 1024), in registers for the whole bootstrap
#pragma unroll
    for (int mm = 0; mm < 8; mm++) {
        const int j = lane + 32 * (8 * q + mm);
        const uint32_t i0 = (uint32_t)(j + bhat) & 4095u, i1 = (i0 + 1024u) & 4095u;
        uint64_t v0 = lut[i0 & 2047u], v1 = lut[i1 & 2047u];
        if (i0 & 2048u) v0 = 0 - v0;
        if (i1 & 2048u) v1 = 0 - v1;
        g0[mm] = 0 - v0; g1[mm] = 0 - v1;
        rot[j] = g0[mm]; rot[j + kHalf] = g1[mm];
        if (j < kZone5) rot[kN + j] = v0;
    }
    cta_barrier256(bar_ct);   // a~ table and both rotation copies complete

    // ---------------------------------------------------------------- CMUX loop
    for (int i = 0; i < a.n; i++) {
        PBS3_TS(0);
        double xr[8], xi[8], yr[8], yi[8];
        // phase A (group-uniform gather of pbs_kernel5.cuh: slots 8 q .. 8 q + 7 are group q of the first half, 4 + q of the second)
        {
            const uint32_t q0 = (4096u - (uint32_t)ahat[i]) & 4095u;
            const uint32_t qa = (q0 + 256u * (uint32_t)q) & 4095u, qb = (qa + 1024u) & 4095u;
            const uint64_t *pa = rot + lane + (qa & 2047u), *pb = rot + lane + (qb & 2047u);
            const uint32_t ta = (qa >> 11) - 1u, tb = (qb >> 11) - 1u;
            const uint64_t ma = pack64(ta, ta), mb = pack64(tb, tb), ca = pack64(ta & 1u, 0x7FFFFF00u), cb = pack64(tb & 1u, 0x7FFFFF00u);
#pragma unroll
            for (int mm = 0; mm < 8; mm++) {
                const uint64_t e0 = g0[mm] + (pa[32 * mm] ^ ma) + ca;
                const uint64_t e1 = g1[mm] + (pb[32 * mm] ^ mb) + cb;
                const double fr = dbl((uint32_t)(e0 >> 41), 0x43300000u) - 4503599631564799.0;
                const double fi = dbl((uint32_t)(e1 >> 41), 0x43300000u) - 4503599631564799.0;
                const double2 cm = c_twm[8 * q + mm];
                xr[mm] = fma(-fi, cm.y, fr * cm.x);
                xi[mm] = fma(fi, cm.x, fr * cm.y);
            }
        }
        PBS3_TS(1);
        // ---- forward transform (the barriers inside the first pass also order all rotation reads before the buffer is reused)
        lat4_fwd_pass(xr, xi, yr, yi, c, e, tx0_out, tx0_in, bar_pair, ex0_own, ex0_oth, bar_poly);
#pragma unroll
        for (int kap = 0; kap < 8; kap++) {
            const double nr = fma(-yi[kap], twi[kap], yr[kap] * twr[kap]);
            const double ni = fma(yi[kap], twr[kap], yr[kap] * twi[kap]);
            tb_own[lane * kTStride + 4 * kap + r] = make_double2(nr, ni);
        }
        PBS3_TS(2);
        poly_barrier(bar_poly);
#pragma unroll
        for (int ll = 0; ll < 8; ll++) {
            const double2 v = tb_own[(ll + 8 * q) * kTStride + lane];
            xr[ll] = v.x; xi[ll] = v.y;
        }
        lat4_fwd_pass(xr, xi, yr, yi, c, e, tx1_out, tx1_in, bar_pair, ex1_own, ex1_oth, bar_poly);   // its poly barrier orders the transposition reads before the writes below
        PBS3_TS(3);

        // ---- exchange the transforms between the two polynomials, Out_p = B[p][p] F_p + B[1-p][p] F_{1-p}
#pragma unroll
        for (int kap = 0; kap < 8; kap++) tb_own[(4 * kap + r) * 32 + lane] = make_double2(yr[kap], yi[kap]);
        mbar_wait(bsk_bar, (uint32_t)(i & 1));
        double zr[8], zi[8];
        {
            const double2 *b_own = bsk_s + (size_t)(p * 2 + p) * kHalf + lane;
#pragma unroll
            for (int kap = 0; kap < 8; kap++) {
                const double2 bo = b_own[(4 * kap + r) * 32];
                zr[brev3(kap)] = fma(-bo.y, yi[kap], bo.x * yr[kap]);
                zi[brev3(kap)] = fma(bo.y, yr[kap], bo.x * yi[kap]);
            }
        }
        PBS3_TS(4);
        cta_barrier256(bar_ct);
        PBS3_TS(5);
        {
            const double2 *b_oth = bsk_s + (size_t)((1 - p) * 2 + p) * kHalf + lane;
#pragma unroll
            for (int kap = 0; kap < 8; kap++) {
                const int qf = 4 * kap + r;
                const double2 bx = b_oth[qf * 32], g = tb_oth[qf * 32 + lane];
                const double o_r = fma(bx.x, g.x, zr[brev3(kap)]), o_i = fma(bx.x, g.y, zi[brev3(kap)]);
                zr[brev3(kap)] = fma(-bx.y, g.y, o_r); zi[brev3(kap)] = fma(bx.y, g.x, o_i);
            }
        }
        PBS3_TS(6);
        cta_barrier256(bar_ct);   // the sibling polynomial has read this one's transform before the buffer is reused ...
        // ... and every warp of the CTA is done with the BSK slice: no consumer counter, one thread refills the buffer
        if (threadIdx.x == 0 && i + 1 < a.n) issue_bsk_slice(bsk_s, a.bsk, i + 1, bsk_bar);
        PBS3_TS(7);

        // ---- inverse transform
        lat4_inv_pass(zr, zi, yr, yi, c, e, tx0_out, tx0_in, bar_pair, ex0_own, ex0_oth, bar_poly);   // y[ll]: index l = ll + 8 q of lane k1
#pragma unroll
        for (int ll = 0; ll < 8; ll++) tb_own[lane * kTStride + ll + 8 * q] = make_double2(yr[ll], yi[ll]);
        poly_barrier(bar_poly);
#pragma unroll
        for (int kap = 0; kap < 8; kap++) {
            const double2 v = tb_own[(4 * kap + r) * kTStride + lane];
            zr[brev3(kap)] = fma(v.y, twi[kap], v.x * twr[kap]);          // times conj(T')
            zi[brev3(kap)] = fma(v.y, twr[kap], -(v.x * twi[kap]));
        }
        lat4_inv_pass(zr, zi, yr, yi, c, e, tx1_out, tx1_in, bar_pair, ex1_own, ex1_oth, bar_poly);   // its poly barrier orders the transposition reads before the rotation copy below
        PBS3_TS(8);

        // ---- phase D: untwist, from_torus, G -= delta (registers), refresh the rotation copy
#pragma unroll
        for (int mm = 0; mm < 8; mm++) {
            const int j = lane + 32 * (mm + 8 * q);
            const double2 cm = c_twm[8 * q + mm];
            const double ur = fma(yi[mm], cm.y, yr[mm] * cm.x);
            const double ui = fma(yi[mm], cm.x, -(yr[mm] * cm.y));
            g0[mm] = g0[mm] + kFtBias - from_torus_fp(ur);   // acc += delta <=> G -= delta
            g1[mm] = g1[mm] + kFtBias - from_torus_fp(ui);
            rot[j] = g0[mm]; rot[j + kHalf] = g1[mm];
            if (j < kZone5) rot[kN + j] = 0 - g0[mm];        // (warp-uniform: q == 0)
        }
        PBS3_TS(9);
        poly_barrier(bar_poly);   // rotation copy complete (all four quarters) before the next step's gather
        PBS3_TS(10);
    }

    // ---------------------------------------------------------------- sample extraction (acc = -G)
    uint64_t *o = a.out + (size_t)ct * (kN + 1);
    if (p == 0) {
#pragma unroll
        for (int mm = 0; mm < 8; mm++) {
            const int j = lane + 32 * (8 * q + mm);
            if (j == 0) o[0] = 0 - g0[mm]; else o[kN - j] = g0[mm];
            o[kHalf - j] = g1[mm];   // coefficient j + 1024 -> index N - (j + 1024), negated
        }
    } else if (q == 0 && lane == 0) {
        o[kN] = 0 - g0[0];
    }
    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 128);
}

}  // namespace b200
