// fft.cuh -- warp-level negacyclic FFT for N = 2048 (1024 complex points), FP64, sm_100a.
//
// One warp transforms one polynomial.  Lane l owns the 32 folded points j = l + 32*m
// (m = 0..31), i.e. coefficients j (real part) and j + 1024 (imaginary part): the "home layout".
// 1024 = 32 x 32 Cooley-Tukey: a 32-point transform entirely in registers (over m), one twiddle,
// one transposition through shared memory, and a second 32-point transform in registers.
//
// The negacyclic twist w_j = exp(i*pi*j/2048) (reference: Twisties::new,
// core_crypto/fft_impl/fft64/math/fft/mod.rs:58-69) is split as w_j = A_l * C_m:
//   C_m = exp(i*pi*m/64)    depends on the register index only -> compile-time constant
//   A_l = exp(i*pi*l/2048)  depends on the lane only -> folded into the inter-pass twiddle table
// so the twist costs no table traffic.  T'[k1][l] = exp(-2*pi*i*l*k1/1024) * A_l.
//
// Frequency k = k1 + 32*k2 ends up in lane k1, register k2: "Fourier layout" [q = k2][lane],
// which is also how the Fourier bootstrap key is stored, so its loads are 512 B coalesced.
// This ordering is private to this library (the reference's ordering is plan dependent,
// fft/mod.rs:161,335-350, which is why the C ABI ingests the standard-domain BSK).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "fft_consts.h"

namespace b200 {

constexpr int kN = 2048;         // polynomial size
constexpr int kHalf = 1024;      // complex points
constexpr int kTStride = 33;     // padded row stride (in double2) of the transposition buffer
constexpr int kTBufElems = 32 * kTStride;  // double2 elements per warp buffer (16,896 B)

__device__ __constant__ double2 c_twm[32] = {B200_TWIST_M_TABLE};

// W32^t = exp(-2 pi i t / 32) in constant memory.  (Measured, tools/mb/pipes.cu: a DFMA with three distinct vector-register
// operands takes 3 issue cycles on the FP64 pipe, one with a uniform-register / constant operand 2; a 7-value
// deduplicated table was tried to keep every twiddle in uniform registers -- it cut the 3-operand DFMAs from 752 to 448
// per CMUX step but cost 2.5x the spill traffic and no time, so the plain table stays.)
__device__ __constant__ double2 c_w32[16] = {B200_W32_TABLE};
__device__ __forceinline__ double w32_re(const int t) { return c_w32[t].x; }
__device__ __forceinline__ double w32_im(const int t) { return c_w32[t].y; }

__host__ __device__ constexpr int brev5(int x) {
    return ((x & 1) << 4) | ((x & 2) << 2) | (x & 4) | ((x & 8) >> 2) | ((x & 16) >> 4);
}

// radix-2 DIT butterfly (a, b) -> (a + w*b, a - w*b), w = W32^TW (conjugated when INV).
// 6 FP64 instructions for a generic twiddle: two chained FMAs per component for a + w*b, then
// a - w*b = 2a - (a + w*b) as one FMA.  4 instructions for w in {1, -i, +i}.
template <bool INV>
__device__ __forceinline__ void bfly(double &ar, double &ai, double &br, double &bi, const int tw) {
    if (tw == 0) {
        const double tr = ar - br, ti = ai - bi;
        ar += br; ai += bi; br = tr; bi = ti;
    } else if (tw == 8) {
        // forward: w = -i -> w*b = (bi, -br); inverse: w = +i -> w*b = (-bi, br)
        const double pr = INV ? -bi : bi, pi = INV ? br : -br;
        const double tr = ar - pr, ti = ai - pi;
        ar += pr; ai += pi; br = tr; bi = ti;
    } else {
        const double wr = w32_re(tw);
        const double wi = INV ? -w32_im(tw) : w32_im(tw);
        double sr = fma(wr, br, ar);
        sr = fma(-wi, bi, sr);
        double si = fma(wr, bi, ai);
        si = fma(wi, br, si);
        br = fma(2.0, ar, -sr);
        bi = fma(2.0, ai, -si);
        ar = sr; ai = si;
    }
}

// In-register 32-point DFT, decimation in time: input in bit-reversed register order, output in
// natural order.  Register indices are compile-time after unrolling, so the bit reversal is free.
template <bool INV>
__device__ __forceinline__ void fft32_dit(double (&xr)[32], double (&xi)[32]) {
#pragma unroll
    for (int half = 1; half < 32; half <<= 1) {
#pragma unroll
        for (int base = 0; base < 32; base += 2 * half) {
#pragma unroll
            for (int t = 0; t < half; t++) {
                bfly<INV>(xr[base + t], xi[base + t], xr[base + t + half], xi[base + t + half], t * (16 / half));
            }
        }
    }
}

// The forward first pass with the register part C_m of the twist folded in: X_k = sum_m d_m C_m W32^(m k) is a DFT shifted
// by -1/4 of a frequency bin, and the decimation-in-time recursion keeps the shift at every size, so the butterfly t of
// the stage with distance `half` uses exp(-2 pi i (t - 1/4) / (2 half)) (c_w32s[(half - 1) + t]) and the inputs are the
// untwisted points.  80 generic butterflies = 480 FP64 instructions against 124 (twist) + 388.
__device__ __constant__ double2 c_w32s[32] = {B200_W32_SHIFTED_TABLE};
__device__ __forceinline__ void fft32_dit_twisted(double (&xr)[32], double (&xi)[32]) {
#pragma unroll
    for (int half = 1; half < 32; half <<= 1) {
#pragma unroll
        for (int base = 0; base < 32; base += 2 * half) {
#pragma unroll
            for (int t = 0; t < half; t++) {
                const double wr = c_w32s[half - 1 + t].x, wi = c_w32s[half - 1 + t].y;
                double &ar = xr[base + t], &ai = xi[base + t], &br = xr[base + t + half], &bi = xi[base + t + half];
                double sr = fma(wr, br, ar);
                sr = fma(-wi, bi, sr);
                double si = fma(wr, bi, ai);
                si = fma(wi, br, si);
                br = fma(2.0, ar, -sr);
                bi = fma(2.0, ai, -si);
                ar = sr; ai = si;
            }
        }
    }
}

// Twiddle sources: T'[k1][lane] for k1 = 4*chunk .. 4*chunk+3, as 16 32-bit words
// (re.lo, re.hi, im.lo, im.hi per twiddle).  issue() may be asynchronous; wait() completes it.
struct GlobalTwiddles {   // plain global-memory table (key conversion / unit-test kernels)
    const double2 *twid;
    int lane;
    __device__ __forceinline__ void issue(const int chunk, uint32_t (&r)[16]) const {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const double2 t = __ldg(twid + (chunk * 4 + kk) * 32 + lane);
            r[4 * kk] = (uint32_t)__double2loint(t.x); r[4 * kk + 1] = (uint32_t)__double2hiint(t.x);
            r[4 * kk + 2] = (uint32_t)__double2loint(t.y); r[4 * kk + 3] = (uint32_t)__double2hiint(t.y);
        }
    }
    __device__ __forceinline__ void wait() const {}
};

// y = x * t  (CONJ: x * conj(t)) with t packed as 4 words at r[4*kk..]
template <bool CONJ>
__device__ __forceinline__ void cmul_tw(double &yr, double &yi, const double xr, const double xi,
                                        const uint32_t (&r)[16], const int kk) {
    const double tr = __hiloint2double((int)r[4 * kk + 1], (int)r[4 * kk]);
    const double ti = __hiloint2double((int)r[4 * kk + 3], (int)r[4 * kk + 2]);
    if (!CONJ) {
        yr = fma(-xi, ti, xr * tr);
        yi = fma(xi, tr, xr * ti);
    } else {
        yr = fma(xi, ti, xr * tr);
        yi = fma(xi, tr, -(xr * ti));
    }
}

// Forward transform.  In: x[brev5(m)] = twisted-by-C_m folded point l + 32*m (A_l NOT applied).
// Out: x[k2] = F[lane + 32*k2].  tbuf: this warp's private 32x33 double2 buffer.
template <class TW>
__device__ __forceinline__ void fwd1024(double (&xr)[32], double (&xi)[32], double2 *tbuf, const TW &tw,
                                        const int lane) {
    uint32_t t0[16], t1[16];
    fft32_dit<false>(xr, xi);
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
            tbuf[lane * kTStride + k1] = y;
        }
        tw.wait();
        if (c2 < 3) tw.issue(2 * c2 + 2, t0);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const int k1 = 8 * c2 + 4 + kk;
            double2 y;
            cmul_tw<false>(y.x, y.y, xr[k1], xi[k1], t1, kk);
            tbuf[lane * kTStride + k1] = y;
        }
    }
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 32; l++) {
        const double2 v = tbuf[l * kTStride + lane];
        xr[brev5(l)] = v.x; xi[brev5(l)] = v.y;
    }
    __syncwarp();
    fft32_dit<false>(xr, xi);
}

// Inverse transform (unnormalised; the 1/1024 is folded into the Fourier BSK).
// In: x[brev5(k2)] = G[lane + 32*k2].  Out: x[m] = conj(A_l)-untwisted point l + 32*m; the caller
// still has to multiply by conj(C_m).  inv1024_pass1 touches registers only (the transposition buffer may still be
// in use by the sibling warp); inv1024_rest needs the buffer.
__device__ __forceinline__ void inv1024_pass1(double (&xr)[32], double (&xi)[32]) { fft32_dit<true>(xr, xi); }
template <class TW>
__device__ __forceinline__ void inv1024_rest(double (&xr)[32], double (&xi)[32], double2 *tbuf, const TW &tw,
                                             const int lane) {
    uint32_t t0[16], t1[16];
    tw.issue(0, t0);
#pragma unroll
    for (int l = 0; l < 32; l++) tbuf[lane * kTStride + l] = make_double2(xr[l], xi[l]);
    __syncwarp();
#pragma unroll
    for (int c2 = 0; c2 < 4; c2++) {
        tw.wait();
        tw.issue(2 * c2 + 1, t1);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const int k1 = 8 * c2 + kk;
            const double2 v = tbuf[k1 * kTStride + lane];
            cmul_tw<true>(xr[brev5(k1)], xi[brev5(k1)], v.x, v.y, t0, kk);
        }
        tw.wait();
        if (c2 < 3) tw.issue(2 * c2 + 2, t0);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const int k1 = 8 * c2 + 4 + kk;
            const double2 v = tbuf[k1 * kTStride + lane];
            cmul_tw<true>(xr[brev5(k1)], xi[brev5(k1)], v.x, v.y, t1, kk);
        }
    }
    __syncwarp();
    fft32_dit<true>(xr, xi);
}
template <class TW>
__device__ __forceinline__ void inv1024(double (&xr)[32], double (&xi)[32], double2 *tbuf, const TW &tw,
                                        const int lane) {
    inv1024_pass1(xr, xi);
    inv1024_rest(xr, xi, tbuf, tw, lane);
}

// multiply register m by C_m (forward twist) / conj(C_m) (inverse untwist)
__device__ __forceinline__ void twist_m(double &r, double &i, const int m) {
    if (m == 0) return;
    const double cr = c_twm[m].x, ci = c_twm[m].y;
    const double nr = fma(-i, ci, r * cr);
    i = fma(i, cr, r * ci);
    r = nr;
}
__device__ __forceinline__ void untwist_m(double &r, double &i, const int m) {
    if (m == 0) return;
    const double cr = c_twm[m].x, ci = c_twm[m].y;
    const double nr = fma(i, ci, r * cr);
    i = fma(i, cr, -(r * ci));
    r = nr;
}

}  // namespace b200
