// pbs_common.cuh -- pieces shared by the programmable-bootstrap kernels for k = 1, N = 2048, one decomposition
// level (pbs_kernel5.cuh: throughput; pbs_kernel_lat.cuh: latency): launch arguments, shared-memory geometry,
// mbarrier / bulk-copy helpers, TMEM twiddle source, the standard -> Fourier key conversion kernel and the FFT
// product unit-test kernel.  Reference path being replaced: FourierLweBootstrapKeyView::bootstrap
// (core_crypto/fft_impl/fft64/crypto/bootstrap.rs:333-364).
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
    uint32_t n_luts;            // registered tables: ids >= n_luts are rejected (table 0 is used, *err_flag set)
    uint32_t *err_flag;         // device word, or-ed with 1 when a lut id was out of range (nullptr = unchecked)
    long long *dbg;             // development: per-warp phase timestamps of CTA 0 (nullptr = off)
};

constexpr int kMaxSmallDim = 1024;       // capacity of the per-ciphertext a~ table (u16 entries)
constexpr int kPbsHeaderBytes = 128;     // tmem slot, mbarrier, consumer counter
constexpr int kBskSliceBytes = 4 * kHalf * (int)sizeof(double2);   // 65,536
constexpr uint32_t kTmemTwCols = 128;    // twiddle columns per quadrant

__host__ __device__ constexpr size_t pbs_ct_smem_bytes() {
    return (size_t)2 * kTBufElems * sizeof(double2) + kMaxSmallDim * sizeof(uint16_t);
}
// fast_pbs_modulus_switch, core_crypto/fft_impl/common.rs:26-43 (log2 N = 11): result in [0, 2N]
__device__ __forceinline__ uint32_t modswitch2048(uint64_t x) { return (uint32_t)(((x >> 51) + 1) >> 1); }

// from_torus, core_crypto/commons/math/torus/mod.rs:72-78 (round-half-even like the x86 SIMD path)
__device__ __forceinline__ uint64_t from_torus_dev(double x) {
    const double f = x - rint(x);
    return (uint64_t)__double2ll_rn(f * 18446744073709551616.0);
}

__device__ __forceinline__ uint64_t pack64(const uint32_t lo, const uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
constexpr uint64_t kFtBias = 0x4338000000000000ull;           // bit pattern of 1.5 * 2^52
// from_torus on the FP64 pipe only, three instructions.  |x| < 2^37.  Returns d with
// round_half_even(x * 2^64) mod 2^64 = d - kFtBias.
//   u1 = x * 2^64 + 1.5*2^102 has ulp 2^50: its low mantissa word holds T = round(x * 2^14) (two's complement);
//   c = (1.5*2^102 + 1.5*2^52) - u1 = 1.5*2^52 - T * 2^50 is exact (a multiple of 2^50 below 2^103);
//   u = x * 2^64 + c is ONE rounding of x * 2^64 - T * 2^50 + 1.5*2^52 (|x * 2^64 - T * 2^50| <= 2^49), i.e. it holds
//   round_half_even(x * 2^64) - T * 2^50 as a 52-bit two's complement mantissa.
// (Round 2 started with the four-instruction form t = x + 1.5*2^38, l = x - (t - 1.5*2^38), u = l * 2^64 + 1.5*2^52: the
// same T and the same u, one more dependent FP64 instruction per coefficient.)
__device__ __forceinline__ uint64_t from_torus_fp(const double x) {
    const double u1 = fma(x, 18446744073709551616.0, __longlong_as_double(0x4658000000000000ll));   // 1.5 * 2^102
    const double c = __longlong_as_double(0x4658000000000006ll) - u1;                               // 1.5 * 2^102 + 1.5 * 2^52
    const double u = fma(x, 18446744073709551616.0, c);
    return pack64((uint32_t)__double2loint(u), (uint32_t)__double2hiint(u) + ((uint32_t)__double2loint(u1) << 18));
}
// lut id of ciphertext ct; device-buffer callers cannot be validated on the host, so the range check is here
__device__ __forceinline__ uint32_t pbs_lut_id(const PbsArgs &a, const int ct) {
    const uint32_t id = a.lut_idx ? a.lut_idx[ct] : 0u;
    if (id < a.n_luts) return id;
    if (a.err_flag) atomicOr(a.err_flag, 1u);
    return 0u;
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

// ---- pieces shared by pbs_kernel5.cuh, pbs_kernel_lat.cuh and the regression reference tools/lab/pbs_kernel3.cuh ----
constexpr uint64_t kAccC = 0x7FFFFF0000000000ull;     // C = 2^63 - 2^40
constexpr uint32_t kTmemAcc0 = 128, kTmemXchg = 384;   // column offsets inside a quadrant

// ---- TMEM accessors without a "memory" clobber: ordering is carried by the register operands, so
// the compiler stays free to schedule shared-memory loads across them.
__device__ __forceinline__ void tmem_ld16_nc(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ void tmem_st16_nc(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :
                 : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}

__device__ __forceinline__ double dbl(const uint32_t lo, const uint32_t hi) { return __hiloint2double((int)hi, (int)lo); }
__device__ __forceinline__ void undbl(const double d, uint32_t &lo, uint32_t &hi) {
    lo = (uint32_t)__double2loint(d); hi = (uint32_t)__double2hiint(d);
}

// from_torus (torus/mod.rs:72-78; round-half-even like fft/x86.rs:864): fractional part centred at
// 0, times 2^64 by adding 64 to the exponent field (f is 0 or |f| >= 2^-1022; +-0 / subnormal
// inputs become < 2^-950 and convert to 0), rounded to i64.  (Measured: replacing the F2I by an
// all-FP64 split conversion does not shorten the phase.)
__device__ __forceinline__ uint64_t from_torus_exp(const double x) {
    const double f = x - rint(x);
    const double s = __hiloint2double(__double2hiint(f) + (64 << 20), __double2loint(f));
#ifdef B200TFHE_LAB_NOSAT
    // development (tools/lab): a fractional part of exactly +1/2 gives 2^63 instead of the saturated 2^63 - 1, so that
    // pbs_kernel5 (whose from_torus does not saturate) can be compared bit for bit
    if (s == 9223372036854775808.0) return 0x8000000000000000ull;
#endif
    return (uint64_t)__double2ll_rn(s);
}

// Development aid (tools/timeline.py): compile with -DB200TFHE_TIMELINE to record per-warp phase
// timestamps of CTA 0, CMUX steps 100..107, into PbsArgs::dbg.
#ifdef B200TFHE_TIMELINE
#define PBS3_TS(k) do { if (a.dbg && blockIdx.x == 0 && lane == 0 && i >= 100 && i < 108) a.dbg[((i - 100) * 8 + warp) * 16 + (k)] = clock64(); } while (0)
#else
#define PBS3_TS(k) do { } while (0)
#endif
// Development aid (tools/lab): compile with -DB200TFHE_DUMP to record, for CTA 0 and the last CMUX step, every lane's digits,
// inverse-transform outputs and torus increments into PbsArgs::dbg ([warp][m][lane][6]).
#ifdef B200TFHE_DUMP
#define PBS_DUMP(k, v) do { if (a.dbg && blockIdx.x == 0 && i == a.n - 1) a.dbg[(((size_t)warp * 32 + m) * 32 + lane) * 6 + (k)] = (long long)(v); } while (0)
#else
#define PBS_DUMP(k, v) do { } while (0)
#endif
// named-barrier hand-over (bar.arrive / bar.sync over n threads)
__device__ __forceinline__ void bar_arrive(const int id, const int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_sync_n(const int id, const int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }


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

// Debug/unit-test kernel: both from_torus routines on caller data
__global__ void from_torus_test_kernel(const double *__restrict__ x, uint64_t *__restrict__ out_fp, uint64_t *__restrict__ out_cvt, const size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out_fp[i] = from_torus_fp(x[i]) - kFtBias;
    out_cvt[i] = from_torus_dev(x[i]);
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
        po[j] += from_torus_fp(yr) - kFtBias;           // the production conversion (pbs_kernel5.cuh, pbs_kernel_lat.cuh)
        po[j + kHalf] += from_torus_fp(yi) - kFtBias;
    }
}

}  // namespace b200
