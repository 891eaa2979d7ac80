// microbench.cu -- measures the chip-level rates the PBS kernel is bounded by (B200, sm_100a):
// FP64 FMA peak (the roofline denominator MEASURED_PEAKS.json does not carry), FP64+INT co-issue,
// shared-memory and TMEM load throughput, f64<->i64 conversion throughput, L2 streaming reads.
// Prints one JSON object per line.  Build: make -C tfhe_rs_string_b200/csrc microbench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

#include "tmem.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s: %s\"}\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 12345.678) out[0] = s;
}

// 8 DFMA chains + 8 integer (IMAD/LOP) chains: does the integer work hide under the FP64 pipe?
__global__ void __launch_bounds__(256) k_dfma_int(double *out, int iters, double a, double b, uint32_t m) {
    double x[8];
    uint32_t y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 1e-3 + i; y[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            x[i] = fma(x[i], a, b);
            y[i] = (y[i] ^ m) + (y[i] >> 3);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + y[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_lds128(double *out, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    double2 *s = reinterpret_cast<double2 *>(sm);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            double2 v = s[(idx + u * 256) & 4095];
            acc.x += v.x; acc.y += v.y;
        }
        idx = (idx + 33) & 4095;
    }
    if (acc.x == 1.5) out[0] = acc.y;
}

// same loads but summing with integer ops so the FP64 pipe is idle (pure LDS rate)
__global__ void __launch_bounds__(256) k_lds128_int(uint32_t *out, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint4 *s = reinterpret_cast<uint4 *>(sm);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = make_uint4(i, i, i, i);
    __syncthreads();
    uint32_t acc = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint4 v = s[(idx + u * 256) & 4095];
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
        idx = (idx + 33) & 4095;
    }
    if (acc == 0x12345u) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_tmem_ld(uint32_t *out, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) b200::tmem_alloc(&slot, 512);
    b200::tmem_fence_before();
    __syncthreads();
    b200::tmem_fence_after();
    const uint32_t taddr = slot + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 128u;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = threadIdx.x + i;
    for (int c = 0; c < 4; c++) b200::tmem_st32(taddr + c * 32, r);
    b200::tmem_wait_st();
    uint32_t acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            b200::tmem_ld32(taddr + c * 32, r);
            b200::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i++) acc ^= r[i];
        }
    }
    if (acc == 0x12345u) out[0] = acc;
    b200::tmem_fence_before();
    __syncthreads();
    if (warp == 0) b200::tmem_dealloc(slot, 512);
}

__global__ void __launch_bounds__(256) k_tmem_st(uint32_t *out, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) b200::tmem_alloc(&slot, 512);
    b200::tmem_fence_before();
    __syncthreads();
    b200::tmem_fence_after();
    const uint32_t taddr = slot + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 128u;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; i++) r[i] = threadIdx.x + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            r[c] += it;
            b200::tmem_st32(taddr + c * 32, r);
        }
        b200::tmem_wait_st();
    }
    b200::tmem_ld32(taddr, r);
    b200::tmem_wait_ld();
    if (r[0] == 0x12345u) out[0] = r[1];
    b200::tmem_fence_before();
    __syncthreads();
    if (warp == 0) b200::tmem_dealloc(slot, 512);
}

// from_torus as the PBS kernel does it: rint + F2I.S64.F64
__global__ void __launch_bounds__(256) k_from_torus(uint64_t *out, int iters, double seed) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = seed * (threadIdx.x + 1) + i * 0.37;
    uint64_t acc = 0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double f = x[i] - rint(x[i]);
            acc += (uint64_t)__double2ll_rn(f * 18446744073709551616.0);
            x[i] += 0.123456789;
        }
    }
    if (acc == 0x12345u) out[0] = acc;
}

// magic-number variant: 52-bit precision, no FRND / F2I
__global__ void __launch_bounds__(256) k_from_torus_magic(uint64_t *out, int iters, double seed) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = seed * (threadIdx.x + 1) + i * 0.37;
    uint64_t acc = 0;
    const double M = 6755399441055744.0;  // 1.5 * 2^52
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            double r = (x[i] + M) - M;
            double f = x[i] - r;
            double h = fma(f, 4503599627370496.0, M);
            acc += (uint64_t)(__double_as_longlong(h) - __double_as_longlong(M)) << 12;
            x[i] += 0.123456789;
        }
    }
    if (acc == 0x12345u) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_l2_read(const double2 *__restrict__ src, size_t n_elems, double *out, int reps) {
    double2 acc = make_double2(0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += stride) {
            double2 v = __ldg(src + i);
            acc.x += v.x; acc.y += v.y;
        }
    if (acc.x == 1.2345) out[0] = acc.y;
}

template <typename F>
float time_ms(F f, int reps = 3) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %d}\n", prop.name, sms, clk_khz / 1000);
    void *scratch; CK(cudaMalloc(&scratch, 1 << 20));

    for (int bps : {1, 2, 4}) {
        const int iters = 20000;
        float ms = time_ms([&] { k_dfma<<<sms * bps, 256>>>((double *)scratch, iters, 1.0000001, 1e-9); });
        double flops = 2.0 * 8 * iters * 256.0 * sms * bps;
        printf("{\"bench\": \"dfma\", \"warps_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", bps * 8, ms, flops / ms / 1e9);
    }
    {
        const int iters = 20000, bps = 4;
        float ms = time_ms([&] { k_dfma_int<<<sms * bps, 256>>>((double *)scratch, iters, 1.0000001, 1e-9, 0x5bd1e995u); });
        double flops = 2.0 * 8 * iters * 256.0 * sms * bps;
        printf("{\"bench\": \"dfma_plus_3int_per_dfma\", \"ms\": %.3f, \"tflops\": %.2f}\n", ms, flops / ms / 1e9);
    }
    {
        const int iters = 4000, bps = 2;
        CK(cudaFuncSetAttribute(k_lds128, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_lds128_int, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        float ms = time_ms([&] { k_lds128<<<sms * bps, 256, 65536>>>((double *)scratch, iters); });
        double bytes = 16.0 * 8 * iters * 256.0 * sms * bps;
        printf("{\"bench\": \"lds128_dadd\", \"ms\": %.3f, \"TBps\": %.2f, \"B_per_clk_per_sm_at_max_clock\": %.1f}\n", ms, bytes / ms / 1e9,
               bytes / ms / 1e3 / sms / (clk_khz * 1.0));
        ms = time_ms([&] { k_lds128_int<<<sms * bps, 256, 65536>>>((uint32_t *)scratch, iters); });
        printf("{\"bench\": \"lds128_int\", \"ms\": %.3f, \"TBps\": %.2f, \"B_per_clk_per_sm_at_max_clock\": %.1f}\n", ms, bytes / ms / 1e9,
               bytes / ms / 1e3 / sms / (clk_khz * 1.0));
    }
    {
        const int iters = 4000;
        float ms = time_ms([&] { k_tmem_ld<<<sms, 256>>>((uint32_t *)scratch, iters); });
        double bytes = 4.0 * 128 * iters * 256.0 * sms;
        printf("{\"bench\": \"tmem_ld_x32\", \"ms\": %.3f, \"TBps\": %.2f, \"B_per_clk_per_sm_at_max_clock\": %.1f}\n", ms, bytes / ms / 1e9,
               bytes / ms / 1e3 / sms / (clk_khz * 1.0));
        ms = time_ms([&] { k_tmem_st<<<sms, 256>>>((uint32_t *)scratch, iters); });
        printf("{\"bench\": \"tmem_st_x32\", \"ms\": %.3f, \"TBps\": %.2f, \"B_per_clk_per_sm_at_max_clock\": %.1f}\n", ms, bytes / ms / 1e9,
               bytes / ms / 1e3 / sms / (clk_khz * 1.0));
        CK(cudaGetLastError());
    }
    {
        const int iters = 4000, bps = 4;
        float ms = time_ms([&] { k_from_torus<<<sms * bps, 256>>>((uint64_t *)scratch, iters, 0.7310585); });
        double n = 4.0 * iters * 256.0 * sms * bps;
        printf("{\"bench\": \"from_torus_rint_f2i\", \"ms\": %.3f, \"Gconv_per_s\": %.1f, \"conv_per_clk_per_sm_at_max_clock\": %.2f}\n", ms, n / ms / 1e6,
               n / ms / 1e3 / sms / (clk_khz * 1.0));
        ms = time_ms([&] { k_from_torus_magic<<<sms * bps, 256>>>((uint64_t *)scratch, iters, 0.7310585); });
        printf("{\"bench\": \"from_torus_magic\", \"ms\": %.3f, \"Gconv_per_s\": %.1f, \"conv_per_clk_per_sm_at_max_clock\": %.2f}\n", ms, n / ms / 1e6,
               n / ms / 1e3 / sms / (clk_khz * 1.0));
    }
    {
        const size_t bytes = 48627712;  // one Fourier BSK
        double2 *src; CK(cudaMalloc(&src, bytes));
        CK(cudaMemset(src, 0, bytes));
        const int reps = 20;
        float ms = time_ms([&] { k_l2_read<<<sms * 8, 256>>>(src, bytes / 16, (double *)scratch, reps); });
        printf("{\"bench\": \"l2_read_bsk_sized_48MB\", \"ms\": %.3f, \"TBps\": %.2f}\n", ms, (double)bytes * reps / ms / 1e9);
        cudaFree(src);
    }
    CK(cudaDeviceSynchronize());
    cudaFree(scratch);
    return 0;
}
