// Microbenchmark: does a co-running warp on the same SM sub-partition slow a warp's FP64 chain work?
// 8 warps per CTA, one CTA per SM: warps 0-3 (one per sub-partition) run NCH independent DFMA chains;
// warps 4-7 run a co-runner selected by `mode`: 0 idle, 1 ALU-pipe ints (LOP3/IADD3), 2 IMAD (fma pipe),
// 3 LDS.128 stream, 4 another DFMA stream.  Reports cycles of the FP64 warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int NCH>
__global__ void __launch_bounds__(256, 1) k(long long *out, int iters, int mode, double seed) {
    __shared__ double4 sm[1024];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_double4(i, 1, 2, 3);
    __syncthreads();
    long long t0 = clock64();
    if (warp < 4) {
        double x[NCH];
#pragma unroll
        for (int c = 0; c < NCH; c++) x[c] = seed + c;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < NCH; c++) x[c] = fma(x[c], 1.0000001, 0.5);
        }
        double s = 0;
#pragma unroll
        for (int c = 0; c < NCH; c++) s += x[c];
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + warp] = t1 - t0;
        if (s == 1.2345) out[0] = 0;
    } else if (mode == 1) {
        uint32_t a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = threadIdx.x + c;
        for (int it = 0; it < iters * NCH; it++) {
#pragma unroll
            for (int c = 0; c < 8; c++) a[c] = (a[c] ^ (a[c] >> 3)) + 0x9E3779B9u;
        }
        uint32_t s = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) s += a[c];
        if (s == 0x1234567u) out[1] = 0;
    } else if (mode == 2) {
        uint32_t a[8];
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = threadIdx.x + c;
        for (int it = 0; it < iters * NCH; it++) {
#pragma unroll
            for (int c = 0; c < 8; c++) a[c] = a[c] * 0x9E3779B1u + 12345u;
        }
        uint32_t s = 0;
#pragma unroll
        for (int c = 0; c < 8; c++) s += a[c];
        if (s == 0x1234567u) out[1] = 0;
    } else if (mode == 3) {
        double acc = 0;
        int idx = threadIdx.x & 31;
        for (int it = 0; it < iters * NCH / 2; it++) {
#pragma unroll
            for (int c = 0; c < 8; c++) { double4 v = sm[(idx + 32 * c) & 1023]; acc += v.x; idx = (idx + 7) & 1023; }
        }
        if (acc == 1.2345) out[1] = 0;
    } else if (mode == 4) {
        double x[NCH];
#pragma unroll
        for (int c = 0; c < NCH; c++) x[c] = seed + c;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int c = 0; c < NCH; c++) x[c] = fma(x[c], 1.0000001, 0.5);
        }
        double s = 0;
#pragma unroll
        for (int c = 0; c < NCH; c++) s += x[c];
        if (s == 1.2345) out[1] = 0;
    }
}
template <int NCH>
void run(long long *d, int iters) {
    for (int mode = 0; mode <= 4; mode++) {
        k<NCH><<<148, 256>>>(d, iters, mode, 0.7);
        cudaDeviceSynchronize();
        k<NCH><<<148, 256>>>(d, iters, mode, 0.7);
        cudaDeviceSynchronize();
        long long h[4];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double per = (double)h[0] / ((double)iters * 8 * NCH);
        printf("{\"chains\": %d, \"corunner\": %d, \"cycles_per_dfma\": %.3f}\n", NCH, mode, per);
    }
}
int main() {
    long long *d; cudaMalloc(&d, 148 * 4 * sizeof(long long));
    run<1>(d, 2000); run<2>(d, 2000); run<4>(d, 2000); run<8>(d, 2000);
    return 0;
}
