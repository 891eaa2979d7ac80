// Microbenchmark: one warp per SM sub-partition issues 4 independent DFMA chains interleaved with K
// integer/FP32 ops of one kind per DFMA.  Reports cycles per DFMA: which op kinds share the FP64 issue path.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int KIND, int K>
__global__ void __launch_bounds__(128, 1) k(long long *out, int iters, double seed) {
    double x[4];
    uint32_t a[8];
    float f[8];
#pragma unroll
    for (int c = 0; c < 4; c++) x[c] = seed + c;
#pragma unroll
    for (int c = 0; c < 8; c++) { a[c] = threadIdx.x * 7 + c; f[c] = (float)c + seed; }
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int c = 0; c < 4; c++) {
                x[c] = fma(x[c], 1.0000001, 0.5);
#pragma unroll
                for (int j = 0; j < K; j++) {
                    const int i = (c * K + j + r) & 7;
                    if (KIND == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(a[(i + 3) & 7]));
                    if (KIND == 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
                    if (KIND == 2) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]));
                    if (KIND == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(a[(i + 3) & 7]));
                    if (KIND == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 3) & 7]));
                    if (KIND == 5) asm volatile("{.reg .pred p; setp.gt.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(a[(i + 3) & 7]));
                    if (KIND == 6) asm volatile("mov.b32 %0, %1;" : "=r"(a[i]) : "r"(a[(i + 1) & 7]));
                    if (KIND == 7) asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(a[(i + 3) & 7]));
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0; uint32_t u = 0; float g = 0;
#pragma unroll
    for (int c = 0; c < 4; c++) s += x[c];
#pragma unroll
    for (int c = 0; c < 8; c++) { u += a[c]; g += f[c]; }
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 4 + (threadIdx.x >> 5)] = t1 - t0;
    if (s == 1.2345 || u == 0x12345u || g == 1.2345f) out[0] = 0;
}
template <int KIND, int K>
void run(long long *d, const char *name) {
    const int iters = 2000;
    k<KIND, K><<<148, 128>>>(d, iters, 0.7); cudaDeviceSynchronize();
    k<KIND, K><<<148, 128>>>(d, iters, 0.7); cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"op\": \"%s\", \"per_dfma\": %d, \"cycles_per_dfma\": %.3f}\n", name, K, (double)h[1] / (iters * 32.0));
}
int main() {
    long long *d; cudaMalloc(&d, 148 * 4 * sizeof(long long));
    run<0, 0>(d, "none");
    run<0, 1>(d, "lop3"); run<0, 2>(d, "lop3");
    run<1, 1>(d, "shf"); run<1, 2>(d, "shf");
    run<2, 1>(d, "add"); run<2, 2>(d, "add");
    run<3, 1>(d, "imad"); run<3, 2>(d, "imad");
    run<4, 1>(d, "ffma"); run<4, 2>(d, "ffma");
    run<5, 1>(d, "setp+selp"); run<5, 2>(d, "setp+selp");
    run<6, 1>(d, "mov"); run<6, 2>(d, "mov");
    run<7, 1>(d, "mad_by_1"); run<7, 2>(d, "mad_by_1");
    return 0;
}
