// ifetch2.cu -- development microbenchmark: instruction delivery of one SM sub-partition as a function of the loop body
// size.  Body = NI independent FFMAs on 16 chains (1 issue slot, 1 pipe cycle each: peak IPC 1 per sub-partition),
// executed in a loop by 1, 2 or 4 warps per sub-partition (all warps the same code, started together).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NI>
__global__ void __launch_bounds__(512, 1) k(long long *out, int iters, float seed, float *sink) {
    float u[16];
#pragma unroll
    for (int c = 0; c < 16; c++) u[c] = threadIdx.x * 0.5f + c + seed;
    const float k1 = seed * 0.999f, k2 = seed * 0.001f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NI; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(u[i & 15]) : "f"(k1), "f"(k2));
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int c = 0; c < 16; c++) s += u[c];
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + (threadIdx.x >> 5)] = t1 - t0;
    if (s == 1234.5f) sink[0] = s;
}
static long long *d_out; static float *d_sink;
template <int NI>
void run() {
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int iters = 4000000 / NI + 4;
        for (int rep = 0; rep < 2; rep++) { k<NI><<<148, warps * 32>>>(d_out, iters, 1.0f, d_sink); cudaDeviceSynchronize(); }
        long long h[16]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
        const double ipc = (double)iters * NI / h[0];
        printf("{\"test\": \"ifetch_ffma\", \"body_kb\": %.1f, \"warps_per_subpartition\": %d, \"ipc_per_warp\": %.3f, \"ipc_per_subpartition\": %.3f}\n", NI * 16 / 1024.0, warps / 4, ipc, ipc * warps / 4);
    }
}
int main() {
    cudaMalloc(&d_out, 148 * 16 * sizeof(long long)); cudaMalloc(&d_sink, 64);
    run<32>(); run<64>(); run<128>(); run<256>(); run<512>(); run<768>(); run<1024>(); run<1536>(); run<2048>(); run<3072>(); run<4096>(); run<6144>(); run<8192>();
    return 0;
}
