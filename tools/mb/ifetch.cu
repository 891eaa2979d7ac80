// ifetch.cu -- development microbenchmark: instruction-fetch bandwidth of an SM for a loop body that does not fit
// the instruction caches.  Body = two regions X, Y of NI instructions each (alternating LOP3 / IMAD on 16 chains,
// so neither pipe nor dependencies bind: peak IPC 1 per warp).  Modes: all warps run X,Y in lock step; or the
// warps of the upper half of the CTA run Y,X (two fetch streams per sub-partition).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NI>
__device__ __forceinline__ void region(uint32_t (&u)[16], const uint32_t k1, const uint32_t k2) {
#pragma unroll
    for (int i = 0; i < NI / 2; i++) {
        const int a = (2 * i) & 15;     // 16 independent chains, dependency distance 16 instructions
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[a]) : "r"(k1), "r"(k2));
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(u[a + 1]) : "r"(k1), "r"(k2));
    }
}
template <int NI>
__global__ void __launch_bounds__(512, 1) k(long long *out, int iters, int swap_upper, uint32_t seed, uint32_t *sink) {
    uint32_t u[16];
#pragma unroll
    for (int c = 0; c < 16; c++) u[c] = threadIdx.x * 7 + c + seed;
    const int warp = threadIdx.x >> 5;
    const bool upper = swap_upper && (warp >= (int)(blockDim.x >> 6));
    __syncthreads();
    long long t0 = clock64();
    if (upper) region<NI>(u, seed, seed * 3 + 1);    // upper warps are half a loop ahead: they run Y while the others run X
    for (int it = 0; it < iters; it++) {
        region<NI>(u, seed, seed * 3 + 1);           // X   (two distinct copies: the compiler cannot merge asm volatile bodies)
        asm volatile("" ::: "memory");
        region<NI>(u, seed, seed * 3 + 1);           // Y
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < 16; c++) s += u[c];
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
    if (s == 0x1234567u) sink[0] = s;
}
static long long *d_out; static uint32_t *d_sink;
template <int NI>
void run(int threads, int swap) {
    const int iters = 200;
    for (int rep = 0; rep < 2; rep++) { k<NI><<<148, threads>>>(d_out, iters, swap, 3, d_sink); cudaDeviceSynchronize(); }
    long long h[16]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    const double instr = (double)iters * 2 * NI + (swap ? 0 : 0);
    printf("{\"body_kb\": %d, \"warps_per_sm\": %d, \"two_streams\": %d, \"ipc_per_warp\": %.3f, \"ipc_per_smsp\": %.3f, \"fetch_B_per_clk_per_warp\": %.2f}\n",
           2 * NI * 16 / 1024, threads / 32, swap, instr / h[0], instr / h[0] * (threads / 128.0 < 1 ? 1 : threads / 128.0), instr / h[0] * 16);
}
int main() {
    cudaMalloc(&d_out, 148 * 16 * sizeof(long long)); cudaMalloc(&d_sink, 64);
    run<2048>(32, 0); run<2048>(128, 0); run<2048>(256, 0); run<2048>(256, 1); run<2048>(512, 0); run<2048>(512, 1);
    return 0;
}
