// tmem_bw.cu -- development microbenchmark: TMEM load / store throughput per SM sub-partition and per SM
// (tcgen05.ld / tcgen05.st 32x32b.x16, the shape the PBS kernels use), 1..8 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../tfhe_rs_string_b200/csrc/tmem.cuh"
using namespace b200;

template <int MODE>   // 0 = ld x16, 1 = st x16, 2 = ld x16 with a wait after every load (latency)
__global__ void __launch_bounds__(256, 1) k(long long *out, int iters, uint32_t *sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 512);
    tmem_fence_before(); __syncthreads(); tmem_fence_after();
    const uint32_t tq = slot + (((uint32_t)(warp & 3) * 32u) << 16) + (uint32_t)(warp >> 2) * 256u;
    uint32_t r[16];
#pragma unroll
    for (int c = 0; c < 16; c++) r[c] = lane + c;
    for (int c = 0; c < 16; c++) tmem_st16(tq + c * 16, r);
    tmem_wait_st();
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 16; c++) {
            if (MODE == 0) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(tq + c * 16)); }
            else if (MODE == 2) { tmem_ld16(tq + c * 16, r); tmem_wait_ld(); acc += r[0]; }
            else { asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(tq + c * 16), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])); }
        }
        if (MODE == 0) { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += r[3]; }
        if (MODE == 1) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    const long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
    if (acc == 0x12345u) sink[0] = acc;
    tmem_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}
int main() {
    long long *d; uint32_t *s; cudaMalloc(&d, 148 * 8 * 8); cudaMalloc(&s, 64);
    const int iters = 2000;
    const char *names[3] = {"ld_x16", "st_x16", "ld_x16_wait_each"};
    for (int mode = 0; mode < 3; mode++)
        for (int warps = 1; warps <= 8; warps *= 2) {
            for (int rep = 0; rep < 2; rep++) {
                if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, s); else if (mode == 1) k<1><<<148, warps * 32>>>(d, iters, s); else k<2><<<148, warps * 32>>>(d, iters, s);
                cudaDeviceSynchronize();
            }
            long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            const double cyc = (double)h[0] / (iters * 16.0);
            printf("{\"test\": \"tmem_%s\", \"warps_per_sm\": %d, \"cycles_per_instr_per_warp\": %.2f, \"bytes_per_clk_per_sm\": %.1f, \"err\": \"%s\"}\n", names[mode], warps, cyc, warps * 2048.0 / cyc, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
