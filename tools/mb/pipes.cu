// pipes.cu -- development microbenchmark: issue/pipe model of one SM sub-partition on B200 (sm_100a).
// One CTA per SM, warps w and w+4 share a sub-partition.  Every test reports cycles per "unit" for the
// measured warp(s).  Integer streams use 16 independent chains, FP64 streams 8, so nothing is latency bound.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { DFMA_RCC, DFMA_RRR, DADD_RR, IADD, LOP, SHFT, IMAD, IMADW, ADD64, F2I, FRND, LDS128, NOPS };

template <int OP>
__device__ __forceinline__ void stream(int iters, double seed, uint64_t *sink, const double4 *sm) {
    double x[8], a[8], b[8];
    uint32_t u[16];
    uint64_t v[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { x[c] = seed + c; a[c] = 1.0 + 1e-9 * (seed + c); b[c] = 1e-3 * (seed - c); v[c] = (uint64_t)(seed * 1e6) + c; }
#pragma unroll
    for (int c = 0; c < 16; c++) u[c] = threadIdx.x * 3 + c;
    double acc = 0;
    int idx = threadIdx.x & 31;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (OP == DFMA_RCC) {
#pragma unroll
                for (int c = 0; c < 8; c++) x[c] = fma(x[c], 1.0000001, 0.5);
            } else if (OP == DFMA_RRR) {
#pragma unroll
                for (int c = 0; c < 8; c++) x[c] = fma(a[c], b[(c + r) & 7], x[c]);
            } else if (OP == DADD_RR) {
#pragma unroll
                for (int c = 0; c < 8; c++) x[c] = x[c] + a[(c + r) & 7];
            } else if (OP == IADD) {
#pragma unroll
                for (int c = 0; c < 16; c++) asm volatile("add.u32 %0, %0, 0x9E3779B9;" : "+r"(u[c]));
            } else if (OP == LOP) {
#pragma unroll
                for (int c = 0; c < 16; c++) asm volatile("xor.b32 %0, %0, 0x9E3779B9;" : "+r"(u[c]));
            } else if (OP == SHFT) {
#pragma unroll
                for (int c = 0; c < 16; c++) asm volatile("shf.r.wrap.b32 %0, %0, %0, 7;" : "+r"(u[c]));
            } else if (OP == IMAD) {
#pragma unroll
                for (int c = 0; c < 16; c++) asm volatile("mad.lo.u32 %0, %0, 0x9E3779B1, 12345;" : "+r"(u[c]));
            } else if (OP == IMADW) {   // 64-bit add as mad.wide.u32 (lo) + mad.lo (hi): fma pipe
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    uint32_t lo = (uint32_t)v[c], hi = (uint32_t)(v[c] >> 32);
                    uint64_t t;
                    asm volatile("mad.wide.u32 %0, %1, 1, %2;" : "=l"(t) : "r"(u[c]), "l"(v[c]));
                    uint32_t thi = (uint32_t)(t >> 32);
                    asm volatile("mad.lo.u32 %0, %1, 1, %0;" : "+r"(thi) : "r"(u[c + 8]));
                    v[c] = ((uint64_t)thi << 32) | (uint32_t)t;
                    (void)lo; (void)hi;
                }
            } else if (OP == ADD64) {
#pragma unroll
                for (int c = 0; c < 8; c++) v[c] += ((uint64_t)u[c + 8] << 32) | u[c];
            } else if (OP == F2I) {
#pragma unroll
                for (int c = 0; c < 8; c++) { long long t = __double2ll_rn(x[c]); x[c] = __hiloint2double((int)(t >> 32) & 0x000FFFFF | 0x40000000, (int)t); }
            } else if (OP == FRND) {
#pragma unroll
                for (int c = 0; c < 8; c++) x[c] = rint(x[c]) * 1.0000001 + 0.3;
            } else if (OP == LDS128) {
#pragma unroll
                for (int c = 0; c < 8; c++) { double4 q = sm[(idx + 32 * c) & 1023]; acc += q.x; }
                idx = (idx + 7) & 1023;
            }
        }
    }
    double s = acc;
    uint64_t t = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) { s += x[c] + a[c] + b[c]; t += v[c]; }
#pragma unroll
    for (int c = 0; c < 16; c++) t += u[c];
    if (s == 1.2345 || t == 0x1234567ull) sink[0] = t;
}

// warps 0-3 run OPA (measured), warps 4-7 run OPB for itersB iterations (also measured)
template <int OPA, int OPB>
__global__ void __launch_bounds__(256, 1) k2(long long *out, int itersA, int itersB, double seed, uint64_t *sink) {
    __shared__ double4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_double4(i, 1, 2, 3);
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    long long t0 = clock64();
    if (warp < 4) stream<OPA>(itersA, seed, sink, sm);
    else if (OPB != NOPS) stream<OPB>(itersB, seed, sink, sm);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 8 + warp] = t1 - t0;
}

// single warp per sub-partition interleaving OPA and OPB in one instruction stream (ratio 8 FP64 : KB*8 others)
template <int OPB, int KB>
__global__ void __launch_bounds__(128, 1) kmix(long long *out, int iters, double seed, uint64_t *sink) {
    double x[8], a[8], b[8];
    uint32_t u[16];
#pragma unroll
    for (int c = 0; c < 8; c++) { x[c] = seed + c; a[c] = 1.0 + 1e-9 * (seed + c); b[c] = 1e-3 * (seed - c); }
#pragma unroll
    for (int c = 0; c < 16; c++) u[c] = threadIdx.x * 3 + c;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                x[c] = fma(a[c], b[(c + r) & 7], x[c]);
#pragma unroll
                for (int j = 0; j < KB; j++) {
                    const int i = (c * KB + j) & 15;
                    if (OPB == IADD) asm volatile("add.u32 %0, %0, 0x9E3779B9;" : "+r"(u[i]));
                    if (OPB == LOP) asm volatile("xor.b32 %0, %0, 0x9E3779B9;" : "+r"(u[i]));
                    if (OPB == IMAD) asm volatile("mad.lo.u32 %0, %0, 0x9E3779B1, 12345;" : "+r"(u[i]));
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0; uint32_t t = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) s += x[c];
#pragma unroll
    for (int c = 0; c < 16; c++) t += u[c];
    if ((threadIdx.x & 31) == 0) out[blockIdx.x * 8 + (threadIdx.x >> 5)] = t1 - t0;
    if (s == 1.2345 || t == 0x1234567u) sink[0] = t;
}

static long long *d_out; static uint64_t *d_sink;
static const char *names[] = {"dfma_rcc", "dfma_rrr", "dadd_rr", "iadd", "lop", "shf", "imad", "imadw64", "add64", "f2i_s64_f64", "frnd_f64", "lds128", "none"};
static int units(int op) { return (op == DFMA_RCC || op == DFMA_RRR || op == DADD_RR || op == IMADW || op == ADD64 || op == F2I || op == FRND || op == LDS128) ? 8 : 16; }

template <int OPA, int OPB>
void corun(int itA, int itB) {
    for (int rep = 0; rep < 2; rep++) { k2<OPA, OPB><<<148, 256>>>(d_out, itA, itB, 0.7, d_sink); cudaDeviceSynchronize(); }
    long long h[8]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"test\": \"corun\", \"a\": \"%s\", \"b\": \"%s\", \"a_units\": %d, \"b_units\": %d, \"a_cycles\": %lld, \"b_cycles\": %lld, \"a_cyc_per_unit\": %.3f, \"b_cyc_per_unit\": %.3f}\n",
           names[OPA], names[OPB], itA * 4 * units(OPA), OPB == NOPS ? 0 : itB * 4 * units(OPB), h[0], h[4],
           (double)h[0] / (itA * 4.0 * units(OPA)), OPB == NOPS ? 0.0 : (double)h[4] / (itB * 4.0 * units(OPB)));
}
template <int OPB, int KB>
void mix(int it) {
    for (int rep = 0; rep < 2; rep++) { kmix<OPB, KB><<<148, 128>>>(d_out, it, 0.7, d_sink); cudaDeviceSynchronize(); }
    long long h[8]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"test\": \"mix_one_warp\", \"b\": \"%s\", \"b_per_dfma\": %d, \"cycles_per_dfma\": %.3f}\n", names[OPB], KB, (double)h[0] / (it * 32.0));
}
int main() {
    cudaMalloc(&d_out, 148 * 8 * sizeof(long long)); cudaMalloc(&d_sink, 64);
    const int N = 2000;
    // alone
    corun<DFMA_RCC, NOPS>(N, 0); corun<DFMA_RRR, NOPS>(N, 0); corun<DADD_RR, NOPS>(N, 0);
    corun<IADD, NOPS>(N, 0); corun<LOP, NOPS>(N, 0); corun<SHFT, NOPS>(N, 0); corun<IMAD, NOPS>(N, 0);
    corun<IMADW, NOPS>(N, 0); corun<ADD64, NOPS>(N, 0); corun<F2I, NOPS>(N, 0); corun<FRND, NOPS>(N, 0); corun<LDS128, NOPS>(N, 0);
    // two warps on one sub-partition; B sized to take about as long as A alone (FP64: 2 cyc/unit, 32 units per iter => 64 cyc/iter)
    corun<DFMA_RRR, DFMA_RRR>(N, N);
    corun<DFMA_RRR, IADD>(N, N);      // 64 int per iter vs 32 dfma
    corun<DFMA_RRR, LOP>(N, N);
    corun<DFMA_RRR, IMAD>(N, N);
    corun<DFMA_RRR, IMADW>(N, N);
    corun<DFMA_RRR, ADD64>(N, N);
    corun<DFMA_RRR, F2I>(N, N / 4);
    corun<DFMA_RRR, FRND>(N, N / 4);
    corun<DFMA_RRR, LDS128>(N, N);
    corun<IADD, IADD>(N, N); corun<IADD, IMAD>(N, N); corun<IADD, LOP>(N, N); corun<F2I, FRND>(N / 4, N / 4);
    // one warp, interleaved streams
    mix<IADD, 0>(N); mix<IADD, 1>(N); mix<IADD, 2>(N); mix<LOP, 1>(N); mix<LOP, 2>(N); mix<IMAD, 1>(N); mix<IMAD, 2>(N);
    return 0;
}
