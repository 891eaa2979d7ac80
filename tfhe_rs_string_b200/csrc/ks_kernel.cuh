// ks_kernel.cuh -- batched LWE keyswitch as a register-tiled u64 integer GEMM.
//
// Replaces keyswitch_lwe_ciphertext (core_crypto/algorithms/lwe_keyswitch.rs:96-170) and its inner
// slice_wrapping_sub_scalar_mul_assign (slice_algorithms.rs:363-462) for a whole batch:
//     out[b][:] = (0, .., 0, in[b].body) - sum_{i, lvl} digit(in[b][i], lvl) * KSK[i][lvl][:]
// i.e. Out[B x out_size] = Body - D[B x (n_in*L)] * KSK[(n_in*L) x out_size] in wrapping u64.
// Bit-exact: integer arithmetic mod 2^64 is associative, so the tiling cannot change a bit
// (the reference asserts the same of its own threaded variant, algorithms/test/lwe_keyswitch.rs:93).
//
// Digits are balanced, in [-B/2, B/2] (B = 2^base_log); we add B/2 so they are unsigned and the
// MAC is IMAD.WIDE.U32 + IMAD, and subtract (B/2) * column_sum(KSK) (precomputed at key load) at
// the end.  CTA tile: 64 ciphertexts x 128 output columns; thread tile: 8 x 4; KSK rows are
// staged with 8-byte cp.async (rows are 743 u64 = 5944 B, only 8 B aligned) and double buffered.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

constexpr int kKsBM = 64;    // ciphertexts per CTA
constexpr int kKsBN = 128;   // output columns per CTA
constexpr int kKsIC = 8;     // input mask elements per pipeline stage
constexpr int kKsThreads = 256;

struct KsArgs {
    const uint64_t *in;       // [batch][n_in + 1]
    const uint64_t *ksk;      // [n_in][level][out_size]
    const uint64_t *colsum;   // [out_size]  (B/2) * sum over all rows of KSK
    uint64_t *out;            // [batch][out_size]
    int batch, n_in, out_size, base_log, level;
};

__host__ __device__ inline size_t ks_smem_bytes(int level) {
    const int kc = kKsIC * level;
    return (size_t)2 * kc * kKsBN * sizeof(uint64_t) + (size_t)2 * kc * kKsBM;
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int UNROLL, int MINB>
__global__ void __launch_bounds__(kKsThreads, MINB) ks_kernel(const KsArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int kc = kKsIC * a.level;
    uint64_t *ktile = reinterpret_cast<uint64_t *>(smem);                           // [2][kc][BN]
    uint8_t *dtile = smem + (size_t)2 * kc * kKsBN * sizeof(uint64_t);               // [2][kc][BM]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int col0 = blockIdx.x * kKsBN, ct0 = blockIdx.y * kKsBM;
    const int n_chunks = (a.n_in + kKsIC - 1) / kKsIC;
    const uint64_t mod_b_mask = ((uint64_t)1 << a.base_log) - 1;
    const uint32_t half_b = 1u << (a.base_log - 1);
    const int rep_bits = a.base_log * a.level;

    // acc = alo + (ahi << 32): low words accumulate in 64 bits (32x32+64 IMAD.WIDE, < 2^49 over the
    // whole sum), high words only matter mod 2^32 (one 32-bit IMAD); no carry chain per MAC.
    uint64_t alo[8][4];
    uint32_t ahi[8][4];
#pragma unroll
    for (int c = 0; c < 8; c++)
#pragma unroll
        for (int j = 0; j < 4; j++) { alo[c][j] = 0; ahi[c][j] = 0; }

    auto stage_load = [&](int chunk, int buf) {
        // KSK rows [chunk*IC*L, +kc) x columns [col0, +BN)
        uint64_t *kt = ktile + (size_t)buf * kc * kKsBN;
        const size_t row0 = (size_t)chunk * kc;
        const size_t n_rows_total = (size_t)a.n_in * a.level;
        for (int e = tid; e < kc * kKsBN; e += kKsThreads) {
            const int r = e / kKsBN, c = e % kKsBN;
            size_t row = row0 + r;
            if (row >= n_rows_total) row = n_rows_total - 1;   // padded rows are multiplied by a zero byte
            int col = col0 + c;
            if (col >= a.out_size) col = a.out_size - 1;
            cp_async8(kt + e, a.ksk + row * a.out_size + col);
        }
        cp_async_commit();
        // digits: SignedDecomposer::decompose (decomposer.rs:98-152, iter.rs:120-127), level L first
        uint8_t *dt = dtile + (size_t)buf * kc * kKsBM;
        for (int e = tid; e < kKsBM * kKsIC; e += kKsThreads) {
            const int ii = e % kKsIC, cl = e / kKsIC;
            const int i = chunk * kKsIC + ii;
            int ct = ct0 + cl;
            if (ct >= a.batch) ct = a.batch - 1;
            if (i < a.n_in) {
                const uint64_t x = a.in[(size_t)ct * (a.n_in + 1) + i];
                const int shift = 64 - rep_bits - 1;
                uint64_t res = ((x >> shift) + 1) & ~(uint64_t)1;
                uint64_t state = (res << shift) >> (64 - rep_bits);
                for (int li = 0; li < a.level; li++) {
                    uint64_t d = state & mod_b_mask;
                    state >>= a.base_log;
                    uint64_t carry = ((d - 1) | state) & d;
                    carry >>= (a.base_log - 1);
                    state += carry;
                    d -= carry << a.base_log;   // signed digit in [-B/2, B/2]
                    dt[(ii * a.level + li) * kKsBM + cl] = (uint8_t)((uint32_t)d + half_b);
                }
            } else {
                for (int li = 0; li < a.level; li++) dt[(ii * a.level + li) * kKsBM + cl] = 0;
            }
        }
    };

    stage_load(0, 0);
    for (int chunk = 0; chunk < n_chunks; chunk++) {
        const int buf = chunk & 1;
        if (chunk + 1 < n_chunks) {
            stage_load(chunk + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint64_t *kt = ktile + (size_t)buf * kc * kKsBN + tx * 4;
        const uint8_t *dt = dtile + (size_t)buf * kc * kKsBM + ty * 8;
#pragma unroll UNROLL
        for (int k = 0; k < kc; k++) {
            const ulonglong2 k01 = *reinterpret_cast<const ulonglong2 *>(kt + (size_t)k * kKsBN);
            const ulonglong2 k23 = *reinterpret_cast<const ulonglong2 *>(kt + (size_t)k * kKsBN + 2);
            const uint2 dd = *reinterpret_cast<const uint2 *>(dt + (size_t)k * kKsBM);
            const uint32_t klo[4] = {(uint32_t)k01.x, (uint32_t)k01.y, (uint32_t)k23.x, (uint32_t)k23.y};
            const uint32_t khi[4] = {(uint32_t)(k01.x >> 32), (uint32_t)(k01.y >> 32), (uint32_t)(k23.x >> 32),
                                     (uint32_t)(k23.y >> 32)};
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const uint32_t w = (c < 4) ? dd.x : dd.y;
                const uint32_t d = (w >> (8 * (c & 3))) & 0xFFu;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    alo[c][j] += (uint64_t)klo[j] * d;
                    ahi[c][j] += khi[j] * d;
                }
            }
        }
        __syncthreads();
    }

#pragma unroll
    for (int c = 0; c < 8; c++) {
        const int ct = ct0 + ty * 8 + c;
        if (ct >= a.batch) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int col = col0 + tx * 4 + j;
            if (col >= a.out_size) continue;
            // acc = sum (digit + B/2) * KSK  =>  -sum digit * KSK = colsum - acc
            uint64_t v = a.colsum[col] - (alo[c][j] + ((uint64_t)ahi[c][j] << 32));
            if (col == a.out_size - 1) v += a.in[(size_t)ct * (a.n_in + 1) + a.n_in];
            a.out[(size_t)ct * a.out_size + col] = v;
        }
    }
}

// colsum[col] = (B/2) * sum_{rows} KSK[row][col]; colsum must be zeroed before the launch.
__global__ void ks_colsum_kernel(const uint64_t *__restrict__ ksk, uint64_t *__restrict__ colsum,
                                 const size_t n_rows, const int out_size, const uint64_t half_b) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= out_size) return;
    const size_t rows_per = (n_rows + gridDim.y - 1) / gridDim.y;
    const size_t r0 = (size_t)blockIdx.y * rows_per;
    const size_t r1 = r0 + rows_per < n_rows ? r0 + rows_per : n_rows;
    uint64_t s = 0;
    for (size_t r = r0; r < r1; r++) s += ksk[r * out_size + col];
    atomicAdd(reinterpret_cast<unsigned long long *>(colsum + col), (unsigned long long)(s * half_b));
}

// ---------------------------------------------------------------------------------------------
// LWE linear algebra on device-resident batches (core_crypto/algorithms/lwe_linear_algebra.rs:
// add_assign :68, plaintext_add_assign :276, cleartext_mul_assign :556, sub_assign :703) fused as
//     out[b][:] = ca[b] * x[ia[b]][:] + cb[b] * y[ib[b]][:] ; body += pt[b]
// which covers add/sub/scalar-mul/plaintext-add and the bivariate pack lhs*factor + rhs
// (shortint/server_key/bivariate_pbs.rs:173-181) with one launch per dependency level.
struct LinArgs {
    const uint64_t *x;       // [*][size]
    const uint64_t *y;       // [*][size] (may be nullptr when cb == 0 everywhere)
    const int32_t *ia, *ib;  // [batch] row indices into x / y (nullptr = identity)
    const int64_t *ca, *cb;  // [batch] small signed scalars
    const uint64_t *pt;      // [batch] plaintext added to the body (nullptr = none)
    uint64_t *out;           // [batch][size]
    int batch, size;
};

__global__ void lwe_linear_kernel(const LinArgs a) {
    const int b = blockIdx.y;
    if (b >= a.batch) return;
    const uint64_t ca = (uint64_t)a.ca[b], cb = a.cb ? (uint64_t)a.cb[b] : 0;
    const uint64_t *xr = a.x + (size_t)(a.ia ? a.ia[b] : b) * a.size;
    const uint64_t *yr = a.y ? a.y + (size_t)(a.ib ? a.ib[b] : b) * a.size : nullptr;
    uint64_t *o = a.out + (size_t)b * a.size;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < a.size; j += gridDim.x * blockDim.x) {
        uint64_t v = ca * xr[j];
        if (yr && cb) v += cb * yr[j];
        if (a.pt && j == a.size - 1) v += a.pt[b];
        o[j] = v;
    }
}


// ---------------------------------------------------------------------------------------------
// Stage kernels of the level executor: out[k] = sum_t coeff[t] * pool[block[t]] (+ plaintext on the
// body) for every node k of one dependency level (generalises lwe_linear_kernel to CSR term lists),
// and the final gather of the program outputs.
__global__ void lwe_lincomb_kernel(const uint64_t *__restrict__ pool, const int32_t *__restrict__ term_block,
                                   const int64_t *__restrict__ term_coeff, const uint32_t *__restrict__ node_tbeg,
                                   const uint64_t *__restrict__ node_pt, uint64_t *__restrict__ out,
                                   const int first_node, const int size) {
    const int k = first_node + blockIdx.y;
    const uint32_t t0 = node_tbeg[k], t1 = node_tbeg[k + 1];
    uint64_t *o = out + (size_t)blockIdx.y * size;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < size; j += gridDim.x * blockDim.x) {
        uint64_t v = (j == size - 1) ? node_pt[k] : 0;
        for (uint32_t t = t0; t < t1; t++) v += (uint64_t)term_coeff[t] * pool[(size_t)term_block[t] * size + j];
        o[j] = v;
    }
}
__global__ void lwe_gather_kernel(const uint64_t *__restrict__ pool, const int32_t *__restrict__ ids,
                                  uint64_t *__restrict__ out, const int size) {
    const uint64_t *src = pool + (size_t)ids[blockIdx.y] * size;
    uint64_t *o = out + (size_t)blockIdx.y * size;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < size; j += gridDim.x * blockDim.x) o[j] = src[j];
}

}  // namespace b200
