// lwe_linear.cuh -- leveled LWE operations between bootstraps (HBM-bound element-wise u64 kernels):
// lwe_linear_kernel (a*x[ia] + b*y[ib] + plaintext), lwe_lincomb_kernel (one dependency level of a circuit:
// arbitrary small-integer linear combinations of earlier blocks) and lwe_gather_kernel (program outputs).
// Reference: core_crypto/algorithms/lwe_linear_algebra.rs:68,276,556,703; shortint/server_key/bivariate_pbs.rs:173-181.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {


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
