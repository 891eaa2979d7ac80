// ks_mma.cuh -- batched LWE keyswitch on the 5th-generation tensor cores (tcgen05.mma kind::i8).
//
// keyswitch_lwe_ciphertext (core_crypto/algorithms/lwe_keyswitch.rs:96-170) for a batch is the
// integer matrix product  Out[B x (n+1)] = Body - D[B x K] * KSK[K x (n+1)]  (K = k*N*level), with D the
// balanced gadget digits (|d| <= B/2 <= 64, decomposer.rs:98-152) and KSK in Z/2^64.  Splitting every
// KSK word into its 8 bytes turns it into an exact s8 x u8 -> s32 GEMM with 8x the columns:
//     acc[b][col][limb] = sum_k D[b][k] * byte_limb(KSK[k][col])         (|acc| <= K * 64 * 255 < 2^31)
//     Out[b][col]      = body - sum_limb (int64)acc[b][col][limb] << (8 limb)   (mod 2^64)
// so the result is bit-identical to the reference (wrapping integer arithmetic is associative).
//
// Three kernels:
//   ksk_limbs_kernel  (key load)   KSK -> byte limbs, pre-tiled in the UMMA "interleaved" (no swizzle)
//                                  K-major core-matrix order so one bulk copy fills one pipeline stage
//   ks_digits_kernel  (per batch)  ciphertext masks -> s8 digits, same pre-tiled order
//   ks_mma_kernel     (per batch)  128 ciphertexts x 32 output words (256 limb columns) per CTA; a
//                                  producer thread streams the tiles with cp.async.bulk + mbarriers, one
//                                  thread issues tcgen05.mma into a 128 x 256 s32 TMEM accumulator, four
//                                  warps recombine the limbs and write the small ciphertexts.
//
// Tile geometry: a K-stage is 32 input mask elements x level digits = 32*level bytes of K per row
// (a multiple of the MMA K of 32 bytes).  Shared-memory tile: [16-byte K chunk][row][16 B], i.e. core
// matrices of 8 rows x 16 B, SBO = 128 B between 8-row groups, LBO = rows*16 B between K chunks.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "pbs_common.cuh"   // mbarrier / bulk-copy helpers, tmem helpers

namespace b200 {

constexpr int kKmM = 128;        // ciphertexts per CTA (UMMA M)
constexpr int kKmCols = 32;      // output words per CTA
constexpr int kKmN = kKmCols * 8;  // limb columns per CTA (UMMA N = 256)
constexpr int kKmIC = 32;        // input mask elements per K-stage
constexpr int kKmMaxLevel = 8;

struct KsMmaGeom {
    int n_in, out_size, level, base_log;
    int k_stages;       // ceil(n_in / 32)
    int n_tiles;        // ceil(out_size / 32)
    int stage_k_bytes;  // 32 * level
    __host__ __device__ size_t a_stage_bytes() const { return (size_t)kKmM * stage_k_bytes; }
    __host__ __device__ size_t b_stage_bytes() const { return (size_t)kKmN * stage_k_bytes; }
    __host__ __device__ size_t b_total_bytes() const { return (size_t)n_tiles * k_stages * b_stage_bytes(); }
    __host__ __device__ size_t a_total_bytes(size_t batch) const {
        return ((batch + kKmM - 1) / kKmM) * (size_t)k_stages * a_stage_bytes();
    }
};

inline KsMmaGeom ks_mma_geom(int n_in, int out_size, int level, int base_log) {
    KsMmaGeom g{};
    g.n_in = n_in; g.out_size = out_size; g.level = level; g.base_log = base_log;
    g.k_stages = (n_in + kKmIC - 1) / kKmIC;
    g.n_tiles = (out_size + kKmCols - 1) / kKmCols;
    g.stage_k_bytes = kKmIC * level;
    return g;
}
inline int ks_mma_pipeline_stages(const KsMmaGeom &g) {
    const size_t per = g.a_stage_bytes() + g.b_stage_bytes();
    int s = (int)((size_t)(220 * 1024) / per);
    return s > 4 ? 4 : s;
}
inline size_t ks_mma_smem_bytes(const KsMmaGeom &g) {
    return 1024 + (size_t)ks_mma_pipeline_stages(g) * (g.a_stage_bytes() + g.b_stage_bytes());
}

// ---- key load: KSK words -> byte limbs in tile order ------------------------------------------
// b_tiled[n_tile][k_stage][chunk c][n_local = col_local*8 + limb][16 B]; byte b of chunk c is K index
// kk = 16 c + b of the stage = (mask element 32*ks + kk / level, level index kk % level), i.e. KSK row
// (i * level + li) exactly as the reference stores it (lwe_keyswitch_key_generation.rs:109-111).
__global__ void __launch_bounds__(kKmN) ksk_limbs_kernel(const uint64_t *__restrict__ ksk, uint8_t *__restrict__ b_tiled,
                                                         const KsMmaGeom g) {
    const int ks = blockIdx.x, nt = blockIdx.y, nl = threadIdx.x;
    const int col = nt * kKmCols + (nl >> 3), limb = nl & 7;
    const int chunks = g.stage_k_bytes / 16;
    uint8_t *dst = b_tiled + ((size_t)nt * g.k_stages + ks) * g.b_stage_bytes();
    for (int c = 0; c < chunks; c++) {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int b = 0; b < 16; b++) {
            const int kk = 16 * c + b;
            const int i = ks * kKmIC + kk / g.level, li = kk % g.level;
            uint32_t byte = 0;
            if (col < g.out_size && i < g.n_in)
                byte = (uint32_t)(ksk[((size_t)i * g.level + li) * g.out_size + col] >> (8 * limb)) & 0xFFu;
            w[b >> 2] |= byte << (8 * (b & 3));
        }
        *reinterpret_cast<uint4 *>(dst + ((size_t)c * kKmN + nl) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---- per batch: mask elements -> signed digits in tile order ----------------------------------
// a_tiled[m_tile][k_stage][chunk c][row][16 B].  SignedDecomposer::decompose (decomposer.rs:98-152,
// iter.rs:120-127): digit li is the li-th one the reference's iterator yields (level l first).
__global__ void __launch_bounds__(256) ks_digits_kernel(const uint64_t *__restrict__ in, uint8_t *__restrict__ a_tiled,
                                                        const KsMmaGeom g, const int batch) {
    extern __shared__ __align__(16) unsigned char dsm[];   // [chunk][row][16]
    const int ks = blockIdx.x, mt = blockIdx.y;
    const int rep_bits = g.base_log * g.level;
    const int shift = 64 - rep_bits - 1;
    const uint64_t mod_b_mask = ((uint64_t)1 << g.base_log) - 1;
    for (int e = threadIdx.x; e < kKmM * kKmIC; e += 256) {
        const int ii = e % kKmIC, row = e / kKmIC;
        const int i = ks * kKmIC + ii, ct = mt * kKmM + row;
        uint64_t state = 0;
        if (ct < batch && i < g.n_in) {
            const uint64_t x = in[(size_t)ct * (g.n_in + 1) + i];
            const uint64_t res = ((x >> shift) + 1) & ~(uint64_t)1;
            state = (res << shift) >> (64 - rep_bits);
        }
        for (int li = 0; li < g.level; li++) {
            uint64_t d = state & mod_b_mask;
            state >>= g.base_log;
            uint64_t carry = ((d - 1) | state) & d;
            carry >>= (g.base_log - 1);
            state += carry;
            d -= carry << g.base_log;   // signed digit in [-B/2, B/2]
            const int kk = ii * g.level + li;
            dsm[((kk >> 4) * kKmM + row) * 16 + (kk & 15)] = (uint8_t)(int8_t)(int64_t)d;
        }
    }
    __syncthreads();
    const size_t bytes = g.a_stage_bytes();
    uint4 *dst = reinterpret_cast<uint4 *>(a_tiled + ((size_t)mt * g.k_stages + ks) * bytes);
    const uint4 *src = reinterpret_cast<const uint4 *>(dsm);
    for (int e = threadIdx.x; e < (int)(bytes / 16); e += 256) dst[e] = src[e];
}

// ---- UMMA plumbing -----------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start >> 4, [16,30) leading (K-chunk) byte offset >> 4, [32,46) stride (8-row group) byte
// offset >> 4, [46,48) version = 1, [61,64) layout type = 0.
__device__ __forceinline__ uint64_t umma_desc_kmajor(const uint32_t smem_addr, const uint32_t lbo_bytes, const uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor (InstrDescriptor): D = s32, A = signed 8 bit, B = unsigned 8 bit, both K-major
constexpr uint32_t kKmIdesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(kKmN >> 3) << 17) | ((uint32_t)(kKmM >> 4) << 24);

__device__ __forceinline__ void umma_i8(const uint32_t tmem_d, const uint64_t desc_a, const uint64_t desc_b, const uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kKmIdesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t *bar, const uint32_t parity) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "KM_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra KM_DONE;\n\t"
        "bra KM_WAIT;\n\t"
        "KM_DONE:\n\t"
        "}" ::"r"(addr), "r"(parity) : "memory");
}

struct KsMmaArgs {
    const uint8_t *a_tiled;   // digits
    const uint8_t *b_tiled;   // KSK limbs
    const uint64_t *in;       // [batch][n_in + 1] (body)
    uint64_t *out;            // [batch][out_size]
    KsMmaGeom g;
    int batch, stages;
};

__global__ void __launch_bounds__(128, 1) ks_mma_kernel(const KsMmaArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // consecutive CTAs share the KSK tile (nt) and differ in the ciphertext tile (mt): a wave of 148 CTAs then touches ~5 of the
    // 24 KSK column tiles instead of all of them, so the 63 MB of limbs stream from HBM once per launch, not once per wave
    const int mt = blockIdx.x, nt = blockIdx.y;
    const KsMmaGeom &g = a.g;
    const int S = a.stages, KS = g.k_stages;
    const uint32_t a_bytes = (uint32_t)g.a_stage_bytes(), b_bytes = (uint32_t)g.b_stage_bytes();

    uint64_t *full = reinterpret_cast<uint64_t *>(smem);          // [S]
    uint64_t *empty = full + 4;                                   // [S]
    uint64_t *accum = full + 8;
    uint32_t *slot = reinterpret_cast<uint32_t *>(full + 9);
    unsigned char *tiles = smem + 1024;                           // [S][A | B]

    if (warp == 0) tmem_alloc(slot, kKmN);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(accum, 1);
    }
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tmem_d = *slot;

    if (threadIdx.x == 0) {
        // producer: one bulk copy per operand per stage
        const uint8_t *ga = a.a_tiled + (size_t)mt * KS * a_bytes;
        const uint8_t *gb = a.b_tiled + (size_t)nt * KS * b_bytes;
        for (int ks = 0; ks < KS; ks++) {
            const int s = ks % S;
            if (ks >= S) mbar_wait_parity(empty + s, (uint32_t)((ks / S - 1) & 1));
            unsigned char *sa = tiles + (size_t)s * (a_bytes + b_bytes);
            mbar_arrive_expect_tx(full + s, a_bytes + b_bytes);
            bulk_g2s(sa, ga + (size_t)ks * a_bytes, a_bytes, full + s);
            bulk_g2s(sa + a_bytes, gb + (size_t)ks * b_bytes, b_bytes, full + s);
        }
    } else if (threadIdx.x == 32) {
        // MMA issuer: `level` instructions of K = 32 bytes per stage
        const uint32_t lbo_a = kKmM * 16, lbo_b = kKmN * 16, sbo = 128;
        for (int ks = 0; ks < KS; ks++) {
            const int s = ks % S;
            mbar_wait_parity(full + s, (uint32_t)((ks / S) & 1));
            tmem_fence_after();
            const uint32_t sa = (uint32_t)__cvta_generic_to_shared(tiles + (size_t)s * (a_bytes + b_bytes));
            const uint32_t sb = sa + a_bytes;
            for (int j = 0; j < g.level; j++) {
                const uint64_t da = umma_desc_kmajor(sa + 2u * j * lbo_a, lbo_a, sbo);
                const uint64_t db = umma_desc_kmajor(sb + 2u * j * lbo_b, lbo_b, sbo);
                umma_i8(tmem_d, da, db, (ks | j) ? 1u : 0u);
            }
            umma_commit(empty + s);   // the stage may be refilled once these MMAs have read it
        }
        umma_commit(accum);
    }

    // epilogue: thread = one ciphertext (TMEM lane), 4 output words (32 limb columns) per TMEM load
    __syncwarp();
    mbar_wait_parity(accum, 0);
    tmem_fence_after();
    // thread = one ciphertext (TMEM lane): recombine the 8 limbs of every output word, stage the warp's 32 x 32 word tile in
    // shared memory (the pipeline buffers are idle by now: every MMA that read them has completed before `accum` fires) and
    // write it out row by row, so that a warp stores 256 contiguous bytes per instruction instead of 32 words 743 words apart
    const int row0 = mt * kKmM + warp * 32;
    const uint32_t t_row = tmem_d + (((uint32_t)warp * 32u) << 16);
    uint64_t *stage = reinterpret_cast<uint64_t *>(tiles) + (size_t)warp * 32 * 33;   // [32 rows][33]: padded against bank conflicts
#pragma unroll 1
    for (int c4 = 0; c4 < kKmCols / 4; c4++) {
        uint32_t r[32];
        tmem_ld32(t_row + c4 * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            uint64_t v = 0;
#pragma unroll
            for (int l = 0; l < 8; l++) v += (uint64_t)(int64_t)(int32_t)r[cc * 8 + l] << (8 * l);
            stage[lane * 33 + c4 * 4 + cc] = v;
        }
    }
    __syncwarp();
    const int col = nt * kKmCols + lane;
#pragma unroll 4
    for (int rr = 0; rr < 32; rr++) {
        const int row = row0 + rr;
        if (row < a.batch && col < g.out_size) {
            const uint64_t body = col == g.out_size - 1 ? a.in[(size_t)row * (g.n_in + 1) + g.n_in] : 0;
            a.out[(size_t)row * g.out_size + col] = body - stage[rr * 33 + lane];
        }
    }

    tmem_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, kKmN);
}

}  // namespace b200
