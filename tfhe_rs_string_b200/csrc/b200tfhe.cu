// b200tfhe.cu -- context, device memory manager and C ABI of libb200tfhe.so (see include/b200tfhe.h).
// Host-side only plumbing lives here; the kernels are in pbs_kernel.cuh / ks_kernel.cuh.
#include "../../include/b200tfhe.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "ks_kernel.cuh"
#include "ks_mma.cuh"
#include "pbs_kernel.cuh"
#include "pbs_kernel3.cuh"
#include "pbs_kernel_lat.cuh"
#include "programs.hpp"

using namespace b200;

namespace {

std::mutex g_err_mu;
std::string g_err;

void set_global_error(const std::string &s) {
    std::lock_guard<std::mutex> l(g_err_mu);
    g_err = s;
}

struct EventPair {
    cudaEvent_t a, b;
};

}  // namespace

struct b200tfhe_ctx {
    b200tfhe_params p{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;   // copy streams of the pipelined host-buffer path
    int sm_count = 148;
    std::mutex mu;
    mutable std::string err;

    // key arena: [Fourier BSK][KSK][colsum]
    unsigned char *arena = nullptr;
    size_t arena_bytes = 0, off_bsk = 0, off_ksk = 0, off_colsum = 0, off_ksk_limbs = 0;
    KsMmaGeom ks_geom{};
    bool ksk_loaded = false, bsk_loaded = false;
    double2 *d_twid = nullptr;

    // LUT store (content addressed)
    uint64_t *d_luts = nullptr;
    size_t lut_cap = 0;
    std::vector<std::vector<uint64_t>> h_luts;
    std::unordered_multimap<uint64_t, uint32_t> lut_hash;

    // workspace for the host-buffer entry points and the fused KS->PBS
    uint64_t *d_in = nullptr, *d_small = nullptr, *d_out = nullptr;
    uint32_t *d_lut_idx = nullptr;
    uint8_t *d_digits = nullptr;   // keyswitch digits, tile order (ks_mma.cuh)
    size_t ws_cap = 0;

    int pbs_variant = 3;
    bool profiling = false;
    std::vector<EventPair> ev_ks, ev_pbs;
    double ks_ms = 0, pbs_ms = 0;
    uint64_t ks_launches = 0, pbs_launches = 0;

    size_t big_size() const { return (size_t)p.glwe_dimension * p.polynomial_size + 1; }
    size_t small_size() const { return (size_t)p.lwe_dimension + 1; }
    size_t glwe_len() const { return (size_t)(p.glwe_dimension + 1) * p.polynomial_size; }
    size_t ksk_len() const { return (size_t)p.glwe_dimension * p.polynomial_size * p.ks_level * small_size(); }
    size_t bsk_len() const {
        return (size_t)p.lwe_dimension * p.pbs_level * (p.glwe_dimension + 1) * (p.glwe_dimension + 1) * p.polynomial_size;
    }
    double2 *d_bsk() const { return reinterpret_cast<double2 *>(arena + off_bsk); }
    uint64_t *d_ksk() const { return reinterpret_cast<uint64_t *>(arena + off_ksk); }
    uint64_t *d_colsum() const { return reinterpret_cast<uint64_t *>(arena + off_colsum); }
    uint8_t *d_ksk_limbs() const { return arena + off_ksk_limbs; }
};

namespace {

int fail(const b200tfhe_ctx *ctx, const std::string &msg) {
    if (ctx) ctx->err = msg;
    set_global_error(msg);
    return 1;
}

#define CU_TRY(ctx, expr)                                                                        \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return fail(ctx, std::string(#expr) + ": " + cudaGetErrorString(e__));              \
    } while (0)

#define ARG_TRY(ctx, cond, msg)                                                                  \
    do {                                                                                         \
        if (!(cond)) return fail(ctx, std::string("invalid argument: ") + (msg));                \
    } while (0)

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// T'[k1][l] = exp(-2*pi*i*l*k1/1024) * exp(i*pi*l/2048) = exp(i*pi*(l*(1-4*k1) mod 4096)/2048)
void make_twiddles(std::vector<double2> &t) {
    t.resize(1024);
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k1 = 0; k1 < 32; k1++)
        for (int l = 0; l < 32; l++) {
            int r = (l * (1 - 4 * k1)) % 4096;
            if (r < 0) r += 4096;
            long double ang = pi * (long double)r / 2048.0L;
            t[k1 * 32 + l] = make_double2((double)cosl(ang), (double)sinl(ang));
        }
}

int ensure_workspace(b200tfhe_ctx *ctx, size_t batch) {
    if (batch <= ctx->ws_cap) return 0;
    size_t cap = std::max<size_t>(batch, 256);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_in); cudaFree(ctx->d_small); cudaFree(ctx->d_out); cudaFree(ctx->d_lut_idx); cudaFree(ctx->d_digits);
    ctx->d_in = ctx->d_small = ctx->d_out = nullptr; ctx->d_lut_idx = nullptr; ctx->d_digits = nullptr; ctx->ws_cap = 0;
    CU_TRY(ctx, cudaMalloc(&ctx->d_in, cap * ctx->big_size() * sizeof(uint64_t)));
    CU_TRY(ctx, cudaMalloc(&ctx->d_small, cap * ctx->small_size() * sizeof(uint64_t)));
    CU_TRY(ctx, cudaMalloc(&ctx->d_out, cap * ctx->big_size() * sizeof(uint64_t)));
    CU_TRY(ctx, cudaMalloc(&ctx->d_lut_idx, cap * sizeof(uint32_t)));
    CU_TRY(ctx, cudaMalloc(&ctx->d_digits, ctx->ks_geom.a_total_bytes(cap)));
    ctx->ws_cap = cap;
    return 0;
}

int ensure_lut_capacity(b200tfhe_ctx *ctx, size_t n) {
    if (n <= ctx->lut_cap) return 0;
    size_t cap = std::max<size_t>(64, ctx->lut_cap * 2);
    while (cap < n) cap *= 2;
    uint64_t *nd = nullptr;
    CU_TRY(ctx, cudaMalloc(&nd, cap * ctx->glwe_len() * sizeof(uint64_t)));
    if (ctx->d_luts) {
        CU_TRY(ctx, cudaMemcpyAsync(nd, ctx->d_luts, ctx->h_luts.size() * ctx->glwe_len() * sizeof(uint64_t),
                                    cudaMemcpyDeviceToDevice, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->d_luts);
    }
    ctx->d_luts = nd;
    ctx->lut_cap = cap;
    return 0;
}

uint64_t fnv1a(const uint64_t *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 1099511628211ull;
    }
    return h;
}

void prof_begin(b200tfhe_ctx *ctx, std::vector<EventPair> &v) {
    if (!ctx->profiling) return;
    EventPair e{};
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, ctx->stream);
    v.push_back(e);
}
void prof_end(b200tfhe_ctx *ctx, std::vector<EventPair> &v) {
    if (!ctx->profiling) return;
    cudaEventRecord(v.back().b, ctx->stream);
}

int launch_ks(b200tfhe_ctx *ctx, const uint64_t *d_in, uint64_t *d_out, size_t batch) {
    if (!ctx->ksk_loaded) return fail(ctx, "keyswitch key not loaded");
    KsArgs a{};
    a.in = d_in; a.ksk = ctx->d_ksk(); a.colsum = ctx->d_colsum(); a.out = d_out;
    a.batch = (int)batch; a.n_in = (int)(ctx->p.glwe_dimension * ctx->p.polynomial_size);
    a.out_size = (int)ctx->small_size(); a.base_log = (int)ctx->p.ks_base_log; a.level = (int)ctx->p.ks_level;
    dim3 grid((a.out_size + kKsBN - 1) / kKsBN, (unsigned)((batch + kKsBM - 1) / kKsBM));
    prof_begin(ctx, ctx->ev_ks);
    static const int dev_variant = getenv("B200TFHE_KS_VARIANT") ? atoi(getenv("B200TFHE_KS_VARIANT")) : -1;  // development knob
    if (dev_variant < 0) {
        // default: tensor-core path (ks_mma.cuh): digits pre-pass + s8 x u8 tcgen05 GEMM over the KSK byte limbs
        if (int rc = ensure_workspace(ctx, batch)) return rc;
        const KsMmaGeom &g = ctx->ks_geom;
        const unsigned m_tiles = (unsigned)((batch + kKmM - 1) / kKmM);
        ks_digits_kernel<<<dim3(g.k_stages, m_tiles), 256, g.a_stage_bytes(), ctx->stream>>>(d_in, ctx->d_digits, g, (int)batch);
        KsMmaArgs m{};
        m.a_tiled = ctx->d_digits; m.b_tiled = ctx->d_ksk_limbs(); m.in = d_in; m.out = d_out; m.g = g;
        m.batch = (int)batch; m.stages = ks_mma_pipeline_stages(g);
        ks_mma_kernel<<<dim3(g.n_tiles, m_tiles), 128, ks_mma_smem_bytes(g), ctx->stream>>>(m);
        prof_end(ctx, ctx->ev_ks);
        CU_TRY(ctx, cudaGetLastError());
        ctx->ks_launches++;
        return 0;
    }
    const size_t smem = ks_smem_bytes(a.level);
    switch (dev_variant) {
        case 1: ks_kernel<4, 2><<<grid, kKsThreads, smem, ctx->stream>>>(a); break;
        case 2: ks_kernel<4, 1><<<grid, kKsThreads, smem, ctx->stream>>>(a); break;
        case 3: ks_kernel<1, 1><<<grid, kKsThreads, smem, ctx->stream>>>(a); break;
        default: ks_kernel<2, 1><<<grid, kKsThreads, smem, ctx->stream>>>(a); break;
    }
    prof_end(ctx, ctx->ev_ks);
    CU_TRY(ctx, cudaGetLastError());
    ctx->ks_launches++;
    return 0;
}

template <int CTS, bool BSK_SMEM>
int launch_pbs_variant(b200tfhe_ctx *ctx, const PbsArgs &a) {
    static bool configured[16] = {};
    constexpr size_t smem = pbs_smem_bytes<CTS, BSK_SMEM>();
    if (!configured[ctx->device & 15]) {
        CU_TRY(ctx, cudaFuncSetAttribute(pbs_kernel<CTS, BSK_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = true;
    }
    const unsigned grid = (unsigned)((a.batch + CTS - 1) / CTS);
    pbs_kernel<CTS, BSK_SMEM><<<grid, CTS * 64, smem, ctx->stream>>>(a);
    return 0;
}

template <int CTS>
int launch_pbs3_cts(b200tfhe_ctx *ctx, const PbsArgs &a) {
    static bool configured[16] = {};
    constexpr size_t smem = pbs3_smem_bytes<CTS>();
    if (!configured[ctx->device & 15]) {
        CU_TRY(ctx, cudaFuncSetAttribute(pbs_kernel3<CTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = true;
    }
    const unsigned grid = (unsigned)((a.batch + CTS - 1) / CTS);
#ifdef B200TFHE_TIMELINE
    if (const char *dump = getenv("B200TFHE_PBS_TIMELINE")) {   // development: phase timestamps of CTA 0, steps 100..107
        PbsArgs b = a;
        const size_t n = 8 * 8 * 16;
        CU_TRY(ctx, cudaMalloc(&b.dbg, n * sizeof(long long)));
        CU_TRY(ctx, cudaMemsetAsync(b.dbg, 0, n * sizeof(long long), ctx->stream));
        pbs_kernel3<CTS><<<grid, CTS * 64, smem, ctx->stream>>>(b);
        std::vector<long long> h(n);
        CU_TRY(ctx, cudaMemcpyAsync(h.data(), b.dbg, n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(b.dbg);
        if (FILE *f = fopen(dump, "w")) {
            for (size_t r = 0; r < 64; r++) {
                fprintf(f, "%zu %zu", r / 8, r % 8);
                for (int k = 0; k < 11; k++) fprintf(f, " %lld", h[r * 16 + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
        return 0;
    }
#endif
    pbs_kernel3<CTS><<<grid, CTS * 64, smem, ctx->stream>>>(a);
    return 0;
}

template <int CTS>
int launch_pbs_lat(b200tfhe_ctx *ctx, const PbsArgs &a) {
    static bool configured[16] = {};
    constexpr size_t smem = pbs_lat_smem_bytes<CTS>();
    if (!configured[ctx->device & 15]) {
        CU_TRY(ctx, cudaFuncSetAttribute(pbs_lat_kernel<CTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[ctx->device & 15] = true;
    }
#ifdef B200TFHE_TIMELINE
    if (const char *dump = getenv("B200TFHE_PBS_TIMELINE")) {
        PbsArgs b = a;
        const size_t n = 8 * 8 * 16;
        CU_TRY(ctx, cudaMalloc(&b.dbg, n * sizeof(long long)));
        CU_TRY(ctx, cudaMemsetAsync(b.dbg, 0, n * sizeof(long long), ctx->stream));
        pbs_lat_kernel<CTS><<<(unsigned)((a.batch + CTS - 1) / CTS), CTS == 4 ? 512 : 256, smem, ctx->stream>>>(b);
        std::vector<long long> h(n);
        CU_TRY(ctx, cudaMemcpyAsync(h.data(), b.dbg, n * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(b.dbg);
        if (FILE *f = fopen(dump, "w")) {
            for (size_t r = 0; r < 64; r++) {
                fprintf(f, "%zu %zu", r / 8, r % 8);
                for (int k = 0; k < 11; k++) fprintf(f, " %lld", h[r * 16 + k]);
                fprintf(f, "\n");
            }
            fclose(f);
        }
        return 0;
    }
#endif
    pbs_lat_kernel<CTS><<<(unsigned)((a.batch + CTS - 1) / CTS), CTS == 4 ? 512 : 256, smem, ctx->stream>>>(a);
    return 0;
}

// Ciphertexts per CTA: the fewest that still fit the batch into the minimum number of waves over the
// SMs (one CTA per SM).  Large batches get 4 (throughput); a dependency level with few bootstraps
// gets 1-3, which shortens every CMUX step (fewer warps share an SM sub-partition) and so the
// latency of the level.
int launch_pbs3(b200tfhe_ctx *ctx, const PbsArgs &a) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const long long waves = (a.batch + 4LL * sms - 1) / (4LL * sms);
    const long long per_cta = (a.batch + waves * sms - 1) / (waves * sms);
    // one or two ciphertexts per SM: the latency kernel (two warps per polynomial, pbs_kernel_lat.cuh)
    static const int dev_lat = getenv("B200TFHE_PBS_LAT") ? atoi(getenv("B200TFHE_PBS_LAT")) : 1;  // development knob
    if (dev_lat && per_cta == 1) return launch_pbs_lat<1>(ctx, a);
    if (dev_lat && per_cta == 2) return launch_pbs_lat<2>(ctx, a);
    if (dev_lat >= 2 && per_cta == 4) return launch_pbs_lat<4>(ctx, a);   // experiment: 16-warp throughput configuration
    switch ((int)per_cta) {
        case 1: return launch_pbs3_cts<1>(ctx, a);
        case 2: return launch_pbs3_cts<2>(ctx, a);
        case 3: return launch_pbs3_cts<3>(ctx, a);
        default: return launch_pbs3_cts<4>(ctx, a);
    }
}

int launch_pbs(b200tfhe_ctx *ctx, const uint64_t *d_small, const uint32_t *d_lut_idx, uint64_t *d_out, size_t batch) {
    if (!ctx->bsk_loaded) return fail(ctx, "bootstrap key not loaded");
    if (ctx->h_luts.empty()) return fail(ctx, "no lookup table registered");
    PbsArgs a{};
    a.lwe_small = d_small; a.lut_idx = d_lut_idx; a.luts = ctx->d_luts; a.bsk = ctx->d_bsk(); a.twid = ctx->d_twid;
    a.out = d_out; a.batch = (int)batch; a.n = (int)ctx->p.lwe_dimension;
    prof_begin(ctx, ctx->ev_pbs);
    int rc;
    switch (ctx->pbs_variant) {
        case 1: rc = launch_pbs_variant<4, false>(ctx, a); break;
        case 2: rc = launch_pbs_variant<6, false>(ctx, a); break;
        case 0: rc = launch_pbs_variant<4, true>(ctx, a); break;
        default: rc = launch_pbs3(ctx, a); break;
    }
    prof_end(ctx, ctx->ev_pbs);
    if (rc) return rc;
    CU_TRY(ctx, cudaGetLastError());
    ctx->pbs_launches++;
    return 0;
}

int check_ready(b200tfhe_ctx *ctx) {
    if (!ctx) return fail(nullptr, "null context");
    CU_TRY(ctx, cudaSetDevice(ctx->device));
    return 0;
}

}  // namespace

extern "C" {

int b200tfhe_last_global_error(char *buf, size_t buf_len) {
    std::lock_guard<std::mutex> l(g_err_mu);
    if (buf && buf_len) {
        std::snprintf(buf, buf_len, "%s", g_err.c_str());
    }
    return 0;
}

int b200tfhe_last_error(const b200tfhe_ctx *ctx, char *buf, size_t buf_len) {
    if (!ctx) return b200tfhe_last_global_error(buf, buf_len);
    if (buf && buf_len) std::snprintf(buf, buf_len, "%s", ctx->err.c_str());
    return 0;
}

int b200tfhe_ctx_create(const b200tfhe_params *params, int device, b200tfhe_ctx **out) {
    if (!out) return fail(nullptr, "invalid argument: out is null");
    *out = nullptr;  // c_api/shortint/server_key/pbs.rs:54-58 nulls results on entry too
    if (!params) return fail(nullptr, "invalid argument: params is null");
    const b200tfhe_params &p = *params;
    if (p.glwe_dimension != 1 || p.polynomial_size != 2048 || p.pbs_level != 1)
        return fail(nullptr, "unsupported parameters: kernels require glwe_dimension=1, polynomial_size=2048, pbs_level=1");
    if (p.pbs_base_log != 23)
        return fail(nullptr, "unsupported parameters: pbs_base_log must be 23");
    if (p.lwe_dimension == 0 || p.lwe_dimension > (uint32_t)kMaxSmallDim)
        return fail(nullptr, "unsupported parameters: lwe_dimension must be in [1, 1024]");
    if (p.ks_base_log == 0 || p.ks_base_log > 7 || p.ks_level == 0 || p.ks_base_log * p.ks_level >= 64)
        return fail(nullptr, "unsupported parameters: ks_base_log must be in [1,7] and base_log*level < 64");
    if (p.message_modulus == 0 || p.carry_modulus == 0 ||
        p.polynomial_size % (p.message_modulus * p.carry_modulus) != 0)
        return fail(nullptr, "unsupported parameters: message_modulus*carry_modulus must divide polynomial_size");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                 (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
    if (device < 0 || device >= n_dev) return fail(nullptr, "invalid argument: device index out of range");
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, "unsupported GPU: libb200tfhe is built for sm_100a (B200) only");

    auto *ctx = new b200tfhe_ctx();
    ctx->p = p;
    ctx->device = device;
    if (const char *v = getenv("B200TFHE_PBS_VARIANT")) ctx->pbs_variant = atoi(v);   // development knob
    auto bail = [&](const std::string &m) {
        fail(nullptr, m);
        delete ctx;
        return 1;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking) != cudaSuccess)
        return bail("cudaStreamCreate failed");
    ctx->sm_count = prop.multiProcessorCount;
    ctx->off_bsk = 0;
    ctx->off_ksk = align_up(ctx->bsk_len() / 2 * sizeof(double2), 256);  // N/2 complex per poly
    ctx->off_colsum = ctx->off_ksk + align_up(ctx->ksk_len() * sizeof(uint64_t), 256);
    ctx->ks_geom = ks_mma_geom((int)(p.glwe_dimension * p.polynomial_size), (int)ctx->small_size(), (int)p.ks_level, (int)p.ks_base_log);
    ctx->off_ksk_limbs = ctx->off_colsum + align_up(ctx->small_size() * sizeof(uint64_t), 256);
    ctx->arena_bytes = ctx->off_ksk_limbs + align_up(ctx->ks_geom.b_total_bytes(), 256);
    if (cudaMalloc(&ctx->arena, ctx->arena_bytes) != cudaSuccess) return bail("cudaMalloc(key arena) failed");
    std::vector<double2> tw;
    make_twiddles(tw);
    if (cudaMalloc(&ctx->d_twid, tw.size() * sizeof(double2)) != cudaSuccess) return bail("cudaMalloc(twiddles) failed");
    if (cudaMemcpy(ctx->d_twid, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess)
        return bail("cudaMemcpy(twiddles) failed");
    {
        const int ks_smem = (int)ks_smem_bytes((int)p.ks_level);
        if (cudaFuncSetAttribute(ks_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ks_smem) != cudaSuccess ||
            cudaFuncSetAttribute(ks_kernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ks_smem) != cudaSuccess ||
            cudaFuncSetAttribute(ks_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ks_smem) != cudaSuccess ||
            cudaFuncSetAttribute(ks_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ks_smem) != cudaSuccess)
            return bail("cudaFuncSetAttribute(ks_kernel) failed");
        if (ks_mma_pipeline_stages(ctx->ks_geom) < 2 || p.ks_level > (uint32_t)kKmMaxLevel)
            return bail("unsupported parameters: keyswitch level too large for the tensor-core pipeline");
        if (cudaFuncSetAttribute(ks_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ks_mma_smem_bytes(ctx->ks_geom)) != cudaSuccess ||
            cudaFuncSetAttribute(ks_digits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->ks_geom.a_stage_bytes()) != cudaSuccess)
            return bail("cudaFuncSetAttribute(ks_mma_kernel) failed");
    }
    *out = ctx;
    return 0;
}

int b200tfhe_ctx_destroy(b200tfhe_ctx *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &e : ctx->ev_ks) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto &e : ctx->ev_pbs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    cudaFree(ctx->arena); cudaFree(ctx->d_twid); cudaFree(ctx->d_luts);
    cudaFree(ctx->d_in); cudaFree(ctx->d_small); cudaFree(ctx->d_out); cudaFree(ctx->d_lut_idx); cudaFree(ctx->d_digits);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->s_h2d);
    cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
    return 0;
}

int b200tfhe_load_ksk(b200tfhe_ctx *ctx, const uint64_t *ksk, size_t n_u64) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, ksk != nullptr, "ksk is null");
    ARG_TRY(ctx, n_u64 == ctx->ksk_len(), "ksk length does not match parameters (k*N*ks_level*(n+1))");
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_ksk(), ksk, n_u64 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaMemsetAsync(ctx->d_colsum(), 0, ctx->small_size() * sizeof(uint64_t), ctx->stream));
    const int out_size = (int)ctx->small_size();
    const size_t n_rows = (size_t)ctx->p.glwe_dimension * ctx->p.polynomial_size * ctx->p.ks_level;
    dim3 grid((out_size + 127) / 128, 64);
    ks_colsum_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->d_ksk(), ctx->d_colsum(), n_rows, out_size,
                                                    (uint64_t)1 << (ctx->p.ks_base_log - 1));
    CU_TRY(ctx, cudaGetLastError());
    ksk_limbs_kernel<<<dim3(ctx->ks_geom.k_stages, ctx->ks_geom.n_tiles), kKmN, 0, ctx->stream>>>(ctx->d_ksk(), ctx->d_ksk_limbs(), ctx->ks_geom);
    CU_TRY(ctx, cudaGetLastError());
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->ksk_loaded = true;
    return 0;
}

int b200tfhe_load_bsk_standard(b200tfhe_ctx *ctx, const uint64_t *bsk, size_t n_u64) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, bsk != nullptr, "bsk is null");
    ARG_TRY(ctx, n_u64 == ctx->bsk_len(), "bsk length does not match parameters (n*pbs_level*(k+1)^2*N)");
    uint64_t *tmp = nullptr;
    CU_TRY(ctx, cudaMalloc(&tmp, n_u64 * sizeof(uint64_t)));
    cudaError_t e = cudaMemcpyAsync(tmp, bsk, n_u64 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        const int n_polys = (int)(n_u64 / kN);
        bsk_to_fourier_kernel<<<std::min(n_polys / 2 + 1, 148 * 8), 64, 0, ctx->stream>>>(tmp, ctx->d_bsk(), ctx->d_twid, n_polys);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    CU_TRY(ctx, e);
    ctx->bsk_loaded = true;
    return 0;
}

int b200tfhe_key_arena(b200tfhe_ctx *ctx, void **device_ptr, size_t *bytes) {
    if (int rc = check_ready(ctx)) return rc;
    ARG_TRY(ctx, device_ptr && bytes, "null output pointer");
    *device_ptr = ctx->arena;
    *bytes = ctx->arena_bytes;
    return 0;
}

int b200tfhe_keys_adopt(b200tfhe_ctx *ctx) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ctx->ksk_loaded = ctx->bsk_loaded = true;
    return 0;
}

int b200tfhe_register_lut(b200tfhe_ctx *ctx, const uint64_t *glwe_acc, uint32_t *id) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, glwe_acc && id, "null pointer");
    const size_t len = ctx->glwe_len();
    const uint64_t h = fnv1a(glwe_acc, len);
    auto range = ctx->lut_hash.equal_range(h);
    for (auto it = range.first; it != range.second; ++it)
        if (std::memcmp(ctx->h_luts[it->second].data(), glwe_acc, len * sizeof(uint64_t)) == 0) {
            *id = it->second;
            return 0;
        }
    if (int rc = ensure_lut_capacity(ctx, ctx->h_luts.size() + 1)) return rc;
    const uint32_t nid = (uint32_t)ctx->h_luts.size();
    ctx->h_luts.emplace_back(glwe_acc, glwe_acc + len);
    // copy from our own stable host copy so the caller's buffer can be reused immediately
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_luts + (size_t)nid * len, ctx->h_luts.back().data(), len * sizeof(uint64_t),
                                cudaMemcpyHostToDevice, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->lut_hash.emplace(h, nid);
    *id = nid;
    return 0;
}

int b200tfhe_register_lut_from_table(b200tfhe_ctx *ctx, const uint64_t *table, size_t table_len, uint32_t *id) {
    if (!ctx) return fail(nullptr, "null context");
    ARG_TRY(ctx, table && id, "null pointer");
    const b200tfhe_params &p = ctx->p;
    const size_t modulus_sup = (size_t)p.message_modulus * p.carry_modulus;
    ARG_TRY(ctx, table_len == modulus_sup, "table length must be message_modulus*carry_modulus");
    // fill_accumulator, shortint/engine/mod.rs:72-128
    const size_t N = p.polynomial_size, box = N / modulus_sup, half = box / 2;
    const uint64_t delta = ((uint64_t)1 << 63) / modulus_sup;
    std::vector<uint64_t> acc(ctx->glwe_len(), 0);
    uint64_t *body = acc.data() + (size_t)p.glwe_dimension * N;
    for (size_t i = 0; i < modulus_sup; i++)
        for (size_t j = 0; j < box; j++) body[i * box + j] = table[i] * delta;
    for (size_t j = 0; j < half; j++) body[j] = 0 - body[j];
    std::rotate(body, body + half, body + N);
    return b200tfhe_register_lut(ctx, acc.data(), id);
}

int b200tfhe_keyswitch_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, uint64_t *d_out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    return launch_ks(ctx, d_in, d_out, batch);
}

int b200tfhe_pbs_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    return launch_pbs(ctx, d_in, d_lut_id, d_out, batch);
}

int b200tfhe_ks_pbs_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    if (int rc = ensure_workspace(ctx, batch)) return rc;
    if (int rc = launch_ks(ctx, d_in, ctx->d_small, batch)) return rc;
    return launch_pbs(ctx, ctx->d_small, d_lut_id, d_out, batch);
}

int b200tfhe_pbs_ks_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    if (int rc = ensure_workspace(ctx, batch)) return rc;
    if (int rc = launch_pbs(ctx, d_in, d_lut_id, ctx->d_out, batch)) return rc;
    return launch_ks(ctx, ctx->d_out, d_out, batch);
}

static int validate_lut_ids(b200tfhe_ctx *ctx, const uint32_t *lut_id, size_t batch) {
    if (!lut_id) return 0;
    const uint32_t n = (uint32_t)ctx->h_luts.size();
    for (size_t b = 0; b < batch; b++)
        if (lut_id[b] >= n) return fail(ctx, "invalid argument: lut_id out of range");
    return 0;
}

int b200tfhe_keyswitch_batch(b200tfhe_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in && out, "null pointer");
    if (int rc = ensure_workspace(ctx, batch)) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_in, in, batch * ctx->big_size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_ks(ctx, ctx->d_in, ctx->d_small, batch)) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_small, batch * ctx->small_size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200tfhe_pbs_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in && out, "null pointer");
    if (int rc = validate_lut_ids(ctx, lut_id, batch)) return rc;
    if (int rc = ensure_workspace(ctx, batch)) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_small, in, batch * ctx->small_size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    if (lut_id) CU_TRY(ctx, cudaMemcpyAsync(ctx->d_lut_idx, lut_id, batch * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_pbs(ctx, ctx->d_small, lut_id ? ctx->d_lut_idx : nullptr, ctx->d_out, batch)) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_out, batch * ctx->big_size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200tfhe_ks_pbs_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in && out, "null pointer");
    if (int rc = validate_lut_ids(ctx, lut_id, batch)) return rc;
    if (int rc = ensure_workspace(ctx, batch)) return rc;
    const size_t big = ctx->big_size(), small = ctx->small_size();
    // Pipelined over chunks of one full wave (4 ciphertexts per SM): the H2D copy of chunk c+1 and the
    // D2H copy of chunk c-1 run on their own streams under the kernels of chunk c, so for pinned host
    // buffers only the first upload and the last download are exposed.
    const size_t chunk = (size_t)ctx->sm_count * 4;
    const size_t n_chunks = (batch + chunk - 1) / chunk;
    if (lut_id) CU_TRY(ctx, cudaMemcpyAsync(ctx->d_lut_idx, lut_id, batch * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->s_h2d));
    std::vector<cudaEvent_t> ev_in(n_chunks), ev_done(n_chunks);
    int rc = 0;
    cudaEvent_t ev_start;   // everything already queued on the compute stream (earlier async work on the workspaces) first
    CU_TRY(ctx, cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    CU_TRY(ctx, cudaEventRecord(ev_start, ctx->stream));
    CU_TRY(ctx, cudaStreamWaitEvent(ctx->s_h2d, ev_start, 0));
    for (size_t c = 0; c < n_chunks; c++) {
        cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ev_done[c], cudaEventDisableTiming);
    }
    for (size_t c = 0; c < n_chunks && !rc; c++) {
        const size_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
        cudaError_t e = cudaMemcpyAsync(ctx->d_in + b0 * big, in + b0 * big, nb * big * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->s_h2d);
        if (e == cudaSuccess) e = cudaEventRecord(ev_in[c], ctx->s_h2d);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev_in[c], 0);
        if (e != cudaSuccess) { rc = fail(ctx, std::string("ks_pbs_batch (upload): ") + cudaGetErrorString(e)); break; }
        rc = launch_ks(ctx, ctx->d_in + b0 * big, ctx->d_small + b0 * small, nb);
        if (!rc) rc = launch_pbs(ctx, ctx->d_small + b0 * small, lut_id ? ctx->d_lut_idx + b0 : nullptr, ctx->d_out + b0 * big, nb);
        if (rc) break;
        e = cudaEventRecord(ev_done[c], ctx->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_d2h, ev_done[c], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + b0 * big, ctx->d_out + b0 * big, nb * big * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->s_d2h);
        if (e != cudaSuccess) rc = fail(ctx, std::string("ks_pbs_batch (download): ") + cudaGetErrorString(e));
    }
    cudaError_t e1 = cudaStreamSynchronize(ctx->s_h2d), e2 = cudaStreamSynchronize(ctx->stream), e3 = cudaStreamSynchronize(ctx->s_d2h);
    for (size_t c = 0; c < n_chunks; c++) { cudaEventDestroy(ev_in[c]); cudaEventDestroy(ev_done[c]); }
    cudaEventDestroy(ev_start);
    if (rc) return rc;
    CU_TRY(ctx, e1); CU_TRY(ctx, e2); CU_TRY(ctx, e3);
    return 0;
}

int b200tfhe_pbs_ks_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in && out, "null pointer");
    if (int rc = validate_lut_ids(ctx, lut_id, batch)) return rc;
    if (int rc = ensure_workspace(ctx, batch)) return rc;
    const size_t small_bytes = batch * ctx->small_size() * sizeof(uint64_t);
    CU_TRY(ctx, cudaMemcpyAsync(ctx->d_small, in, small_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (lut_id) CU_TRY(ctx, cudaMemcpyAsync(ctx->d_lut_idx, lut_id, batch * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = launch_pbs(ctx, ctx->d_small, lut_id ? ctx->d_lut_idx : nullptr, ctx->d_out, batch)) return rc;
    if (int rc = launch_ks(ctx, ctx->d_out, ctx->d_small, batch)) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_small, small_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200tfhe_lwe_linear_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_x, const uint64_t *d_y, const int32_t *d_ia,
                                     const int32_t *d_ib, const int64_t *d_ca, const int64_t *d_cb,
                                     const uint64_t *d_pt, uint64_t *d_out, size_t batch, size_t lwe_size) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_x && d_ca && d_out, "null pointer");
    ARG_TRY(ctx, batch <= 65535, "batch too large for one launch (max 65535)");
    LinArgs a{};
    a.x = d_x; a.y = d_y; a.ia = d_ia; a.ib = d_ib; a.ca = d_ca; a.cb = d_cb; a.pt = d_pt; a.out = d_out;
    a.batch = (int)batch; a.size = (int)lwe_size;
    dim3 grid((unsigned)((lwe_size + 255) / 256), (unsigned)batch);
    lwe_linear_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    CU_TRY(ctx, cudaGetLastError());
    return 0;
}

int b200tfhe_sync(b200tfhe_ctx *ctx) {
    if (int rc = check_ready(ctx)) return rc;
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200tfhe_stream(b200tfhe_ctx *ctx, void **stream) {
    if (!ctx) return fail(nullptr, "null context");
    ARG_TRY(ctx, stream != nullptr, "null pointer");
    *stream = (void *)ctx->stream;
    return 0;
}

int b200tfhe_set_profiling(b200tfhe_ctx *ctx, int enabled) {
    if (!ctx) return fail(nullptr, "null context");
    std::lock_guard<std::mutex> l(ctx->mu);
    ctx->profiling = enabled != 0;
    return 0;
}

int b200tfhe_get_kernel_times(b200tfhe_ctx *ctx, double *ks_ms, uint64_t *ks_launches, double *pbs_ms,
                              uint64_t *pbs_launches, int reset) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    auto drain = [&](std::vector<EventPair> &v, double &acc) {
        for (auto &e : v) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) acc += ms;
            cudaEventDestroy(e.a);
            cudaEventDestroy(e.b);
        }
        v.clear();
    };
    drain(ctx->ev_ks, ctx->ks_ms);
    drain(ctx->ev_pbs, ctx->pbs_ms);
    if (ks_ms) *ks_ms = ctx->ks_ms;
    if (pbs_ms) *pbs_ms = ctx->pbs_ms;
    if (ks_launches) *ks_launches = ctx->ks_launches;
    if (pbs_launches) *pbs_launches = ctx->pbs_launches;
    if (reset) {
        ctx->ks_ms = ctx->pbs_ms = 0;
        ctx->ks_launches = ctx->pbs_launches = 0;
    }
    return 0;
}

int b200tfhe_set_pbs_variant(b200tfhe_ctx *ctx, int variant) {
    if (!ctx) return fail(nullptr, "null context");
    ARG_TRY(ctx, variant >= 0 && variant <= 3, "variant must be in [0, 3]");
    std::lock_guard<std::mutex> l(ctx->mu);
    ctx->pbs_variant = variant;
    return 0;
}

// ------------------------------------------------------------------------------------------
// level-synchronous programs
}  // extern "C"

struct b200tfhe_program {
    b200tfhe_ctx *ctx = nullptr;
    std::unique_ptr<Circuit> c;
    int32_t *d_term_block = nullptr, *d_outputs = nullptr;
    int64_t *d_term_coeff = nullptr;
    uint32_t *d_node_tbeg = nullptr, *d_node_lut = nullptr;
    uint64_t *d_node_pt = nullptr, *d_pool = nullptr, *d_stage = nullptr, *d_io = nullptr;
    size_t max_stage = 0;
};

namespace {

void program_free(b200tfhe_program *p) {
    if (!p) return;
    cudaFree(p->d_term_block); cudaFree(p->d_outputs); cudaFree(p->d_term_coeff); cudaFree(p->d_node_tbeg);
    cudaFree(p->d_node_lut); cudaFree(p->d_node_pt); cudaFree(p->d_pool); cudaFree(p->d_stage); cudaFree(p->d_io);
    delete p;
}

template <typename T>
int upload(b200tfhe_ctx *ctx, T **dst, const std::vector<T> &src) {
    const size_t bytes = std::max<size_t>(1, src.size()) * sizeof(T);
    CU_TRY(ctx, cudaMalloc(dst, bytes));
    if (!src.empty()) CU_TRY(ctx, cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// inputs must already be in the pool's first n_inputs slots
int program_execute(b200tfhe_program *p, uint64_t *d_out) {
    b200tfhe_ctx *ctx = p->ctx;
    const Circuit &c = *p->c;
    const int size = (int)ctx->big_size();
    const size_t n_in = c.n_inputs();
    for (const Circuit::Stage &st : c.stages) {
        const size_t n = st.end - st.begin;
        uint64_t *dst_nodes = p->d_pool + (n_in + st.begin) * (size_t)size;
        uint64_t *lin_out = st.bootstrap ? p->d_stage : dst_nodes;
        for (size_t off = 0; off < n; off += 65535) {   // gridDim.y limit
            const size_t cnt = std::min<size_t>(65535, n - off);
            dim3 grid((size + 255) / 256, (unsigned)cnt);
            lwe_lincomb_kernel<<<grid, 256, 0, ctx->stream>>>(p->d_pool, p->d_term_block, p->d_term_coeff, p->d_node_tbeg,
                                                              p->d_node_pt, lin_out + off * (size_t)size,
                                                              (int)(st.begin + off), size);
        }
        CU_TRY(ctx, cudaGetLastError());
        if (st.bootstrap) {
            if (int rc = ensure_workspace(ctx, n)) return rc;
            if (int rc = launch_ks(ctx, p->d_stage, ctx->d_small, n)) return rc;
            if (int rc = launch_pbs(ctx, ctx->d_small, p->d_node_lut + st.begin, dst_nodes, n)) return rc;
        }
    }
    const size_t n_out = c.outputs.size();
    for (size_t off = 0; off < n_out; off += 65535) {
        const size_t cnt = std::min<size_t>(65535, n_out - off);
        dim3 grid((size + 255) / 256, (unsigned)cnt);
        lwe_gather_kernel<<<grid, 256, 0, ctx->stream>>>(p->d_pool, p->d_outputs + off, d_out + off * (size_t)size, size);
    }
    CU_TRY(ctx, cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

int b200tfhe_program_create(b200tfhe_ctx *ctx, const char *op, const uint64_t *shape, size_t n_shape,
                            b200tfhe_program **out) {
    if (!out) return fail(ctx, "invalid argument: out is null");
    *out = nullptr;
    if (int rc = check_ready(ctx)) return rc;
    ARG_TRY(ctx, op && (shape || n_shape == 0), "null pointer");
    std::unique_ptr<Circuit> c;
    try {
        c = build_program(op, std::vector<uint64_t>(shape, shape + n_shape), ctx->p.message_modulus, ctx->p.carry_modulus);
    } catch (const std::exception &e) {
        return fail(ctx, std::string("program_create: ") + e.what());
    }
    auto *p = new b200tfhe_program();
    p->ctx = ctx;
    // lookup tables -> engine ids (content addressed, shared with every other user of the context)
    std::vector<uint32_t> lut_ids(c->luts.size());
    for (size_t l = 0; l < c->luts.size(); l++)
        if (b200tfhe_register_lut_from_table(ctx, c->luts[l].data(), c->luts[l].size(), &lut_ids[l])) { program_free(p); return 1; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    const uint64_t delta = ((uint64_t)1 << 63) / ((uint64_t)ctx->p.message_modulus * ctx->p.carry_modulus);
    std::vector<int32_t> tb(c->terms.size());
    std::vector<int64_t> tc(c->terms.size());
    for (size_t t = 0; t < c->terms.size(); t++) { tb[t] = c->terms[t].block; tc[t] = c->terms[t].coeff; }
    std::vector<uint32_t> tbeg(c->nodes.size() + 1), nlut(c->nodes.size());
    std::vector<uint64_t> npt(c->nodes.size());
    for (size_t k = 0; k < c->nodes.size(); k++) {
        tbeg[k] = c->nodes[k].term_begin;
        nlut[k] = c->nodes[k].lut >= 0 ? lut_ids[c->nodes[k].lut] : 0;
        npt[k] = c->nodes[k].plaintext * delta;
    }
    tbeg[c->nodes.size()] = (uint32_t)c->terms.size();
    for (const Circuit::Stage &st : c->stages)
        if (st.bootstrap) p->max_stage = std::max<size_t>(p->max_stage, st.end - st.begin);
    const size_t big = ctx->big_size();
    int rc = upload(ctx, &p->d_term_block, tb);
    if (!rc) rc = upload(ctx, &p->d_term_coeff, tc);
    if (!rc) rc = upload(ctx, &p->d_node_tbeg, tbeg);
    if (!rc) rc = upload(ctx, &p->d_node_lut, nlut);
    if (!rc) rc = upload(ctx, &p->d_node_pt, npt);
    if (!rc) rc = upload(ctx, &p->d_outputs, c->outputs);
    auto alloc = [&](uint64_t **ptr, size_t n_blocks) {
        if (rc) return;
        cudaError_t e = cudaMalloc(ptr, std::max<size_t>(1, n_blocks) * big * sizeof(uint64_t));
        if (e != cudaSuccess) rc = fail(ctx, std::string("cudaMalloc(program pool): ") + cudaGetErrorString(e));
    };
    alloc(&p->d_pool, c->n_blocks());
    alloc(&p->d_stage, p->max_stage);
    alloc(&p->d_io, c->outputs.size());
    if (rc) { program_free(p); return rc; }
    p->c = std::move(c);
    *out = p;
    return 0;
}

int b200tfhe_program_info(const b200tfhe_program *prog, uint64_t *info) {
    if (!prog || !info) return fail(nullptr, "invalid argument: null pointer");
    const Circuit &c = *prog->c;
    info[0] = c.n_inputs(); info[1] = c.outputs.size(); info[2] = c.n_pbs(); info[3] = c.depth();
    info[4] = c.stages.size(); info[5] = c.luts.size();
    return 0;
}

int b200tfhe_program_run_device(b200tfhe_program *prog, const uint64_t *d_in, uint64_t *d_out) {
    if (!prog) return fail(nullptr, "invalid argument: null program");
    b200tfhe_ctx *ctx = prog->ctx;
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    CU_TRY(ctx, cudaMemcpyAsync(prog->d_pool, d_in, prog->c->n_inputs() * ctx->big_size() * sizeof(uint64_t),
                                cudaMemcpyDeviceToDevice, ctx->stream));
    return program_execute(prog, d_out);
}

int b200tfhe_program_run(b200tfhe_program *prog, const uint64_t *in, uint64_t *out) {
    if (!prog) return fail(nullptr, "invalid argument: null program");
    b200tfhe_ctx *ctx = prog->ctx;
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, in && out, "null pointer");
    const size_t big_bytes = ctx->big_size() * sizeof(uint64_t);
    CU_TRY(ctx, cudaMemcpyAsync(prog->d_pool, in, prog->c->n_inputs() * big_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (int rc = program_execute(prog, prog->d_io)) return rc;
    CU_TRY(ctx, cudaMemcpyAsync(out, prog->d_io, prog->c->outputs.size() * big_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int b200tfhe_program_destroy(b200tfhe_program *prog) {
    if (!prog) return 0;
    cudaSetDevice(prog->ctx->device);
    cudaStreamSynchronize(prog->ctx->stream);
    program_free(prog);
    return 0;
}

int b200tfhe_debug_negacyclic_mul(b200tfhe_ctx *ctx, const uint64_t *a_int, const uint64_t *b_torus, uint64_t *out, size_t count) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (count == 0) return 0;
    ARG_TRY(ctx, a_int && b_torus && out, "null pointer");
    const size_t bytes = count * kN * sizeof(uint64_t);
    uint64_t *da = nullptr, *db = nullptr, *dout = nullptr;
    CU_TRY(ctx, cudaMalloc(&da, bytes));
    CU_TRY(ctx, cudaMalloc(&db, bytes));
    CU_TRY(ctx, cudaMalloc(&dout, bytes));
    cudaError_t e = cudaMemcpyAsync(da, a_int, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db, b_torus, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dout, out, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        negacyclic_mul_test_kernel<<<(unsigned)count, 32, 0, ctx->stream>>>(da, db, dout, ctx->d_twid, (int)count);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    CU_TRY(ctx, e);
    return 0;
}

}  // extern "C"
