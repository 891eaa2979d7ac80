// b200tfhe.cu -- C ABI of libb200tfhe.so (include/b200tfhe.h): context lifetime, key upload, LUT store, batched
// keyswitch / bootstrap entry points on host and device buffers, multi-GPU sharding, level-synchronous programs.
// Kernels: pbs_kernel5.cuh / pbs_kernel_lat.cuh (k = 1, N = 2048, l = 1), pbs_generic.cuh (every other classic
// parameter set), ks_mma.cuh (keyswitch on tcgen05), lwe_linear.cuh.  State and scheduler: context.hpp.
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

#include "context.hpp"
#include "ks_mma.cuh"
#include "lwe_linear.cuh"
#include "pbs_common.cuh"
#include "pbs_generic.cuh"
#include "pbs_kernel5.cuh"
#include "pbs_kernel_lat.cuh"
#include "pbs_kernel_lat4.cuh"
#include "programs.hpp"
#include "key_import.hpp"

using namespace b200;

namespace {

std::mutex g_err_mu;
std::string g_err;

void set_global_error(const std::string &s) {
    std::lock_guard<std::mutex> l(g_err_mu);
    g_err = s;
}

int fail(const b200tfhe_ctx *ctx, const std::string &msg) {
    if (ctx) {
        std::lock_guard<std::mutex> l(ctx->err_mu);
        ctx->err = msg;
    }
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
// generic path: roots[t] = exp(-2 pi i t / (N/2)) for t < N/4, twist[j] = exp(i pi j / N) for j < N/2 (fft/mod.rs:58-69)
void make_generic_tables(uint32_t N, std::vector<double2> &roots, std::vector<double2> &twist) {
    const long double pi = 3.14159265358979323846264338327950288L;
    roots.resize(N / 4);
    twist.resize(N / 2);
    for (uint32_t t = 0; t < N / 4; t++) {
        const long double a = -2.0L * pi * (long double)t / (long double)(N / 2);
        roots[t] = make_double2((double)cosl(a), (double)sinl(a));
    }
    for (uint32_t j = 0; j < N / 2; j++) {
        const long double a = pi * (long double)j / (long double)N;
        twist[j] = make_double2((double)cosl(a), (double)sinl(a));
    }
}

// ---- per-device one-time kernel configuration (attributes are per function and device, not per context) ------
constexpr int kMaxOptinSmem = 232448;   // 227 KB
std::mutex g_cfg_mu;
bool g_cfg_done[64] = {};

int configure_device_once(int device) {
    std::lock_guard<std::mutex> l(g_cfg_mu);
    if (device < 64 && g_cfg_done[device]) return 0;
    cudaError_t e = cudaSuccess;
    auto set = [&](const void *fn, int bytes) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    set((const void *)pbs_kernel5<3>, (int)pbs5_smem_bytes<3>());
    set((const void *)pbs_kernel5<4>, (int)pbs5_smem_bytes<4>());
    set((const void *)pbs_lat_kernel<2>, (int)pbs_lat_smem_bytes<2>());
    set((const void *)pbs_lat4_kernel, (int)pbs_lat4_smem_bytes());
    // parameter-independent maxima: two live contexts with different keyswitch levels share these functions
    set((const void *)ks_mma_kernel, kMaxOptinSmem);
    set((const void *)ks_digits_kernel, 64 * 1024);
    set((const void *)pbs_generic_kernel<uint64_t>, kMaxOptinSmem);
    set((const void *)bsk_to_fourier_generic_kernel<uint64_t>, kMaxOptinSmem);
    if (e != cudaSuccess) {
        set_global_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
        return 1;
    }
    if (device < 64) g_cfg_done[device] = true;
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

cudaEvent_t pool_event(DevCtx &d) {
    if (d.ev_next == d.ev_pool.size()) {
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        d.ev_pool.push_back(e);
    }
    return d.ev_pool[d.ev_next++];
}

int dev_fail(b200tfhe_ctx *ctx, DevCtx &d, const std::string &msg) {
    d.err = msg;
    return fail(ctx, "device " + std::to_string(d.device) + ": " + msg);
}
#define DEV_TRY(ctx, d, expr)                                                                    \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) return dev_fail(ctx, d, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

void free_workspace(DevCtx &d) {
    cudaFree(d.d_in); cudaFree(d.d_small); cudaFree(d.d_out); cudaFree(d.d_lut_idx); cudaFree(d.d_digits);
    cudaFree(d.d_acc_ws); cudaFree(d.d_fourier_ws);
    d.d_in = d.d_small = d.d_out = nullptr; d.d_lut_idx = nullptr; d.d_digits = nullptr;
    d.d_acc_ws = nullptr; d.d_fourier_ws = nullptr; d.ws_cap = 0;
}

// Workspaces grow geometrically (so a ragged sequence of level sizes reallocates O(log) times) and only after the
// work already queued on the device has drained.
int ensure_workspace(b200tfhe_ctx *ctx, DevCtx &d, size_t batch) {
    if (batch <= d.ws_cap) return 0;
    size_t cap = std::max<size_t>(256, d.ws_cap * 2);
    while (cap < batch) cap *= 2;
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.s_h2d));
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.s_d2h));
    free_workspace(d);
    DEV_TRY(ctx, d, cudaMalloc(&d.d_in, cap * ctx->big_size() * sizeof(uint64_t)));
    DEV_TRY(ctx, d, cudaMalloc(&d.d_small, cap * ctx->small_size() * sizeof(uint64_t)));
    DEV_TRY(ctx, d, cudaMalloc(&d.d_out, cap * ctx->big_size() * sizeof(uint64_t)));
    DEV_TRY(ctx, d, cudaMalloc(&d.d_lut_idx, cap * sizeof(uint32_t)));
    if (ctx->ks_tensor) DEV_TRY(ctx, d, cudaMalloc(&d.d_digits, ctx->ks_geom.a_total_bytes(cap)));
    if (!ctx->fast_path) {
        DEV_TRY(ctx, d, cudaMalloc(&d.d_acc_ws, cap * ctx->glwe_len() * sizeof(uint64_t)));
        DEV_TRY(ctx, d, cudaMalloc(&d.d_fourier_ws, cap * ctx->fourier_per_ct() * sizeof(double2)));
    }
    d.ws_cap = cap;
    return 0;
}

int ensure_lut_capacity(b200tfhe_ctx *ctx, DevCtx &d, size_t n) {
    if (n <= d.lut_cap) return 0;
    size_t cap = std::max<size_t>(64, d.lut_cap * 2);
    while (cap < n) cap *= 2;
    uint64_t *nd = nullptr;
    DEV_TRY(ctx, d, cudaMalloc(&nd, cap * ctx->glwe_len() * sizeof(uint64_t)));
    if (d.d_luts) {
        DEV_TRY(ctx, d, cudaMemcpyAsync(nd, d.d_luts, d.lut_count * ctx->glwe_len() * sizeof(uint64_t), cudaMemcpyDeviceToDevice, d.stream));
        DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
        cudaFree(d.d_luts);
    }
    d.d_luts = nd;
    d.lut_cap = cap;
    return 0;
}

int ensure_staging(b200tfhe_ctx *ctx, DevCtx &d, size_t slab_bytes) {
    if (slab_bytes <= d.slab_bytes) return 0;
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.s_h2d));
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.s_d2h));
    for (int s = 0; s < kStageSlabs; s++) {
        cudaFreeHost(d.h_in[s]); cudaFreeHost(d.h_out[s]);
        d.h_in[s] = d.h_out[s] = nullptr;
    }
    d.slab_bytes = 0;
    for (int s = 0; s < kStageSlabs; s++) {
        DEV_TRY(ctx, d, cudaHostAlloc((void **)&d.h_in[s], slab_bytes, cudaHostAllocDefault));
        DEV_TRY(ctx, d, cudaHostAlloc((void **)&d.h_out[s], slab_bytes, cudaHostAllocDefault));
    }
    d.slab_bytes = slab_bytes;
    return 0;
}

bool is_pinned_host(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

void prof_begin(DevCtx &d, std::vector<EventPair> &v) {
    if (!d.profiling) return;
    EventPair e{};
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, d.stream);
    v.push_back(e);
}
void prof_end(DevCtx &d, std::vector<EventPair> &v) {
    if (!d.profiling) return;
    cudaEventRecord(v.back().b, d.stream);
}

// ---- kernel launches (device d must be current) -------------------------------------------------------------
int launch_ks(b200tfhe_ctx *ctx, DevCtx &d, const uint64_t *d_in, uint64_t *d_out, size_t batch) {
    if (!ctx->ksk_loaded) return fail(ctx, "keyswitch key not loaded");
    if (int rc = ensure_workspace(ctx, d, batch)) return rc;
    prof_begin(d, d.ev_ks);
    if (ctx->ks_tensor) {
        // digits pre-pass + s8 x u8 tcgen05 GEMM over the KSK byte limbs (ks_mma.cuh)
        const KsMmaGeom &g = ctx->ks_geom;
        const unsigned m_tiles = (unsigned)((batch + kKmM - 1) / kKmM);
        ks_digits_kernel<<<dim3(g.k_stages, m_tiles), 256, g.a_stage_bytes(), d.stream>>>(d_in, d.d_digits, g, (int)batch);
        KsMmaArgs m{};
        m.a_tiled = d.d_digits; m.b_tiled = ctx->d_ksk_limbs(d); m.in = d_in; m.out = d_out; m.g = g;
        m.batch = (int)batch; m.stages = ks_mma_pipeline_stages(g);
        ks_mma_kernel<<<dim3(m_tiles, g.n_tiles), 128, ks_mma_smem_bytes(g), d.stream>>>(m);
        d.kernel_launches += 2;
    } else {
        ks_generic_kernel<uint64_t><<<(unsigned)batch, 256, 0, d.stream>>>(d_in, ctx->d_ksk(d), d_out, (int)batch,
                                                                          (int)(ctx->p.glwe_dimension * ctx->p.polynomial_size),
                                                                          (int)ctx->small_size(), (int)ctx->p.ks_base_log, (int)ctx->p.ks_level);
        d.kernel_launches += 1;
    }
    prof_end(d, d.ev_ks);
    DEV_TRY(ctx, d, cudaGetLastError());
    d.ks_launches++;
    return 0;
}

// Ciphertexts per CTA of the specialised kernels (one CTA per SM).  One launch with c ciphertexts per CTA takes a whole
// number of waves of t(c) each; measured on this pool's B200 per wave: 3.25 ms for 1 per SM (pbs_lat4_kernel: four warps per
// polynomial), 4.7 ms for 2 (pbs_lat_kernel<2>: two warps per polynomial), 6.1 ms for 3 and 6.47 ms for 4 (pbs_kernel5).
// A batch is served either by ONE launch with the fewest ciphertexts per CTA that still fits it into the minimum number of
// waves, or by a launch of full 4-per-SM waves followed by a second launch for the remainder with the kernel that suits the
// remainder (700 = 592 + 108: 6.47 + 3.25 ms instead of two waves of 3 per SM = 12.2 ms) -- whichever the wave model says is
// shorter.
struct PbsWaveModel { double t1 = 3.25, t2 = 4.7, t3 = 6.1, t4 = 6.47, split_penalty = 0.15; };
inline int pbs_per_cta(long long batch, long long sms) {
    const long long waves = (batch + 4 * sms - 1) / (4 * sms);
    return (int)((batch + waves * sms - 1) / (waves * sms));
}
void launch_pbs_one(DevCtx &d, const PbsArgs &a, int per_cta) {
    switch (per_cta) {
#ifdef B200TFHE_LAB_LAT2   // development A/B switch: the two-warps-per-polynomial kernel for one ciphertext per SM
        case 1: cudaFuncSetAttribute(pbs_lat_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pbs_lat_smem_bytes<1, true>());
                pbs_lat_kernel<1, true><<<(unsigned)a.batch, 128, pbs_lat_smem_bytes<1, true>(), d.stream>>>(a); break;
#else
        case 1: pbs_lat4_kernel<<<(unsigned)a.batch, 256, pbs_lat4_smem_bytes(), d.stream>>>(a); break;
#endif
        case 2: pbs_lat_kernel<2><<<(unsigned)((a.batch + 1) / 2), 256, pbs_lat_smem_bytes<2>(), d.stream>>>(a); break;
        case 3: pbs_kernel5<3><<<(unsigned)((a.batch + 2) / 3), 192, pbs5_smem_bytes<3>(), d.stream>>>(a); break;
        default: pbs_kernel5<4><<<(unsigned)((a.batch + 3) / 4), 256, pbs5_smem_bytes<4>(), d.stream>>>(a); break;
    }
}
void launch_pbs_fast(DevCtx &d, const PbsArgs &a) {
    const PbsWaveModel m;
    const long long sms = d.sm_count, full = a.batch / (4 * sms), rest = a.batch - full * 4 * sms;
    const double t[5] = {0, m.t1, m.t2, m.t3, m.t4};
    const int c_one = pbs_per_cta(a.batch, sms);
    const double one = (double)((a.batch + c_one * sms - 1) / (c_one * sms)) * t[c_one];
    if (full > 0 && rest > 0) {
        const int c_rest = pbs_per_cta(rest, sms);
        if ((double)full * m.t4 + t[c_rest] + m.split_penalty < one) {
            PbsArgs head = a, tail = a;
            head.batch = (int)(full * 4 * sms);
            tail.batch = (int)rest;
            tail.lwe_small = a.lwe_small + (size_t)head.batch * (a.n + 1);
            tail.lut_idx = a.lut_idx ? a.lut_idx + head.batch : nullptr;
            tail.out = a.out + (size_t)head.batch * (kN + 1);
            launch_pbs_one(d, head, 4);
            launch_pbs_one(d, tail, c_rest);
            d.kernel_launches++;   // (the caller counts one)
            return;
        }
    }
    launch_pbs_one(d, a, c_one);
}

int launch_pbs(b200tfhe_ctx *ctx, DevCtx &d, const uint64_t *d_small, const uint32_t *d_lut_idx, uint64_t *d_out, size_t batch) {
    if (!ctx->bsk_loaded) return fail(ctx, "bootstrap key not loaded");
    if (d.lut_count == 0) return fail(ctx, "no lookup table registered");
    prof_begin(d, d.ev_pbs);
    if (ctx->fast_path) {
        PbsArgs a{};
        a.lwe_small = d_small; a.lut_idx = d_lut_idx; a.luts = d.d_luts; a.bsk = ctx->d_bsk(d); a.twid = d.d_twid;
        a.out = d_out; a.batch = (int)batch; a.n = (int)ctx->p.lwe_dimension;
        a.n_luts = (uint32_t)d.lut_count; a.err_flag = d.d_err_flag;
        launch_pbs_fast(d, a);
    } else {
        if (int rc = ensure_workspace(ctx, d, batch)) return rc;
        GenPbsArgs g{};
        g.lwe_small = d_small; g.lut_idx = d_lut_idx; g.luts = d.d_luts; g.bsk = ctx->d_bsk(d);
        g.roots = d.d_roots; g.twist = d.d_twist; g.acc_ws = d.d_acc_ws; g.fourier_ws = d.d_fourier_ws; g.out = d_out;
        g.batch = (int)batch; g.n = (int)ctx->p.lwe_dimension; g.k = (int)ctx->p.glwe_dimension; g.log2N = ctx->log2N;
        g.base_log = (int)ctx->p.pbs_base_log; g.level = (int)ctx->p.pbs_level; g.fft_in_smem = ctx->fft_in_smem ? 1 : 0;
        g.n_luts = (uint32_t)d.lut_count; g.err_flag = d.d_err_flag;
        const GenLaunch gl = gen_launch_shape(ctx->p.polynomial_size, ctx->p.glwe_dimension, ctx->p.pbs_level, ctx->fft_in_smem, kMaxOptinSmem);
        g.polys_in_smem = gl.polys_in_smem;
        const size_t smem = gl.smem;
        const unsigned threads = gl.threads;
        pbs_generic_kernel<uint64_t><<<(unsigned)batch, threads, smem, d.stream>>>(g);
    }
    prof_end(d, d.ev_pbs);
    DEV_TRY(ctx, d, cudaGetLastError());
    d.pbs_launches++;
    d.kernel_launches++;
    return 0;
}

int check_ready(b200tfhe_ctx *ctx) {
    if (!ctx) return fail(nullptr, "null context");
    CU_TRY(ctx, cudaSetDevice(ctx->devs[0]->device));
    return 0;
}

int validate_lut_ids(b200tfhe_ctx *ctx, const uint32_t *lut_id, size_t batch) {
    if (!lut_id) return 0;
    const uint32_t n = (uint32_t)ctx->h_luts.size();
    for (size_t b = 0; b < batch; b++)
        if (lut_id[b] >= n) return fail(ctx, "invalid argument: lut_id out of range");
    return 0;
}

// Runs fn(device index) for every GPU of the context: GPU 0 on the calling thread, the others on their workers.
template <typename F>
int for_each_device(b200tfhe_ctx *ctx, F fn) {
    const int n = (int)ctx->devs.size();
    for (int i = 1; i < n; i++) ctx->devs[i]->worker->post([ctx, fn, i]() -> int {
        if (cudaSetDevice(ctx->devs[i]->device) != cudaSuccess) return 1;
        return fn(i);
    });
    int rc = 0;
    if (cudaSetDevice(ctx->devs[0]->device) != cudaSuccess) rc = fail(ctx, "cudaSetDevice failed");
    if (!rc) rc = fn(0);
    for (int i = 1; i < n; i++) {
        const int r = ctx->devs[i]->worker->wait();
        if (r && !rc) rc = r;
    }
    return rc;
}

// ---- one GPU's share of a host-buffer KS+PBS batch ----------------------------------------------------------
// Pipelined over chunks of one full wave (4 ciphertexts per SM): the H2D copy of chunk c+1 and the D2H copy of
// chunk c-1 run on their own streams under the kernels of chunk c.  Pinned host buffers are copied in place;
// pageable ones (a Rust Vec<u64>, a numpy array) go through the context's pinned slabs: the host thread memcpy's
// chunk c+2 into a slab while chunk c computes, so only the first upload and the last download are exposed.
int run_ks_pbs_host(b200tfhe_ctx *ctx, DevCtx &d, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch) {
    if (batch == 0) return 0;
    if (int rc = ensure_workspace(ctx, d, batch)) return rc;
    const size_t big = ctx->big_size(), small = ctx->small_size();
    const size_t chunk = (size_t)d.sm_count * 4;
    const size_t n_chunks = (batch + chunk - 1) / chunk;
    const bool pin_in = is_pinned_host(in), pin_out = is_pinned_host(out);
    if (!pin_in || !pin_out)
        if (int rc = ensure_staging(ctx, d, std::min(chunk, batch) * big * sizeof(uint64_t))) return rc;
    d.ev_next = 0;
    cudaEvent_t ev_start = pool_event(d);   // work already queued on the compute stream (earlier async calls) goes first
    DEV_TRY(ctx, d, cudaEventRecord(ev_start, d.stream));
    DEV_TRY(ctx, d, cudaStreamWaitEvent(d.s_h2d, ev_start, 0));
    if (lut_id) DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_lut_idx, lut_id, batch * sizeof(uint32_t), cudaMemcpyHostToDevice, d.s_h2d));
    std::vector<cudaEvent_t> ev_in(n_chunks), ev_done(n_chunks), ev_out(n_chunks);
    for (size_t c = 0; c < n_chunks; c++) { ev_in[c] = pool_event(d); ev_done[c] = pool_event(d); ev_out[c] = pool_event(d); }
    auto drain = [&](size_t c) -> int {    // staged download of chunk c: wait for its D2H, then copy slab -> caller's buffer
        const size_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
        DEV_TRY(ctx, d, cudaEventSynchronize(ev_out[c]));
        std::memcpy(out + b0 * big, d.h_out[c % kStageSlabs], nb * big * sizeof(uint64_t));
        return 0;
    };
    for (size_t c = 0; c < n_chunks; c++) {
        const size_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
        const size_t bytes = nb * big * sizeof(uint64_t);
        const uint64_t *src = in + b0 * big;
        if (!pin_in) {
            if (c >= kStageSlabs) DEV_TRY(ctx, d, cudaEventSynchronize(ev_in[c - kStageSlabs]));   // slab free again
            std::memcpy(d.h_in[c % kStageSlabs], src, bytes);
            src = reinterpret_cast<const uint64_t *>(d.h_in[c % kStageSlabs]);
        }
        DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_in + b0 * big, src, bytes, cudaMemcpyHostToDevice, d.s_h2d));
        DEV_TRY(ctx, d, cudaEventRecord(ev_in[c], d.s_h2d));
        DEV_TRY(ctx, d, cudaStreamWaitEvent(d.stream, ev_in[c], 0));
        if (int rc = launch_ks(ctx, d, d.d_in + b0 * big, d.d_small + b0 * small, nb)) return rc;
        if (int rc = launch_pbs(ctx, d, d.d_small + b0 * small, lut_id ? d.d_lut_idx + b0 : nullptr, d.d_out + b0 * big, nb)) return rc;
        DEV_TRY(ctx, d, cudaEventRecord(ev_done[c], d.stream));
        DEV_TRY(ctx, d, cudaStreamWaitEvent(d.s_d2h, ev_done[c], 0));
        if (!pin_out && c >= kStageSlabs)
            if (int rc = drain(c - kStageSlabs)) return rc;                                       // frees slab c % kStageSlabs
        uint64_t *dst = pin_out ? out + b0 * big : reinterpret_cast<uint64_t *>(d.h_out[c % kStageSlabs]);
        DEV_TRY(ctx, d, cudaMemcpyAsync(dst, d.d_out + b0 * big, bytes, cudaMemcpyDeviceToHost, d.s_d2h));
        DEV_TRY(ctx, d, cudaEventRecord(ev_out[c], d.s_d2h));
    }
    if (!pin_out)
        for (size_t c = n_chunks > kStageSlabs ? n_chunks - kStageSlabs : 0; c < n_chunks; c++)
            if (int rc = drain(c)) return rc;
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.s_h2d));
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.s_d2h));
    return 0;
}

int check_device_errors(b200tfhe_ctx *ctx, DevCtx &d) {
    uint32_t flag = 0;
    DEV_TRY(ctx, d, cudaMemcpyAsync(&flag, d.d_err_flag, sizeof(flag), cudaMemcpyDeviceToHost, d.stream));
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
    if (flag) {
        cudaMemsetAsync(d.d_err_flag, 0, sizeof(uint32_t), d.stream);
        return dev_fail(ctx, d, "a device-side lut id was out of range (table 0 was used for it)");
    }
    return 0;
}

void destroy_dev(DevCtx &d) {
    cudaSetDevice(d.device);
    if (d.stream) cudaStreamSynchronize(d.stream);
    d.worker.reset();
    for (auto &e : d.ev_ks) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto &e : d.ev_pbs) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto &e : d.ev_pool) cudaEventDestroy(e);
    cudaFree(d.arena); cudaFree(d.d_twid); cudaFree(d.d_roots); cudaFree(d.d_twist); cudaFree(d.d_luts); cudaFree(d.d_err_flag);
    free_workspace(d);
    for (int s = 0; s < kStageSlabs; s++) { cudaFreeHost(d.h_in[s]); cudaFreeHost(d.h_out[s]); }
    if (d.stream) cudaStreamDestroy(d.stream);
    if (d.s_h2d) cudaStreamDestroy(d.s_h2d);
    if (d.s_d2h) cudaStreamDestroy(d.s_d2h);
}

void destroy_ctx(b200tfhe_ctx *ctx) {
    for (auto &d : ctx->devs)
        if (d) destroy_dev(*d);
    delete ctx;
}

// copies the key arena of GPU 0 to every other GPU of the context (NVLink peer copy when available)
int replicate_arena(b200tfhe_ctx *ctx, size_t off, size_t bytes) {
    DevCtx &d0 = *ctx->devs[0];
    for (size_t i = 1; i < ctx->devs.size(); i++) {
        DevCtx &di = *ctx->devs[i];
        DEV_TRY(ctx, d0, cudaMemcpyPeerAsync(di.arena + off, di.device, d0.arena + off, d0.device, bytes, d0.stream));
    }
    DEV_TRY(ctx, d0, cudaStreamSynchronize(d0.stream));
    return 0;
}

}  // namespace

extern "C" {

int b200tfhe_last_global_error(char *buf, size_t buf_len) {
    std::lock_guard<std::mutex> l(g_err_mu);
    if (buf && buf_len) std::snprintf(buf, buf_len, "%s", g_err.c_str());
    return 0;
}

int b200tfhe_last_error(const b200tfhe_ctx *ctx, char *buf, size_t buf_len) {
    if (!ctx) return b200tfhe_last_global_error(buf, buf_len);
    std::lock_guard<std::mutex> l(ctx->err_mu);
    if (buf && buf_len) std::snprintf(buf, buf_len, "%s", ctx->err.c_str());
    return 0;
}

int b200tfhe_ctx_create_multi(const b200tfhe_params *params, const int *devices, int n_devices, b200tfhe_ctx **out) {
    if (!out) return fail(nullptr, "invalid argument: out is null");
    *out = nullptr;  // c_api/shortint/server_key/pbs.rs:54-58 nulls results on entry too
    if (!params) return fail(nullptr, "invalid argument: params is null");
    if (!devices || n_devices < 1 || n_devices > 64) return fail(nullptr, "invalid argument: need 1..64 devices");
    const b200tfhe_params &p = *params;
    // ---- pure parameter checks first: nothing is allocated when they fail
    uint32_t log2N = 0;
    while ((1u << log2N) < p.polynomial_size) log2N++;
    if (p.polynomial_size < 256 || p.polynomial_size > 32768 || (1u << log2N) != p.polynomial_size)
        return fail(nullptr, "unsupported parameters: polynomial_size must be a power of two in [256, 32768]");
    if (p.glwe_dimension == 0 || p.glwe_dimension > 8)
        return fail(nullptr, "unsupported parameters: glwe_dimension must be in [1, 8]");
    if (p.pbs_level == 0 || p.pbs_base_log == 0 || p.pbs_base_log * p.pbs_level >= 64)
        return fail(nullptr, "unsupported parameters: pbs_base_log*pbs_level must be in [1, 63]");
    if (p.lwe_dimension == 0 || p.lwe_dimension > 4096)
        return fail(nullptr, "unsupported parameters: lwe_dimension must be in [1, 4096]");
    if (p.ks_base_log == 0 || p.ks_level == 0 || p.ks_base_log * p.ks_level >= 64)
        return fail(nullptr, "unsupported parameters: ks_base_log*ks_level must be in [1, 63]");
    if (p.message_modulus == 0 || p.carry_modulus == 0 || p.polynomial_size % (p.message_modulus * p.carry_modulus) != 0)
        return fail(nullptr, "unsupported parameters: message_modulus*carry_modulus must divide polynomial_size");
    const bool fast = p.glwe_dimension == 1 && p.polynomial_size == 2048 && p.pbs_level == 1 && p.pbs_base_log == 23 &&
                      p.lwe_dimension <= (uint32_t)kMaxSmallDim;
    const KsMmaGeom geom = ks_mma_geom((int)(p.glwe_dimension * p.polynomial_size), (int)p.lwe_dimension + 1, (int)p.ks_level, (int)p.ks_base_log);
    const bool ks_tensor = p.ks_base_log <= 7 && p.ks_level <= (uint32_t)kKmMaxLevel && ks_mma_pipeline_stages(geom) >= 2;
    if (!ks_tensor && (p.lwe_dimension + 1 > 1024 || p.ks_level > 8))
        return fail(nullptr, "unsupported parameters: keyswitch decomposition outside both keyswitch kernels");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                 (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
    for (int i = 0; i < n_devices; i++) {
        if (devices[i] < 0 || devices[i] >= n_dev) return fail(nullptr, "invalid argument: device index out of range");
        for (int j = 0; j < i; j++)
            if (devices[j] == devices[i]) return fail(nullptr, "invalid argument: duplicate device index");
        cudaDeviceProp prop{};
        if (cudaGetDeviceProperties(&prop, devices[i]) != cudaSuccess) return fail(nullptr, "cudaGetDeviceProperties failed");
        if (prop.major != 10) return fail(nullptr, "unsupported GPU: libb200tfhe is built for sm_100a (B200) only");
    }

    auto *ctx = new b200tfhe_ctx();
    ctx->p = p;
    ctx->fast_path = fast;
    ctx->ks_tensor = ks_tensor;
    ctx->log2N = (int)log2N;
    ctx->fft_in_smem = (size_t)p.polynomial_size * 3 / 4 * sizeof(double2) <= 200 * 1024;   // N/2 points + N/4 roots: N <= 16384
    ctx->ks_geom = geom;
    ctx->off_bsk = 0;
    ctx->off_ksk = align_up(ctx->bsk_len() / 2 * sizeof(double2), 256);  // N/2 complex per polynomial
    ctx->off_ksk_limbs = ctx->off_ksk + align_up(ctx->ksk_len() * sizeof(uint64_t), 256);
    ctx->arena_bytes = ctx->off_ksk_limbs + (ks_tensor ? align_up(geom.b_total_bytes(), 256) : 0);
    auto bail = [&](const std::string &m) {
        fail(nullptr, m);
        destroy_ctx(ctx);   // frees whatever the devices created so far
        return 1;
    };
    std::vector<double2> tw, roots, twist;
    if (fast) make_twiddles(tw);
    else make_generic_tables(p.polynomial_size, roots, twist);
    for (int i = 0; i < n_devices; i++) {
        ctx->devs.emplace_back(new DevCtx());
        DevCtx &d = *ctx->devs.back();
        d.device = devices[i];
        if (cudaSetDevice(d.device) != cudaSuccess) return bail("cudaSetDevice failed");
        if (configure_device_once(d.device)) { destroy_ctx(ctx); return 1; }
        cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.device);
        if (cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&d.s_d2h, cudaStreamNonBlocking) != cudaSuccess)
            return bail("cudaStreamCreate failed");
        if (cudaMalloc(&d.arena, ctx->arena_bytes) != cudaSuccess) return bail("cudaMalloc(key arena) failed");
        if (cudaMalloc(&d.d_err_flag, sizeof(uint32_t)) != cudaSuccess || cudaMemset(d.d_err_flag, 0, sizeof(uint32_t)) != cudaSuccess)
            return bail("cudaMalloc(error flag) failed");
        auto up = [&](double2 **dst, const std::vector<double2> &src) {
            return cudaMalloc(dst, src.size() * sizeof(double2)) == cudaSuccess &&
                   cudaMemcpy(*dst, src.data(), src.size() * sizeof(double2), cudaMemcpyHostToDevice) == cudaSuccess;
        };
        if (fast ? !up(&d.d_twid, tw) : !(up(&d.d_roots, roots) && up(&d.d_twist, twist))) return bail("cudaMalloc(twiddles) failed");
        if (i > 0) {
            d.worker.reset(new Worker());
            // peer access lets cudaMemcpyPeerAsync go GPU to GPU over NVLink; without it the copy is staged by the driver
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctx->devs[0]->device, d.device) == cudaSuccess && can) {
                cudaSetDevice(ctx->devs[0]->device);
                if (cudaDeviceEnablePeerAccess(d.device, 0) != cudaSuccess) cudaGetLastError();
                cudaSetDevice(d.device);
            }
        }
    }
    cudaSetDevice(ctx->devs[0]->device);
    *out = ctx;
    return 0;
}

int b200tfhe_ctx_create(const b200tfhe_params *params, int device, b200tfhe_ctx **out) {
    return b200tfhe_ctx_create_multi(params, &device, 1, out);
}

int b200tfhe_ctx_device_count(const b200tfhe_ctx *ctx, int *n_devices) {
    if (!ctx || !n_devices) return fail(nullptr, "invalid argument: null pointer");
    *n_devices = (int)ctx->devs.size();
    return 0;
}

int b200tfhe_ctx_destroy(b200tfhe_ctx *ctx) {
    if (!ctx) return 0;
    destroy_ctx(ctx);
    return 0;
}

int b200tfhe_load_ksk(b200tfhe_ctx *ctx, const uint64_t *ksk, size_t n_u64) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, ksk != nullptr, "ksk is null");
    ARG_TRY(ctx, n_u64 == ctx->ksk_len(), "ksk length does not match parameters (k*N*ks_level*(n+1))");
    DevCtx &d = *ctx->devs[0];
    DEV_TRY(ctx, d, cudaMemcpyAsync(ctx->d_ksk(d), ksk, n_u64 * sizeof(uint64_t), cudaMemcpyHostToDevice, d.stream));
    if (ctx->ks_tensor) {
        ksk_limbs_kernel<<<dim3(ctx->ks_geom.k_stages, ctx->ks_geom.n_tiles), kKmN, 0, d.stream>>>(ctx->d_ksk(d), ctx->d_ksk_limbs(d), ctx->ks_geom);
        DEV_TRY(ctx, d, cudaGetLastError());
    }
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
    if (int rc = replicate_arena(ctx, ctx->off_ksk, ctx->arena_bytes - ctx->off_ksk)) return rc;
    ctx->ksk_loaded = true;
    return 0;
}

int b200tfhe_load_bsk_standard(b200tfhe_ctx *ctx, const uint64_t *bsk, size_t n_u64) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, bsk != nullptr, "bsk is null");
    ARG_TRY(ctx, n_u64 == ctx->bsk_len(), "bsk length does not match parameters (n*pbs_level*(k+1)^2*N)");
    DevCtx &d = *ctx->devs[0];
    uint64_t *tmp = nullptr;
    double2 *scratch = nullptr;
    DEV_TRY(ctx, d, cudaMalloc(&tmp, n_u64 * sizeof(uint64_t)));
    cudaError_t e = cudaMemcpyAsync(tmp, bsk, n_u64 * sizeof(uint64_t), cudaMemcpyHostToDevice, d.stream);
    const uint32_t N = ctx->p.polynomial_size;
    const size_t n_polys = n_u64 / N;
    if (e == cudaSuccess) {
        if (ctx->fast_path) {
            bsk_to_fourier_kernel<<<std::min<int>((int)n_polys / 2 + 1, 148 * 8), 64, 0, d.stream>>>(tmp, ctx->d_bsk(d), d.d_twid, (int)n_polys);
        } else {
            if (!ctx->fft_in_smem) e = cudaMalloc(&scratch, n_polys * (N / 2) * sizeof(double2));
            if (e == cudaSuccess)
                bsk_to_fourier_generic_kernel<uint64_t><<<(unsigned)n_polys, 512, ctx->fft_in_smem ? (size_t)N / 2 * sizeof(double2) : 0, d.stream>>>(
                    tmp, ctx->d_bsk(d), d.d_roots, d.d_twist, scratch, ctx->log2N, ctx->fft_in_smem ? 1 : 0);
        }
        if (e == cudaSuccess) e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(tmp);
    cudaFree(scratch);
    DEV_TRY(ctx, d, e);
    if (int rc = replicate_arena(ctx, ctx->off_bsk, ctx->off_ksk)) return rc;
    ctx->bsk_loaded = true;
    return 0;
}

int b200tfhe_key_arena(b200tfhe_ctx *ctx, void **device_ptr, size_t *bytes) {
    if (int rc = check_ready(ctx)) return rc;
    ARG_TRY(ctx, device_ptr && bytes, "null output pointer");
    *device_ptr = ctx->devs[0]->arena;
    *bytes = ctx->arena_bytes;
    return 0;
}

int b200tfhe_keys_adopt(b200tfhe_ctx *ctx) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (int rc = replicate_arena(ctx, 0, ctx->arena_bytes)) return rc;
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
    const uint32_t nid = (uint32_t)ctx->h_luts.size();
    ctx->h_luts.emplace_back(glwe_acc, glwe_acc + len);
    // every GPU gets the table; copied from our own stable host copy so the caller's buffer can be reused immediately
    for (auto &dp : ctx->devs) {
        DevCtx &d = *dp;
        cudaSetDevice(d.device);
        int rc = ensure_lut_capacity(ctx, d, (size_t)nid + 1);
        if (!rc) {
            cudaError_t e = cudaMemcpyAsync(d.d_luts + (size_t)nid * len, ctx->h_luts.back().data(), len * sizeof(uint64_t), cudaMemcpyHostToDevice, d.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
            if (e != cudaSuccess) rc = dev_fail(ctx, d, std::string("register_lut: ") + cudaGetErrorString(e));
        }
        if (rc) {
            ctx->h_luts.pop_back();
            cudaSetDevice(ctx->devs[0]->device);
            return rc;
        }
        d.lut_count = (size_t)nid + 1;
    }
    cudaSetDevice(ctx->devs[0]->device);
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

// ---- device-buffer entry points: first GPU of the context, asynchronous on its stream -----------------------
int b200tfhe_keyswitch_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, uint64_t *d_out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    return launch_ks(ctx, *ctx->devs[0], d_in, d_out, batch);
}

int b200tfhe_pbs_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    return launch_pbs(ctx, *ctx->devs[0], d_in, d_lut_id, d_out, batch);
}

int b200tfhe_ks_pbs_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    DevCtx &d = *ctx->devs[0];
    if (int rc = ensure_workspace(ctx, d, batch)) return rc;
    if (int rc = launch_ks(ctx, d, d_in, d.d_small, batch)) return rc;
    return launch_pbs(ctx, d, d.d_small, d_lut_id, d_out, batch);
}

int b200tfhe_pbs_ks_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    DevCtx &d = *ctx->devs[0];
    if (int rc = ensure_workspace(ctx, d, batch)) return rc;
    if (int rc = launch_pbs(ctx, d, d_in, d_lut_id, d.d_out, batch)) return rc;
    return launch_ks(ctx, d, d.d_out, d_out, batch);
}

// One shard per GPU, every pointer resident on that GPU; asynchronous (b200tfhe_sync waits for all GPUs).
int b200tfhe_ks_pbs_batch_device_multi(b200tfhe_ctx *ctx, const uint64_t *const *d_in, const uint32_t *const *d_lut_id,
                                       uint64_t *const *d_out, const size_t *batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, d_in && d_out && batch, "null pointer");
    int rc = 0;
    for (size_t i = 0; i < ctx->devs.size() && !rc; i++) {
        if (batch[i] == 0) continue;
        DevCtx &d = *ctx->devs[i];
        ARG_TRY(ctx, d_in[i] && d_out[i], "null pointer");
        cudaSetDevice(d.device);
        rc = ensure_workspace(ctx, d, batch[i]);
        if (!rc) rc = launch_ks(ctx, d, d_in[i], d.d_small, batch[i]);
        if (!rc) rc = launch_pbs(ctx, d, d.d_small, d_lut_id ? d_lut_id[i] : nullptr, d_out[i], batch[i]);
    }
    cudaSetDevice(ctx->devs[0]->device);
    return rc;
}

// ---- host-buffer entry points: the batch is cut into one contiguous shard per GPU ---------------------------
int b200tfhe_keyswitch_batch(b200tfhe_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in && out, "null pointer");
    const size_t big = ctx->big_size(), small = ctx->small_size();
    const int world = (int)ctx->devs.size();
    return for_each_device(ctx, [=](int i) -> int {
        DevCtx &d = *ctx->devs[i];
        size_t b0, b1;
        shard_bounds(batch, world, i, &b0, &b1);
        if (b1 == b0) return 0;
        const size_t nb = b1 - b0;
        if (int rc = ensure_workspace(ctx, d, nb)) return rc;
        DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_in, in + b0 * big, nb * big * sizeof(uint64_t), cudaMemcpyHostToDevice, d.stream));
        if (int rc = launch_ks(ctx, d, d.d_in, d.d_small, nb)) return rc;
        DEV_TRY(ctx, d, cudaMemcpyAsync(out + b0 * small, d.d_small, nb * small * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.stream));
        DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
        return 0;
    });
}

int b200tfhe_pbs_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in && out, "null pointer");
    if (int rc = validate_lut_ids(ctx, lut_id, batch)) return rc;
    const size_t big = ctx->big_size(), small = ctx->small_size();
    const int world = (int)ctx->devs.size();
    return for_each_device(ctx, [=](int i) -> int {
        DevCtx &d = *ctx->devs[i];
        size_t b0, b1;
        shard_bounds(batch, world, i, &b0, &b1);
        if (b1 == b0) return 0;
        const size_t nb = b1 - b0;
        if (int rc = ensure_workspace(ctx, d, nb)) return rc;
        DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_small, in + b0 * small, nb * small * sizeof(uint64_t), cudaMemcpyHostToDevice, d.stream));
        if (lut_id) DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_lut_idx, lut_id + b0, nb * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
        if (int rc = launch_pbs(ctx, d, d.d_small, lut_id ? d.d_lut_idx : nullptr, d.d_out, nb)) return rc;
        DEV_TRY(ctx, d, cudaMemcpyAsync(out + b0 * big, d.d_out, nb * big * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.stream));
        DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
        return 0;
    });
}

int b200tfhe_ks_pbs_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in && out, "null pointer");
    if (int rc = validate_lut_ids(ctx, lut_id, batch)) return rc;
    const size_t big = ctx->big_size();
    const int world = (int)ctx->devs.size();
    return for_each_device(ctx, [=](int i) -> int {
        size_t b0, b1;
        shard_bounds(batch, world, i, &b0, &b1);
        return run_ks_pbs_host(ctx, *ctx->devs[i], in + b0 * big, lut_id ? lut_id + b0 : nullptr, out + b0 * big, b1 - b0);
    });
}

int b200tfhe_pbs_ks_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in && out, "null pointer");
    if (int rc = validate_lut_ids(ctx, lut_id, batch)) return rc;
    const size_t small = ctx->small_size();
    const int world = (int)ctx->devs.size();
    return for_each_device(ctx, [=](int i) -> int {
        DevCtx &d = *ctx->devs[i];
        size_t b0, b1;
        shard_bounds(batch, world, i, &b0, &b1);
        if (b1 == b0) return 0;
        const size_t nb = b1 - b0, bytes = nb * small * sizeof(uint64_t);
        if (int rc = ensure_workspace(ctx, d, nb)) return rc;
        DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_small, in + b0 * small, bytes, cudaMemcpyHostToDevice, d.stream));
        if (lut_id) DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_lut_idx, lut_id + b0, nb * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
        if (int rc = launch_pbs(ctx, d, d.d_small, lut_id ? d.d_lut_idx : nullptr, d.d_out, nb)) return rc;
        if (int rc = launch_ks(ctx, d, d.d_out, d.d_small, nb)) return rc;
        DEV_TRY(ctx, d, cudaMemcpyAsync(out + b0 * small, d.d_small, bytes, cudaMemcpyDeviceToHost, d.stream));
        DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
        return 0;
    });
}

int b200tfhe_lwe_linear_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_x, const uint64_t *d_y, const int32_t *d_ia,
                                     const int32_t *d_ib, const int64_t *d_ca, const int64_t *d_cb,
                                     const uint64_t *d_pt, uint64_t *d_out, size_t batch, size_t lwe_size) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, d_x && d_ca && d_out, "null pointer");
    ARG_TRY(ctx, batch <= 65535, "batch too large for one launch (max 65535)");
    DevCtx &d = *ctx->devs[0];
    LinArgs a{};
    a.x = d_x; a.y = d_y; a.ia = d_ia; a.ib = d_ib; a.ca = d_ca; a.cb = d_cb; a.pt = d_pt; a.out = d_out;
    a.batch = (int)batch; a.size = (int)lwe_size;
    dim3 grid((unsigned)((lwe_size + 255) / 256), (unsigned)batch);
    lwe_linear_kernel<<<grid, 256, 0, d.stream>>>(a);
    DEV_TRY(ctx, d, cudaGetLastError());
    d.kernel_launches++;
    return 0;
}

int b200tfhe_sync(b200tfhe_ctx *ctx) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    int rc = 0;
    for (auto &dp : ctx->devs) {
        cudaSetDevice(dp->device);
        const int r = check_device_errors(ctx, *dp);   // synchronises the stream
        if (r && !rc) rc = r;
    }
    cudaSetDevice(ctx->devs[0]->device);
    return rc;
}

int b200tfhe_stream(b200tfhe_ctx *ctx, void **stream) {
    if (!ctx) return fail(nullptr, "null context");
    ARG_TRY(ctx, stream != nullptr, "null pointer");
    *stream = (void *)ctx->devs[0]->stream;
    return 0;
}

int b200tfhe_set_profiling(b200tfhe_ctx *ctx, int enabled) {
    if (!ctx) return fail(nullptr, "null context");
    std::lock_guard<std::mutex> l(ctx->mu);
    for (auto &dp : ctx->devs) dp->profiling = enabled != 0;
    return 0;
}

// Device time (ms) and launch counts of the first GPU since the last reset (every GPU of a context runs the same
// schedule on its shard); synchronises.
int b200tfhe_get_kernel_times(b200tfhe_ctx *ctx, double *ks_ms, uint64_t *ks_launches, double *pbs_ms,
                              uint64_t *pbs_launches, int reset) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    for (size_t i = 0; i < ctx->devs.size(); i++) {
        DevCtx &d = *ctx->devs[i];
        cudaSetDevice(d.device);
        DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
        auto drain = [&](std::vector<EventPair> &v, double &acc) {
            for (auto &e : v) {
                float ms = 0;
                if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) acc += ms;
                cudaEventDestroy(e.a);
                cudaEventDestroy(e.b);
            }
            v.clear();
        };
        drain(d.ev_ks, d.ks_ms);
        drain(d.ev_pbs, d.pbs_ms);
    }
    cudaSetDevice(ctx->devs[0]->device);
    DevCtx &d0 = *ctx->devs[0];
    if (ks_ms) *ks_ms = d0.ks_ms;
    if (pbs_ms) *pbs_ms = d0.pbs_ms;
    if (ks_launches) *ks_launches = d0.ks_launches;
    if (pbs_launches) *pbs_launches = d0.pbs_launches;
    if (reset)
        for (auto &dp : ctx->devs) {
            dp->ks_ms = dp->pbs_ms = 0;
            dp->ks_launches = dp->pbs_launches = 0;
        }
    return 0;
}

// Number of kernels this library has launched on all GPUs of the context since creation (bench.py's gpu_launches).
int b200tfhe_kernel_launch_count(b200tfhe_ctx *ctx, uint64_t *count) {
    if (!ctx || !count) return fail(ctx, "invalid argument: null pointer");
    std::lock_guard<std::mutex> l(ctx->mu);
    uint64_t c = 0;
    for (auto &dp : ctx->devs) c += dp->kernel_launches;
    *count = c;
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// level-synchronous programs
struct ProgramPart {   // one GPU's share of a program: its own circuit (for `units` of the batch) and device tables
    int dev = 0;
    size_t unit_begin = 0, units = 0;
    std::unique_ptr<Circuit> c;
    int32_t *d_term_block = nullptr, *d_outputs = nullptr;
    int64_t *d_term_coeff = nullptr;
    uint32_t *d_node_tbeg = nullptr, *d_node_lut = nullptr;
    uint64_t *d_node_pt = nullptr, *d_pool = nullptr, *d_stage = nullptr, *d_io = nullptr;
    size_t max_stage = 0;
};

struct b200tfhe_program {
    b200tfhe_ctx *ctx = nullptr;
    ProgramLayout layout;
    std::vector<ProgramPart> parts;
    size_t n_inputs = 0, n_outputs = 0, n_pbs = 0, depth = 0, n_stages = 0, n_luts = 0;
};

namespace {

void program_free(b200tfhe_program *p) {
    if (!p) return;
    for (ProgramPart &q : p->parts) {
        cudaSetDevice(p->ctx->devs[q.dev]->device);
        cudaFree(q.d_term_block); cudaFree(q.d_outputs); cudaFree(q.d_term_coeff); cudaFree(q.d_node_tbeg);
        cudaFree(q.d_node_lut); cudaFree(q.d_node_pt); cudaFree(q.d_pool); cudaFree(q.d_stage); cudaFree(q.d_io);
    }
    cudaSetDevice(p->ctx->devs[0]->device);
    delete p;
}

template <typename T>
int upload(b200tfhe_ctx *ctx, T **dst, const std::vector<T> &src) {
    const size_t bytes = std::max<size_t>(1, src.size()) * sizeof(T);
    CU_TRY(ctx, cudaMalloc(dst, bytes));
    if (!src.empty()) CU_TRY(ctx, cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// device tables of one part (its GPU must be current); lut ids are the context-wide registered ids
int part_upload(b200tfhe_ctx *ctx, ProgramPart &q, const std::vector<uint32_t> &lut_ids) {
    const Circuit &c = *q.c;
    const uint64_t delta = ((uint64_t)1 << 63) / ((uint64_t)ctx->p.message_modulus * ctx->p.carry_modulus);
    std::vector<int32_t> tb(c.terms.size());
    std::vector<int64_t> tc(c.terms.size());
    for (size_t t = 0; t < c.terms.size(); t++) { tb[t] = c.terms[t].block; tc[t] = c.terms[t].coeff; }
    std::vector<uint32_t> tbeg(c.nodes.size() + 1), nlut(c.nodes.size());
    std::vector<uint64_t> npt(c.nodes.size());
    for (size_t k = 0; k < c.nodes.size(); k++) {
        tbeg[k] = c.nodes[k].term_begin;
        nlut[k] = c.nodes[k].lut >= 0 ? lut_ids[c.nodes[k].lut] : 0;
        npt[k] = c.nodes[k].plaintext * delta + c.nodes[k].plaintext_half * (delta / 2);
    }
    tbeg[c.nodes.size()] = (uint32_t)c.terms.size();
    for (const Circuit::Stage &st : c.stages)
        if (st.bootstrap) q.max_stage = std::max<size_t>(q.max_stage, st.end - st.begin);
    const size_t big = ctx->big_size();
    int rc = upload(ctx, &q.d_term_block, tb);
    if (!rc) rc = upload(ctx, &q.d_term_coeff, tc);
    if (!rc) rc = upload(ctx, &q.d_node_tbeg, tbeg);
    if (!rc) rc = upload(ctx, &q.d_node_lut, nlut);
    if (!rc) rc = upload(ctx, &q.d_node_pt, npt);
    if (!rc) rc = upload(ctx, &q.d_outputs, c.outputs);
    auto alloc = [&](uint64_t **ptr, size_t n_blocks) {
        if (rc) return;
        cudaError_t e = cudaMalloc(ptr, std::max<size_t>(1, n_blocks) * big * sizeof(uint64_t));
        if (e != cudaSuccess) rc = fail(ctx, std::string("cudaMalloc(program pool): ") + cudaGetErrorString(e));
    };
    alloc(&q.d_pool, c.n_blocks());
    alloc(&q.d_stage, q.max_stage);
    alloc(&q.d_io, c.outputs.size());
    return rc;
}

// inputs must already be in the pool's first n_inputs slots; runs every level on the part's GPU (current device)
int part_execute(b200tfhe_ctx *ctx, ProgramPart &q, uint64_t *d_out) {
    DevCtx &d = *ctx->devs[q.dev];
    const Circuit &c = *q.c;
    const int size = (int)ctx->big_size();
    const size_t n_in = c.n_inputs();
    for (const Circuit::Stage &st : c.stages) {
        const size_t n = st.end - st.begin;
        uint64_t *dst_nodes = q.d_pool + (n_in + st.begin) * (size_t)size;
        uint64_t *lin_out = st.bootstrap ? q.d_stage : dst_nodes;
        for (size_t off = 0; off < n; off += 65535) {   // gridDim.y limit
            const size_t cnt = std::min<size_t>(65535, n - off);
            dim3 grid((size + 255) / 256, (unsigned)cnt);
            lwe_lincomb_kernel<<<grid, 256, 0, d.stream>>>(q.d_pool, q.d_term_block, q.d_term_coeff, q.d_node_tbeg,
                                                           q.d_node_pt, lin_out + off * (size_t)size, (int)(st.begin + off), size);
            d.kernel_launches++;
        }
        DEV_TRY(ctx, d, cudaGetLastError());
        if (st.bootstrap) {
            if (int rc = ensure_workspace(ctx, d, n)) return rc;
            if (int rc = launch_ks(ctx, d, q.d_stage, d.d_small, n)) return rc;
            if (int rc = launch_pbs(ctx, d, d.d_small, q.d_node_lut + st.begin, dst_nodes, n)) return rc;
        }
    }
    const size_t n_out = c.outputs.size();
    for (size_t off = 0; off < n_out; off += 65535) {
        const size_t cnt = std::min<size_t>(65535, n_out - off);
        dim3 grid((size + 255) / 256, (unsigned)cnt);
        lwe_gather_kernel<<<grid, 256, 0, d.stream>>>(q.d_pool, q.d_outputs + off, d_out + off * (size_t)size, size);
        d.kernel_launches++;
    }
    DEV_TRY(ctx, d, cudaGetLastError());
    return 0;
}

// registers the circuit's function tables with the context and finishes the program object
int program_finish(b200tfhe_ctx *ctx, b200tfhe_program *p) {
    for (ProgramPart &q : p->parts) {
        const Circuit &c = *q.c;
        std::vector<uint32_t> lut_ids(c.luts.size());
        for (size_t l = 0; l < c.luts.size(); l++) {
            // whole-unit tables go through fill_accumulator on the library side; half-unit tables (the 16-input reductions) are
            // built here with entries scaled by delta / 2
            std::vector<uint64_t> acc(ctx->glwe_len(), 0);
            const std::vector<uint64_t> body = c.lut_body((int)l, ctx->p.polynomial_size);
            std::copy(body.begin(), body.end(), acc.begin() + (size_t)ctx->p.glwe_dimension * ctx->p.polynomial_size);
            if (b200tfhe_register_lut(ctx, acc.data(), &lut_ids[l])) return 1;
        }
        std::lock_guard<std::mutex> lk(ctx->mu);
        if (cudaSetDevice(ctx->devs[q.dev]->device) != cudaSuccess) return fail(ctx, "cudaSetDevice failed");
        const int rc = part_upload(ctx, q, lut_ids);
        cudaSetDevice(ctx->devs[0]->device);
        if (rc) return rc;
        p->n_inputs += c.n_inputs(); p->n_outputs += c.outputs.size(); p->n_pbs += c.n_pbs();
        p->depth = std::max<size_t>(p->depth, c.depth()); p->n_stages = std::max<size_t>(p->n_stages, c.stages.size());
        p->n_luts = std::max<size_t>(p->n_luts, c.luts.size());
    }
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
    const std::vector<uint64_t> sh(shape, shape + n_shape);
    auto *p = new b200tfhe_program();
    p->ctx = ctx;
    try {
        p->layout = program_layout(op, sh);
        // independent units (integers, strings) are dealt contiguously over the GPUs; anything else runs on the first
        const int world = p->layout.splittable ? (int)std::min<size_t>(ctx->devs.size(), std::max<size_t>(1, p->layout.units)) : 1;
        for (int i = 0; i < world; i++) {
            ProgramPart q;
            q.dev = i;
            std::vector<uint64_t> s = sh;
            if (p->layout.splittable) {
                size_t u0, u1;
                shard_bounds(p->layout.units, world, i, &u0, &u1);
                q.unit_begin = u0; q.units = u1 - u0;
                s[0] = q.units;
            } else {
                q.units = 1;
            }
            q.c = build_program(op, s, ctx->p.message_modulus, ctx->p.carry_modulus);
            p->parts.push_back(std::move(q));
        }
    } catch (const std::exception &e) {
        program_free(p);
        return fail(ctx, std::string("program_create: ") + e.what());
    }
    if (program_finish(ctx, p)) { program_free(p); return 1; }
    *out = p;
    return 0;
}

// A caller-built schedule (what the reference's integer layer would hand over instead of one apply_lookup_table call
// per block: integer/server_key/radix_parallel/comparison.rs:22-28, add.rs:529-535).  Blocks 0..n_inputs-1 are the
// inputs; node j produces block n_inputs + j = LUT_j( sum_t coeff[t] * block[t] + plaintext_j * delta ), or just the
// linear combination when node_lut[j] < 0.  Nodes may only reference earlier blocks; the library levels the DAG and
// runs one lwe-linear launch + one KS+PBS launch per dependency level.  luts: n_luts tables of
// message_modulus*carry_modulus entries (function values, as for b200tfhe_register_lut_from_table).
int b200tfhe_program_create_from_circuit(b200tfhe_ctx *ctx, const b200tfhe_circuit_desc *desc, b200tfhe_program **out) {
    if (!out) return fail(ctx, "invalid argument: out is null");
    *out = nullptr;
    if (int rc = check_ready(ctx)) return rc;
    ARG_TRY(ctx, desc != nullptr, "null circuit description");
    ARG_TRY(ctx, desc->n_nodes == 0 || (desc->node_term_begin && desc->node_lut), "null node arrays");
    ARG_TRY(ctx, desc->n_outputs == 0 || desc->outputs, "null outputs");
    auto *p = new b200tfhe_program();
    p->ctx = ctx;
    try {
        const size_t ms = (size_t)ctx->p.message_modulus * ctx->p.carry_modulus;
        ProgramPart q;
        q.dev = 0; q.units = 1;
        q.c.reset(new Circuit(ctx->p.message_modulus, ctx->p.carry_modulus, desc->n_inputs));
        Circuit &c = *q.c;
        for (size_t l = 0; l < desc->n_luts; l++) {
            const uint64_t *t = desc->luts + l * ms;
            c.lut([t](uint64_t x) { return t[x]; });
        }
        if (c.luts.size() != desc->n_luts) throw std::invalid_argument("duplicate lookup tables in the description");
        std::vector<Lin> blocks;
        blocks.reserve(desc->n_inputs + desc->n_nodes);
        for (size_t i = 0; i < desc->n_inputs; i++) blocks.push_back(c.input(i, (uint32_t)ms - 1));
        for (size_t j = 0; j < desc->n_nodes; j++) {
            Lin x;
            for (uint32_t t = desc->node_term_begin[j]; t < desc->node_term_begin[j + 1]; t++) {
                const int32_t b = desc->term_block[t];
                if (b < 0 || (size_t)b >= blocks.size()) throw std::out_of_range("node references a later block");
                x = c.axpy(x, 1, blocks[b], desc->term_coeff[t]);
            }
            if (desc->node_plaintext) x = c.add_const(x, (int64_t)(desc->node_plaintext[j] % (2 * ms)));
            x.degree = 0;   // the caller owns the degree bookkeeping for a custom schedule
            const int32_t lut = desc->node_lut[j];
            if (lut >= (int32_t)desc->n_luts) throw std::out_of_range("node lut index");
            blocks.push_back(lut >= 0 ? c.pbs_unchecked(x, lut) : c.materialize_always(x));
        }
        for (size_t o = 0; o < desc->n_outputs; o++) {
            const int32_t b = desc->outputs[o];
            if (b < 0 || (size_t)b >= blocks.size()) throw std::out_of_range("output block index");
            c.output(blocks[b]);
        }
        c.finalize();
        p->parts.push_back(std::move(q));
    } catch (const std::exception &e) {
        program_free(p);
        return fail(ctx, std::string("program_create_from_circuit: ") + e.what());
    }
    if (program_finish(ctx, p)) { program_free(p); return 1; }
    *out = p;
    return 0;
}

int b200tfhe_program_info(const b200tfhe_program *prog, uint64_t *info) {
    if (!prog || !info) return fail(nullptr, "invalid argument: null pointer");
    info[0] = prog->n_inputs; info[1] = prog->n_outputs; info[2] = prog->n_pbs; info[3] = prog->depth;
    info[4] = prog->n_stages; info[5] = prog->n_luts;
    return 0;
}

int b200tfhe_program_run_device(b200tfhe_program *prog, const uint64_t *d_in, uint64_t *d_out) {
    if (!prog) return fail(nullptr, "invalid argument: null program");
    b200tfhe_ctx *ctx = prog->ctx;
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, d_in && d_out, "null pointer");
    ARG_TRY(ctx, prog->parts.size() == 1, "device-buffer runs need a single-GPU program (create the context on one device)");
    ProgramPart &q = prog->parts[0];
    DevCtx &d = *ctx->devs[0];
    DEV_TRY(ctx, d, cudaMemcpyAsync(q.d_pool, d_in, q.c->n_inputs() * ctx->big_size() * sizeof(uint64_t), cudaMemcpyDeviceToDevice, d.stream));
    return part_execute(ctx, q, d_out);
}

int b200tfhe_program_run(b200tfhe_program *prog, const uint64_t *in, uint64_t *out) {
    if (!prog) return fail(nullptr, "invalid argument: null program");
    b200tfhe_ctx *ctx = prog->ctx;
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    ARG_TRY(ctx, in && out, "null pointer");
    const size_t big = ctx->big_size(), big_bytes = big * sizeof(uint64_t);
    const ProgramLayout &lay = prog->layout;
    auto run_part = [=](ProgramPart &q) -> int {
        DevCtx &d = *ctx->devs[q.dev];
        if (prog->parts.size() == 1) {
            DEV_TRY(ctx, d, cudaMemcpyAsync(q.d_pool, in, q.c->n_inputs() * big_bytes, cudaMemcpyHostToDevice, d.stream));
        } else {
            // this part's units of every input array, packed in the order the sub-program expects
            size_t seg_off = 0, dst = 0;
            for (size_t s = 0; s < lay.in_seg.size(); s++) {
                const size_t per = lay.in_seg[s];
                DEV_TRY(ctx, d, cudaMemcpyAsync(q.d_pool + dst * big, in + (seg_off + q.unit_begin * per) * big, q.units * per * big_bytes,
                                                cudaMemcpyHostToDevice, d.stream));
                dst += q.units * per;
                seg_off += lay.units * per;
            }
        }
        if (int rc = part_execute(ctx, q, q.d_io)) return rc;
        uint64_t *o = out + (prog->parts.size() == 1 ? 0 : q.unit_begin * lay.out_per_unit) * big;
        DEV_TRY(ctx, d, cudaMemcpyAsync(o, q.d_io, q.c->outputs.size() * big_bytes, cudaMemcpyDeviceToHost, d.stream));
        DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
        return 0;
    };
    if (prog->parts.size() == 1) return run_part(prog->parts[0]);
    return for_each_device(ctx, [=](int i) -> int {
        if ((size_t)i >= prog->parts.size()) return 0;
        return run_part(prog->parts[i]);
    });
}

int b200tfhe_program_destroy(b200tfhe_program *prog) {
    if (!prog) return 0;
    for (auto &dp : prog->ctx->devs) {
        cudaSetDevice(dp->device);
        cudaStreamSynchronize(dp->stream);
    }
    program_free(prog);
    return 0;
}

// ---- key wire format import (key_import.hpp) -------------------------------------------------------------
int b200tfhe_parse_server_key(const uint8_t *bytes, size_t n_bytes, b200tfhe_params *params, b200tfhe_key_view *view) {
    if (!bytes || !params || !view) return fail(nullptr, "invalid argument: null pointer");
    std::string err;
    if (!parse_shortint_server_key(bytes, n_bytes, params, view, &err)) return fail(nullptr, "parse_server_key: " + err);
    return 0;
}

int b200tfhe_load_server_key_bytes(b200tfhe_ctx *ctx, const uint8_t *bytes, size_t n_bytes) {
    if (!ctx) return fail(nullptr, "null context");
    ARG_TRY(ctx, bytes != nullptr, "null pointer");
    b200tfhe_params p{};
    b200tfhe_key_view v{};
    std::string err;
    if (!parse_shortint_server_key(bytes, n_bytes, &p, &v, &err)) return fail(ctx, "load_server_key_bytes: " + err);
    if (std::memcmp(&p, &ctx->p, 7 * sizeof(uint32_t)) != 0)   // n, k, N, PBS and KS decompositions
        return fail(ctx, "load_server_key_bytes: the key's dimensions / decompositions differ from the context's parameters");
    // serde stores u64 little endian and unaligned: copy into aligned buffers
    std::vector<uint64_t> ksk(v.ksk_len);
    std::memcpy(ksk.data(), bytes + v.ksk_offset, v.ksk_len * sizeof(uint64_t));
    if (int rc = b200tfhe_load_ksk(ctx, ksk.data(), ksk.size())) return rc;
    if (!v.bsk_is_fourier) {
        std::vector<uint64_t> bsk(v.bsk_len);
        std::memcpy(bsk.data(), bytes + v.bsk_offset, v.bsk_len * sizeof(uint64_t));
        return b200tfhe_load_bsk_standard(ctx, bsk.data(), bsk.size());
    }
    // Fourier key of a serialised shortint::ServerKey.  ASSUMPTION (unpinned, concrete-fft 0.3.0 is not in the tree):
    // Plan::serialize_fourier_buffer writes the N/2 coefficients in natural frequency order k of
    // sum_j z_j exp(-2 pi i j k / (N/2)), z_j the twisted folded input of fft/mod.rs:197-218.  That is the order of the
    // specialised kernels' Fourier key ([q][lane] with k = lane + 32 q) and the bit reversal of the generic kernel's;
    // the 2/N of the inverse transform is folded into the key in both.
    const size_t half = ctx->p.polynomial_size / 2, n_polys = v.bsk_len / half;
    if (n_polys * half * 2 != ctx->bsk_len()) return fail(ctx, "load_server_key_bytes: Fourier key size does not match the parameters");
    std::vector<double2> f(n_polys * half);
    const double norm = 1.0 / (double)half;
    for (size_t poly = 0; poly < n_polys; poly++) {
        const uint8_t *src = bytes + v.bsk_offset + poly * v.bsk_poly_stride_bytes;
        for (size_t k = 0; k < half; k++) {
            double c[2];
            std::memcpy(c, src + k * 16, 16);
            size_t dst = k;
            if (!ctx->fast_path) {   // generic kernel: decimation-in-frequency output order = bit-reversed k
                size_t rv = 0;
                for (int b = 0; b < ctx->log2N - 1; b++) rv |= ((k >> b) & 1) << (ctx->log2N - 2 - b);
                dst = rv;
            }
            f[poly * half + dst] = make_double2(c[0] * norm, c[1] * norm);
        }
    }
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    DevCtx &d = *ctx->devs[0];
    DEV_TRY(ctx, d, cudaMemcpyAsync(ctx->d_bsk(d), f.data(), f.size() * sizeof(double2), cudaMemcpyHostToDevice, d.stream));
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
    if (int rc = replicate_arena(ctx, ctx->off_bsk, ctx->off_ksk)) return rc;
    ctx->bsk_loaded = true;
    return 0;
}

int b200tfhe_debug_negacyclic_mul(b200tfhe_ctx *ctx, const uint64_t *a_int, const uint64_t *b_torus, uint64_t *out, size_t count) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (count == 0) return 0;
    ARG_TRY(ctx, a_int && b_torus && out, "null pointer");
    ARG_TRY(ctx, ctx->fast_path, "the FFT product test hook exercises the N = 2048 warp transform only");
    DevCtx &d = *ctx->devs[0];
    const size_t bytes = count * kN * sizeof(uint64_t);
    uint64_t *da = nullptr, *db = nullptr, *dout = nullptr;
    CU_TRY(ctx, cudaMalloc(&da, bytes));
    CU_TRY(ctx, cudaMalloc(&db, bytes));
    CU_TRY(ctx, cudaMalloc(&dout, bytes));
    cudaError_t e = cudaMemcpyAsync(da, a_int, bytes, cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db, b_torus, bytes, cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dout, out, bytes, cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) {
        negacyclic_mul_test_kernel<<<(unsigned)count, 32, 0, d.stream>>>(da, db, dout, d.d_twid, (int)count);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(da); cudaFree(db); cudaFree(dout);
    CU_TRY(ctx, e);
    return 0;
}

int b200tfhe_debug_from_torus(b200tfhe_ctx *ctx, const double *x, uint64_t *out_fp, uint64_t *out_cvt, size_t n) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (n == 0) return 0;
    ARG_TRY(ctx, x && out_fp && out_cvt, "null pointer");
    DevCtx &d = *ctx->devs[0];
    double *dx = nullptr; uint64_t *da = nullptr, *db = nullptr;
    CU_TRY(ctx, cudaMalloc(&dx, n * sizeof(double)));
    CU_TRY(ctx, cudaMalloc(&da, n * sizeof(uint64_t)));
    CU_TRY(ctx, cudaMalloc(&db, n * sizeof(uint64_t)));
    cudaError_t e = cudaMemcpyAsync(dx, x, n * sizeof(double), cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) {
        from_torus_test_kernel<<<(unsigned)((n + 255) / 256), 256, 0, d.stream>>>(dx, da, db, n);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_fp, da, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_cvt, db, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
    cudaFree(dx); cudaFree(da); cudaFree(db);
    CU_TRY(ctx, e);
    return 0;
}

// One external product on caller-supplied data (unit-test hook for the Fourier stage): runs the production PBS kernel
// selected for `batch` on the first `steps` mask elements only (lwe_dimension is taken as `steps`).
int b200tfhe_debug_pbs_steps(b200tfhe_ctx *ctx, const uint64_t *in_small, const uint32_t *lut_id, uint64_t *out, size_t batch, uint32_t steps) {
    if (int rc = check_ready(ctx)) return rc;
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    ARG_TRY(ctx, in_small && out, "null pointer");
    ARG_TRY(ctx, ctx->fast_path && steps <= ctx->p.lwe_dimension, "only for the specialised kernels, steps <= lwe_dimension");
    if (int rc = validate_lut_ids(ctx, lut_id, batch)) return rc;
    DevCtx &d = *ctx->devs[0];
    if (int rc = ensure_workspace(ctx, d, batch)) return rc;
    // inputs: batch x (steps + 1) words (mask prefix, body)
    DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_small, in_small, batch * (steps + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, d.stream));
    if (lut_id) DEV_TRY(ctx, d, cudaMemcpyAsync(d.d_lut_idx, lut_id, batch * sizeof(uint32_t), cudaMemcpyHostToDevice, d.stream));
    PbsArgs a{};
    a.lwe_small = d.d_small; a.lut_idx = lut_id ? d.d_lut_idx : nullptr; a.luts = d.d_luts; a.bsk = ctx->d_bsk(d); a.twid = d.d_twid;
    a.out = d.d_out; a.batch = (int)batch; a.n = (int)steps; a.n_luts = (uint32_t)d.lut_count; a.err_flag = d.d_err_flag;
    launch_pbs_fast(d, a);
    DEV_TRY(ctx, d, cudaGetLastError());
    d.kernel_launches++;
    DEV_TRY(ctx, d, cudaMemcpyAsync(out, d.d_out, batch * ctx->big_size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, d.stream));
    DEV_TRY(ctx, d, cudaStreamSynchronize(d.stream));
    return 0;
}

}  // extern "C"

#include "boolean_api.hpp"
