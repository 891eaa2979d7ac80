// boolean_api.hpp -- the reference's u32 boolean gate path on the generic kernels of pbs_generic.cuh
// instantiated for 32-bit torus words.  A gate is: a linear pre-combination of the two inputs (boolean/engine/mod.rs:
// 606-850: AND a+b-1/8, NAND -(a+b)+1/8, OR a+b+1/8, NOR -(a+b)-1/8, XOR 2(a+b+1/8), XNOR -2(a+b+1/8)), then the
// bootstrapping pattern of boolean/engine/bootstrapping.rs:257-391 with the constant test polynomial 1/8
// (PLAINTEXT_TRUE, :60): bootstrap then keyswitch for EncryptionKeyChoice::Small, keyswitch then bootstrap for ::Big.
// Included by b200tfhe.cu (one translation unit).
#pragma once

struct b200tfhe_boolean_ctx {
    b200tfhe_params p{};
    int ks_first = 0;         // EncryptionKeyChoice::Big: ciphertexts under the GLWE-derived key, keyswitch first
    int device = 0, log2N = 9;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::string err;
    uint32_t *d_ksk = nullptr, *d_lut = nullptr, *d_a = nullptr, *d_b = nullptr, *d_mid = nullptr, *d_mid2 = nullptr, *d_out = nullptr, *d_acc = nullptr;
    double2 *d_bsk = nullptr, *d_fourier = nullptr, *d_roots = nullptr, *d_twist = nullptr;
    size_t cap = 0;
    bool ksk_loaded = false, bsk_loaded = false;
    size_t big() const { return (size_t)p.glwe_dimension * p.polynomial_size + 1; }
    size_t small() const { return (size_t)p.lwe_dimension + 1; }
    size_t ct_size() const { return ks_first ? big() : small(); }
    size_t ksk_len() const { return (size_t)p.glwe_dimension * p.polynomial_size * p.ks_level * small(); }
    size_t bsk_len() const { return (size_t)p.lwe_dimension * p.pbs_level * (p.glwe_dimension + 1) * (p.glwe_dimension + 1) * p.polynomial_size; }
};

namespace {

int bfail(b200tfhe_boolean_ctx *ctx, const std::string &m) {
    if (ctx) ctx->err = m;
    set_global_error(m);
    return 1;
}
#define BCU_TRY(ctx, expr)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) return bfail(ctx, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// out = coeff * (a + b), body += cst   (wrapping u32; lwe_linear_algebra.rs:68,276,556)
__global__ void boolean_prelin_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t *__restrict__ out,
                                      const uint32_t coeff, const uint32_t cst, const int size, const size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t v = coeff * (a[i] + b[i]);
        if ((int)(i % size) == size - 1) v += cst;
        out[i] = v;
    }
}

void boolean_free(b200tfhe_boolean_ctx *c) {
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->d_ksk); cudaFree(c->d_lut); cudaFree(c->d_a); cudaFree(c->d_b); cudaFree(c->d_mid); cudaFree(c->d_mid2); cudaFree(c->d_out);
    cudaFree(c->d_acc); cudaFree(c->d_bsk); cudaFree(c->d_fourier); cudaFree(c->d_roots); cudaFree(c->d_twist);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int boolean_workspace(b200tfhe_boolean_ctx *c, size_t batch) {
    if (batch <= c->cap) return 0;
    size_t cap = std::max<size_t>(64, c->cap * 2);
    while (cap < batch) cap *= 2;
    BCU_TRY(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_a); cudaFree(c->d_b); cudaFree(c->d_mid); cudaFree(c->d_mid2); cudaFree(c->d_out); cudaFree(c->d_acc); cudaFree(c->d_fourier);
    c->d_a = c->d_b = c->d_mid = c->d_mid2 = c->d_out = c->d_acc = nullptr; c->d_fourier = nullptr; c->cap = 0;
    const size_t ct = std::max(c->big(), c->small()) * sizeof(uint32_t);
    BCU_TRY(c, cudaMalloc(&c->d_a, cap * ct)); BCU_TRY(c, cudaMalloc(&c->d_b, cap * ct));
    BCU_TRY(c, cudaMalloc(&c->d_mid, cap * ct)); BCU_TRY(c, cudaMalloc(&c->d_mid2, cap * ct)); BCU_TRY(c, cudaMalloc(&c->d_out, cap * ct));
    BCU_TRY(c, cudaMalloc(&c->d_acc, cap * (size_t)(c->p.glwe_dimension + 1) * c->p.polynomial_size * sizeof(uint32_t)));
    BCU_TRY(c, cudaMalloc(&c->d_fourier, cap * (size_t)(c->p.glwe_dimension + 1) * (c->p.polynomial_size / 2) * sizeof(double2)));
    c->cap = cap;
    return 0;
}

}  // namespace

extern "C" {

int b200tfhe_boolean_ctx_create(const b200tfhe_params *params, int keyswitch_first, int device, b200tfhe_boolean_ctx **out) {
    if (!out) return fail(nullptr, "invalid argument: out is null");
    *out = nullptr;
    if (!params) return fail(nullptr, "invalid argument: params is null");
    const b200tfhe_params &p = *params;
    uint32_t log2N = 0;
    while ((1u << log2N) < p.polynomial_size) log2N++;
    if (p.polynomial_size < 256 || p.polynomial_size > 4096 || (1u << log2N) != p.polynomial_size)
        return fail(nullptr, "unsupported boolean parameters: polynomial_size must be a power of two in [256, 4096]");
    if (p.glwe_dimension == 0 || p.glwe_dimension > 4 || p.lwe_dimension == 0 || p.lwe_dimension + 1 > 1024)
        return fail(nullptr, "unsupported boolean parameters: glwe_dimension in [1, 4], lwe_dimension < 1024");
    if (p.pbs_level == 0 || p.pbs_base_log == 0 || p.pbs_base_log * p.pbs_level >= 32 || p.ks_level == 0 || p.ks_level > 8 ||
        p.ks_base_log == 0 || p.ks_base_log * p.ks_level >= 32)
        return fail(nullptr, "unsupported boolean parameters: decompositions must fit 32-bit words (ks_level <= 8)");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(nullptr, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                 (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)));
    if (device < 0 || device >= n_dev) return fail(nullptr, "invalid argument: device index out of range");
    auto *c = new b200tfhe_boolean_ctx();
    c->p = p; c->ks_first = keyswitch_first ? 1 : 0; c->device = device; c->log2N = (int)log2N;
    auto bail = [&](const std::string &m) { fail(nullptr, m); boolean_free(c); return 1; };
    if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice failed");
    if (cudaFuncSetAttribute((const void *)pbs_generic_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxOptinSmem) != cudaSuccess ||
        cudaFuncSetAttribute((const void *)bsk_to_fourier_generic_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxOptinSmem) != cudaSuccess)
        return bail("cudaFuncSetAttribute failed");
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate failed");
    std::vector<double2> roots, twist;
    make_generic_tables(p.polynomial_size, roots, twist);
    const size_t glwe = (size_t)(p.glwe_dimension + 1) * p.polynomial_size;
    std::vector<uint32_t> lut(glwe, 0);   // mask polynomials zero, body = 1/8 everywhere (bootstrapping.rs:55-62)
    for (size_t j = (size_t)p.glwe_dimension * p.polynomial_size; j < glwe; j++) lut[j] = 1u << 29;
    if (cudaMalloc(&c->d_ksk, c->ksk_len() * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&c->d_bsk, c->bsk_len() / 2 * sizeof(double2)) != cudaSuccess ||
        cudaMalloc(&c->d_lut, glwe * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&c->d_roots, roots.size() * sizeof(double2)) != cudaSuccess ||
        cudaMalloc(&c->d_twist, twist.size() * sizeof(double2)) != cudaSuccess)
        return bail("cudaMalloc(boolean keys) failed");
    if (cudaMemcpy(c->d_lut, lut.data(), glwe * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->d_roots, roots.data(), roots.size() * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(c->d_twist, twist.data(), twist.size() * sizeof(double2), cudaMemcpyHostToDevice) != cudaSuccess)
        return bail("cudaMemcpy(boolean tables) failed");
    *out = c;
    return 0;
}

int b200tfhe_boolean_ctx_destroy(b200tfhe_boolean_ctx *ctx) {
    if (ctx) boolean_free(ctx);
    return 0;
}

int b200tfhe_boolean_last_error(const b200tfhe_boolean_ctx *ctx, char *buf, size_t buf_len) {
    if (!ctx) return b200tfhe_last_global_error(buf, buf_len);
    if (buf && buf_len) std::snprintf(buf, buf_len, "%s", ctx->err.c_str());
    return 0;
}

int b200tfhe_boolean_load_ksk(b200tfhe_boolean_ctx *ctx, const uint32_t *ksk, size_t n_u32) {
    if (!ctx) return fail(nullptr, "null context");
    std::lock_guard<std::mutex> l(ctx->mu);
    if (!ksk || n_u32 != ctx->ksk_len()) return bfail(ctx, "invalid argument: ksk length does not match parameters (k*N*ks_level*(n+1))");
    BCU_TRY(ctx, cudaSetDevice(ctx->device));
    BCU_TRY(ctx, cudaMemcpy(ctx->d_ksk, ksk, n_u32 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    ctx->ksk_loaded = true;
    return 0;
}

int b200tfhe_boolean_load_bsk_standard(b200tfhe_boolean_ctx *ctx, const uint32_t *bsk, size_t n_u32) {
    if (!ctx) return fail(nullptr, "null context");
    std::lock_guard<std::mutex> l(ctx->mu);
    if (!bsk || n_u32 != ctx->bsk_len()) return bfail(ctx, "invalid argument: bsk length does not match parameters (n*pbs_level*(k+1)^2*N)");
    BCU_TRY(ctx, cudaSetDevice(ctx->device));
    uint32_t *tmp = nullptr;
    BCU_TRY(ctx, cudaMalloc(&tmp, n_u32 * sizeof(uint32_t)));
    cudaError_t e = cudaMemcpyAsync(tmp, bsk, n_u32 * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    const uint32_t N = ctx->p.polynomial_size;
    if (e == cudaSuccess) {
        bsk_to_fourier_generic_kernel<uint32_t><<<(unsigned)(n_u32 / N), 512, (size_t)N / 2 * sizeof(double2), ctx->stream>>>(
            tmp, ctx->d_bsk, ctx->d_roots, ctx->d_twist, nullptr, ctx->log2N, 1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    BCU_TRY(ctx, e);
    ctx->bsk_loaded = true;
    return 0;
}

/* gate: 0 AND, 1 NAND, 2 OR, 3 NOR, 4 XOR, 5 XNOR.  a, b, out: batch ciphertexts of ct_size words (host buffers; out may alias a or b). */
int b200tfhe_boolean_gate_batch(b200tfhe_boolean_ctx *ctx, int gate, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t batch) {
    if (!ctx) return fail(nullptr, "null context");
    std::lock_guard<std::mutex> l(ctx->mu);
    if (batch == 0) return 0;
    if (!a || !b || !out) return bfail(ctx, "invalid argument: null pointer");
    if (gate < 0 || gate > 5) return bfail(ctx, "invalid argument: gate must be in [0, 5]");
    if (!ctx->ksk_loaded || !ctx->bsk_loaded) return bfail(ctx, "server key not loaded");
    BCU_TRY(ctx, cudaSetDevice(ctx->device));
    if (int rc = boolean_workspace(ctx, batch)) return rc;
    const b200tfhe_params &p = ctx->p;
    const size_t ct = ctx->ct_size(), bytes = batch * ct * sizeof(uint32_t);
    static const uint32_t kCoeff[6] = {1u, 0xFFFFFFFFu, 1u, 0xFFFFFFFFu, 2u, 0xFFFFFFFEu};
    static const uint32_t kCst[6] = {7u << 29, 1u << 29, 1u << 29, 7u << 29, 2u << 29, 6u << 29};   // -1/8, 1/8, 1/8, -1/8, 2/8, -2/8
    BCU_TRY(ctx, cudaMemcpyAsync(ctx->d_a, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    BCU_TRY(ctx, cudaMemcpyAsync(ctx->d_b, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
    boolean_prelin_kernel<<<(unsigned)std::min<size_t>(148 * 8, (batch * ct + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->d_a, ctx->d_b, ctx->d_mid, kCoeff[gate], kCst[gate], (int)ct, batch * ct);
    GenPbsArgs g{};
    g.lut_idx = nullptr; g.luts = ctx->d_lut; g.bsk = ctx->d_bsk; g.roots = ctx->d_roots; g.twist = ctx->d_twist;
    g.acc_ws = ctx->d_acc; g.fourier_ws = ctx->d_fourier; g.batch = (int)batch; g.n = (int)p.lwe_dimension; g.k = (int)p.glwe_dimension;
    g.log2N = ctx->log2N; g.base_log = (int)p.pbs_base_log; g.level = (int)p.pbs_level; g.fft_in_smem = 1; g.n_luts = 1; g.err_flag = nullptr;
    const GenLaunch gl = gen_launch_shape(p.polynomial_size, p.glwe_dimension, p.pbs_level, true, kMaxOptinSmem);
    g.polys_in_smem = gl.polys_in_smem;
    const size_t smem = gl.smem;   // FFT buffers + roots of unity
    const int n_in = (int)(p.glwe_dimension * p.polynomial_size);
    const unsigned threads = gl.threads;
    if (ctx->ks_first) {
        ks_generic_kernel<uint32_t><<<(unsigned)batch, 256, 0, ctx->stream>>>(ctx->d_mid, ctx->d_ksk, ctx->d_mid2, (int)batch, n_in, (int)ctx->small(),
                                                                               (int)p.ks_base_log, (int)p.ks_level);
        g.lwe_small = ctx->d_mid2; g.out = ctx->d_out;
        pbs_generic_kernel<uint32_t><<<(unsigned)batch, threads, smem, ctx->stream>>>(g);
    } else {
        g.lwe_small = ctx->d_mid; g.out = ctx->d_mid2;
        pbs_generic_kernel<uint32_t><<<(unsigned)batch, threads, smem, ctx->stream>>>(g);
        ks_generic_kernel<uint32_t><<<(unsigned)batch, 256, 0, ctx->stream>>>(ctx->d_mid2, ctx->d_ksk, ctx->d_out, (int)batch, n_in, (int)ctx->small(),
                                                                               (int)p.ks_base_log, (int)p.ks_level);
    }
    BCU_TRY(ctx, cudaGetLastError());
    BCU_TRY(ctx, cudaMemcpyAsync(out, ctx->d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    BCU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // extern "C"
