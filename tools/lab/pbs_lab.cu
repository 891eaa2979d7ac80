// pbs_lab.cu -- development harness (not product): times the PBS kernels on synthetic keys without the
// library plumbing, and (built with -DB200TFHE_TIMELINE) dumps per-warp phase timestamps of CTA 0.
//   pbs_lab <kernel: 3|4> <cts> <batch> [reps] [timeline_file]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <random>
#include "../../tfhe_rs_string_b200/csrc/pbs_kernel3.cuh"
#ifdef LAB_HAVE_K4
#include "../../tfhe_rs_string_b200/csrc/pbs_kernel4.cuh"
#endif
using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static void make_twiddles(std::vector<double2> &t) {
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

template <int CTS, int PH = 0>
static void launch3(const PbsArgs &a, cudaStream_t s) {
    constexpr size_t smem = pbs3_smem_bytes<CTS>();
    CK(cudaFuncSetAttribute(pbs_kernel3<CTS, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pbs_kernel3<CTS, PH><<<(a.batch + CTS - 1) / CTS, CTS * 64, smem, s>>>(a);
}

int main(int argc, char **argv) {
    const int kernel = argc > 1 ? atoi(argv[1]) : 3;
    const int cts = argc > 2 ? atoi(argv[2]) : 4;
    const int batch = argc > 3 ? atoi(argv[3]) : 4096;
    const int reps = argc > 4 ? atoi(argv[4]) : 3;
    const char *tl = argc > 5 ? argv[5] : nullptr;
    const int n = 742;
    std::mt19937_64 rng(1);
    std::vector<double2> tw; make_twiddles(tw);
    std::vector<double2> bsk((size_t)n * 4 * kHalf);
    std::normal_distribution<double> nd(0.0, 9.0 / 1024.0);
    for (auto &v : bsk) v = make_double2(nd(rng), nd(rng));
    std::vector<uint64_t> lwe((size_t)batch * (n + 1)), lut(2 * kN, 0);
    for (auto &v : lwe) v = rng();
    for (int i = 0; i < kN; i++) lut[kN + i] = (uint64_t)(i / 128) << 59;
    double2 *d_tw, *d_bsk; uint64_t *d_lwe, *d_lut, *d_out, *d_out2;
    CK(cudaMalloc(&d_tw, tw.size() * sizeof(double2))); CK(cudaMemcpy(d_tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_bsk, bsk.size() * sizeof(double2))); CK(cudaMemcpy(d_bsk, bsk.data(), bsk.size() * sizeof(double2), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_lwe, lwe.size() * 8)); CK(cudaMemcpy(d_lwe, lwe.data(), lwe.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_lut, lut.size() * 8)); CK(cudaMemcpy(d_lut, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, (size_t)batch * (kN + 1) * 8)); CK(cudaMalloc(&d_out2, (size_t)batch * (kN + 1) * 8));
    PbsArgs a{};
    a.lwe_small = d_lwe; a.lut_idx = nullptr; a.luts = d_lut; a.bsk = d_bsk; a.twid = d_tw; a.out = d_out; a.batch = batch; a.n = n; a.n_luts = 1; a.err_flag = nullptr; a.dbg = nullptr;
    const size_t ndbg = 8 * 8 * 16;
    if (tl) { CK(cudaMalloc(&a.dbg, ndbg * sizeof(long long))); CK(cudaMemset(a.dbg, 0, ndbg * sizeof(long long))); }
    auto run = [&](const PbsArgs &x) {
        if (kernel == 31) launch3<4, 1>(x, 0);
        else if (kernel == 3) {
            switch (cts) { case 1: launch3<1>(x, 0); break; case 2: launch3<2>(x, 0); break; case 3: launch3<3>(x, 0); break; default: launch3<4>(x, 0); }
        }
#ifdef LAB_HAVE_K4
        else launch4(x, cts, 0);
#endif
    };
    run(a); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); run(a); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // reference output from pbs_kernel3<4> for a bit-compare of the first 8 ciphertexts' bodies (same arithmetic => same bits)
    PbsArgs b = a; b.out = d_out2; b.dbg = nullptr;
    launch3<4>(b, 0); CK(cudaDeviceSynchronize());
    std::vector<uint64_t> o1((size_t)batch * (kN + 1)), o2(o1.size());
    CK(cudaMemcpy(o1.data(), d_out, o1.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(o2.data(), d_out2, o2.size() * 8, cudaMemcpyDeviceToHost));
    size_t ndiff = 0; long long maxd = 0;
    for (size_t i = 0; i < o1.size(); i++) if (o1[i] != o2[i]) { ndiff++; long long d = (long long)(o1[i] - o2[i]); if (d < 0) d = -d; if (d > maxd) maxd = d; }
    printf("{\"kernel\": %d, \"cts\": %d, \"batch\": %d, \"ms\": %.4f, \"pbs_per_s\": %.1f, \"tflops\": %.3f, \"words_differing_from_k3\": %zu, \"max_abs_diff_log2\": %.1f}\n",
           kernel, cts, batch, best, batch / (best * 1e-3), batch * 1.94510848e8 / (best * 1e-3) / 1e12, ndiff, maxd ? log2((double)maxd) : 0.0);
    if (tl) {
        std::vector<long long> h(ndbg);
        CK(cudaMemcpy(h.data(), a.dbg, ndbg * sizeof(long long), cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(tl, "w")) {
            for (size_t r = 0; r < 64; r++) { fprintf(f, "%zu %zu", r / 8, r % 8); for (int k = 0; k < 11; k++) fprintf(f, " %lld", h[r * 16 + k]); fprintf(f, "\n"); }
            fclose(f);
        }
    }
    return 0;
}
