// pbs_lab.cu -- development harness (not product): times the PBS kernels on synthetic keys without the
// library plumbing, and (built with -DB200TFHE_TIMELINE) dumps per-warp phase timestamps of CTA 0.
//   pbs_lab <kernel: 3|4> <cts> <batch> [reps] [timeline_file]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstring>
#include <random>
#include "pbs_kernel3.cuh"
#include "../../tfhe_rs_string_b200/csrc/pbs_kernel5.cuh"
#include "../../tfhe_rs_string_b200/csrc/pbs_kernel_lat.cuh"
#include "../../tfhe_rs_string_b200/csrc/pbs_kernel_lat4.cuh"
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

template <int CTS, int PH = 0>
static void launch5(const PbsArgs &a, cudaStream_t s) {
    constexpr size_t smem = pbs5_smem_bytes<CTS>();
    CK(cudaFuncSetAttribute(pbs_kernel5<CTS, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pbs_kernel5<CTS, PH><<<(a.batch + CTS - 1) / CTS, CTS * 64, smem, s>>>(a);
}

template <int CTS, bool SPREAD = false>
static void launch_lat(const PbsArgs &a, cudaStream_t s) {
    constexpr size_t smem = pbs_lat_smem_bytes<CTS, SPREAD>();
    CK(cudaFuncSetAttribute(pbs_lat_kernel<CTS, SPREAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pbs_lat_kernel<CTS, SPREAD><<<(a.batch + CTS - 1) / CTS, SPREAD ? 128 : 256, smem, s>>>(a);
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
#ifdef B200TFHE_LAB_DELAY
    { const int dl = getenv("LAB_DELAY") ? atoi(getenv("LAB_DELAY")) : 0; CK(cudaMemcpyToSymbol(g_lab_delay, &dl, sizeof(int))); }
#endif
    PbsArgs a{};
    a.lwe_small = d_lwe; a.lut_idx = nullptr; a.luts = d_lut; a.bsk = d_bsk; a.twid = d_tw; a.out = d_out; a.batch = batch; a.n = n; a.n_luts = 1; a.err_flag = nullptr; a.dbg = nullptr;
    const size_t ndbg = 8 * 8 * 16;
    if (tl) { CK(cudaMalloc(&a.dbg, ndbg * sizeof(long long))); CK(cudaMemset(a.dbg, 0, ndbg * sizeof(long long))); }
    auto run = [&](const PbsArgs &x) {
        if (kernel == 31) launch3<4, 1>(x, 0);
        else if (kernel == 51) launch5<4, 1>(x, 0);
        else if (kernel == 7) { if (cts == 1) launch_lat<1>(x, 0); else launch_lat<2>(x, 0); }
        else if (kernel == 74) {   // one ciphertext per SM, four warps per polynomial
            CK(cudaFuncSetAttribute(pbs_lat4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pbs_lat4_smem_bytes()));
            pbs_lat4_kernel<<<x.batch, 256, pbs_lat4_smem_bytes(), 0>>>(x);
        }
        else if (kernel == 72) launch_lat<1, true>(x, 0);   // one ciphertext per SM, the halves of a polynomial on different sub-partitions
        else if (kernel == 5) {
            switch (cts) { case 1: launch5<1>(x, 0); break; case 2: launch5<2>(x, 0); break; case 3: launch5<3>(x, 0); break; default: launch5<4>(x, 0); }
        }
        else if (kernel == 3) {
            switch (cts) { case 1: launch3<1>(x, 0); break; case 2: launch3<2>(x, 0); break; case 3: launch3<3>(x, 0); break; default: launch3<4>(x, 0); }
        }
#ifdef LAB_HAVE_K4
        else launch4(x, cts, 0);
#endif
    };
    if (getenv("LAB_BISECT_CT")) {   // development: first CMUX step at which kernel `kernel` and pbs_kernel3<4> disagree on one ciphertext
        const int c0 = atoi(getenv("LAB_BISECT_CT")), g0 = c0 / 4 * 4;
        PbsArgs s5 = a, s3 = a;
        uint64_t *d_small; CK(cudaMalloc(&d_small, 4 * (n + 1) * 8));
        const size_t ndump = 8 * 32 * 32 * 6;
        long long *d_dump5, *d_dump3; CK(cudaMalloc(&d_dump5, ndump * 8)); CK(cudaMalloc(&d_dump3, ndump * 8));
        s5.lwe_small = d_small; s5.batch = 4; s5.dbg = d_dump5; s3 = s5; s3.out = d_out2; s3.dbg = d_dump3;
        std::vector<uint64_t> p1(4 * (kN + 1)), p2(p1.size()), sm(4 * (n + 1));
        for (int nn = 1; nn <= n; nn++) {
            for (int k = 0; k < 4; k++) {
                for (int i2 = 0; i2 < nn; i2++) sm[(size_t)k * (nn + 1) + i2] = lwe[(size_t)(g0 + k) * (n + 1) + i2];
                sm[(size_t)k * (nn + 1) + nn] = lwe[(size_t)(g0 + k) * (n + 1) + n];
            }
            CK(cudaMemcpy(d_small, sm.data(), 4 * (nn + 1) * 8, cudaMemcpyHostToDevice));
            s5.n = s3.n = nn;
            run(s5); launch3<4>(s3, 0); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(p1.data(), d_out, p1.size() * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(p2.data(), d_out2, p2.size() * 8, cudaMemcpyDeviceToHost));
            long long m = 0; int nd = 0, where = -1;
            for (size_t k = 0; k < p1.size(); k++) { long long d = (long long)(p1[k] - p2[k]); if (d < 0) d = -d; if (d) { nd++; if (d > m) { m = d; where = (int)k; } } }
            if (m > 1) {
                const int cc = where / (kN + 1);
                printf("first divergence after %d steps: %d words differ, max |diff| 2^%.1f at word %d (ct %d, coef %d); a~ of last step = %u\n", nn, nd, log2((double)m), where, g0 + cc, where % (kN + 1),
                       (unsigned)((((lwe[(size_t)(g0 + cc) * (n + 1) + nn - 1]) >> 51) + 1) >> 1));
                std::vector<long long> h5(ndump), h3(ndump);
                CK(cudaMemcpy(h5.data(), d_dump5, ndump * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h3.data(), d_dump3, ndump * 8, cudaMemcpyDeviceToHost));
                int shown = 0;
                for (int w = 0; w < 8; w++) for (int mm = 0; mm < 32; mm++) for (int l = 0; l < 32; l++) {
                    const long long *q5 = &h5[(((size_t)w * 32 + mm) * 32 + l) * 6], *q3 = &h3[(((size_t)w * 32 + mm) * 32 + l) * 6];
                    if (q5[0] != q3[0] || q5[1] != q3[1]) printf("  DIGIT mismatch warp %d m %d lane %d: k5 %lld %lld k3 %lld %lld\n", w, mm, l, q5[0], q5[1], q3[0], q3[1]);
                }
                for (int w = 0; w < 8; w++) for (int mm = 0; mm < 32; mm++) for (int l = 0; l < 32; l++) {
                    const long long *q5 = &h5[(((size_t)w * 32 + mm) * 32 + l) * 6], *q3 = &h3[(((size_t)w * 32 + mm) * 32 + l) * 6];
                    bool dif = false; for (int k = 0; k < 6; k++) if (q5[k] != q3[k] && !(k >= 4 && llabs(q5[k] - q3[k]) <= 1)) dif = true;
                    if (dif && shown < 4) { shown++; double y5r, y3r, y5i, y3i; memcpy(&y5r, &q5[2], 8); memcpy(&y3r, &q3[2], 8); memcpy(&y5i, &q5[3], 8); memcpy(&y3i, &q3[3], 8);
                        printf("  warp %d m %d lane %d: digits k5 %lld %lld k3 %lld %lld | y k5 %.17g %.17g k3 %.17g %.17g | delta k5 %016llx %016llx k3 %016llx %016llx\n", w, mm, l, q5[0], q5[1], q3[0], q3[1], y5r, y5i, y3r, y3i,
                               (unsigned long long)q5[4], (unsigned long long)q5[5], (unsigned long long)q3[4], (unsigned long long)q3[5]); }
                }
                break;
            }
        }
        return 0;
    }
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
    {   // diagnostics: per-ciphertext difference classes, and run-to-run determinism of the kernel under test
        int n_big = 0, n_small = 0, first_big = -1; size_t small_words = 0;
        for (int c = 0; c < batch; c++) {
            long long m = 0; size_t w = 0;
            for (int k = 0; k <= kN; k++) { long long d = (long long)(o1[(size_t)c * (kN + 1) + k] - o2[(size_t)c * (kN + 1) + k]); if (d < 0) d = -d; if (d) w++; if (d > m || d < 0) m = d < 0 ? (1ll << 62) : d; }
            if (m > (1ll << 55)) { n_big++; if (first_big < 0) first_big = c; } else if (m) { n_small++; small_words += w; }
        }
        CK(cudaMemset(d_out2, 0, o2.size() * 8));
        PbsArgs c2 = a; c2.out = d_out2; c2.dbg = nullptr; run(c2); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(o2.data(), d_out2, o2.size() * 8, cudaMemcpyDeviceToHost));
        size_t nrep = 0; for (size_t i = 0; i < o1.size(); i++) if (o1[i] != o2[i]) nrep++;
        printf("{\"cts_big_diff\": %d, \"first_big\": %d, \"cts_small_diff\": %d, \"small_words\": %zu, \"words_differing_between_two_runs\": %zu}\n", n_big, first_big, n_small, small_words, nrep);
    }
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
