// context.hpp -- device memory manager and scheduler state behind the C ABI (include/b200tfhe.h).
//
// One b200tfhe_ctx owns 1..N GPUs of one box.  Per GPU (DevCtx): three streams (compute, H2D, D2H), the key arena
// (Fourier BSK | KSK | KSK byte limbs), the device LUT store, batch workspaces, a pinned staging ring for pageable
// host buffers, an event pool, and (N > 1) one persistent host worker thread.  The context replaces what the
// reference keeps in ShortintEngine's thread-local scratch (shortint/engine/mod.rs:23-25,40-69) and what its
// benches get from rayon (benches/core_crypto/pbs_bench.rs:517-531: one ciphertext per worker): here a batch is cut
// into contiguous shards, one per GPU, and every shard is pipelined wave by wave on its GPU.
#pragma once
#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200tfhe.h"
#include "ks_mma.cuh"

namespace b200 {

struct EventPair {
    cudaEvent_t a, b;
};

// contiguous shard [begin, end) of `total` items for worker `rank` of `world` (same rule as
// tfhe_rs_string_b200/multigpu.py: the first total % world shards get one extra item)
inline void shard_bounds(size_t total, int world, int rank, size_t *begin, size_t *end) {
    const size_t base = total / (size_t)world, extra = total % (size_t)world;
    *begin = (size_t)rank * base + std::min<size_t>((size_t)rank, extra);
    *end = *begin + base + ((size_t)rank < extra ? 1 : 0);
}

// one persistent host thread per GPU: jobs are posted by the API thread and awaited before the call returns
class Worker {
  public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        {
            std::lock_guard<std::mutex> l(m_);
            quit_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<int()> job) {
        {
            std::lock_guard<std::mutex> l(m_);
            job_ = std::move(job);
            has_job_ = true;
            done_ = false;
        }
        cv_.notify_all();
    }
    int wait() {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [this] { return done_; });
        return rc_;
    }

  private:
    void loop() {
        for (;;) {
            std::function<int()> job;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return has_job_ || quit_; });
                if (quit_) return;
                job = std::move(job_);
                has_job_ = false;
            }
            const int rc = job();
            {
                std::lock_guard<std::mutex> l(m_);
                rc_ = rc;
                done_ = true;
            }
            cv_.notify_all();
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::function<int()> job_;
    bool has_job_ = false, done_ = true, quit_ = false;
    int rc_ = 0;
    std::thread th_;
};

constexpr int kStageSlabs = 3;   // pinned slabs per direction (upload runs two waves ahead of the kernels)

struct DevCtx {
    int device = 0, sm_count = 148;
    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;

    // key arena: [Fourier BSK][KSK][KSK byte limbs], contiguous so it can be copied GPU to GPU in one transfer
    unsigned char *arena = nullptr;
    double2 *d_twid = nullptr;                      // fast path: inter-pass twiddles T'[k1][lane]
    double2 *d_roots = nullptr, *d_twist = nullptr; // generic path: roots of unity / negacyclic twist
    uint32_t *d_err_flag = nullptr;                 // or-ed by the kernels when a device-side lut id is out of range

    uint64_t *d_luts = nullptr;                     // [lut_cap][(k+1) N]
    size_t lut_cap = 0, lut_count = 0;

    // batch workspaces (grown on demand, never shrunk)
    uint64_t *d_in = nullptr, *d_small = nullptr, *d_out = nullptr;
    uint32_t *d_lut_idx = nullptr;
    uint8_t *d_digits = nullptr;
    uint64_t *d_acc_ws = nullptr;                   // generic PBS: accumulators
    double2 *d_fourier_ws = nullptr;                // generic PBS: Fourier accumulators (+ FFT buffer when it does not fit smem)
    size_t ws_cap = 0;

    // pinned staging ring for pageable host buffers
    unsigned char *h_in[kStageSlabs] = {}, *h_out[kStageSlabs] = {};
    size_t slab_bytes = 0;

    std::vector<cudaEvent_t> ev_pool;               // timing-disabled events, reused across calls
    size_t ev_next = 0;

    bool profiling = false;
    std::vector<EventPair> ev_ks, ev_pbs;
    double ks_ms = 0, pbs_ms = 0;
    uint64_t ks_launches = 0, pbs_launches = 0, kernel_launches = 0;

    std::unique_ptr<Worker> worker;                 // only for devices 1..N-1 of a multi-GPU context
    std::string err;                                // last error raised on this device's worker thread
};

}  // namespace b200

struct b200tfhe_ctx {
    b200tfhe_params p{};
    bool fast_path = false;                         // k = 1, N = 2048, one level of base 2^23: pbs_kernel5 / pbs_lat_kernel
    bool ks_tensor = true;                          // keyswitch on tcgen05 (else ks_generic_kernel)
    int log2N = 11;
    bool fft_in_smem = true;
    std::vector<std::unique_ptr<b200::DevCtx>> devs;
    std::mutex mu;                                  // serialises API calls on the context
    mutable std::mutex err_mu;
    mutable std::string err;

    size_t arena_bytes = 0, off_bsk = 0, off_ksk = 0, off_ksk_limbs = 0;
    b200::KsMmaGeom ks_geom{};
    bool ksk_loaded = false, bsk_loaded = false;

    std::vector<std::vector<uint64_t>> h_luts;      // host copies (content addressing, re-upload on growth)
    std::unordered_multimap<uint64_t, uint32_t> lut_hash;

    size_t big_size() const { return (size_t)p.glwe_dimension * p.polynomial_size + 1; }
    size_t small_size() const { return (size_t)p.lwe_dimension + 1; }
    size_t glwe_len() const { return (size_t)(p.glwe_dimension + 1) * p.polynomial_size; }
    size_t ksk_len() const { return (size_t)p.glwe_dimension * p.polynomial_size * p.ks_level * small_size(); }
    size_t bsk_len() const {
        return (size_t)p.lwe_dimension * p.pbs_level * (p.glwe_dimension + 1) * (p.glwe_dimension + 1) * p.polynomial_size;
    }
    size_t fourier_per_ct() const {                 // double2 per ciphertext in d_fourier_ws
        return (size_t)(p.glwe_dimension + 1 + (fft_in_smem ? 0 : 1)) * (p.polynomial_size / 2);
    }
    double2 *d_bsk(const b200::DevCtx &d) const { return reinterpret_cast<double2 *>(d.arena + off_bsk); }
    uint64_t *d_ksk(const b200::DevCtx &d) const { return reinterpret_cast<uint64_t *>(d.arena + off_ksk); }
    uint8_t *d_ksk_limbs(const b200::DevCtx &d) const { return d.arena + off_ksk_limbs; }
};
