/*
 * b200tfhe.h -- C ABI of libb200tfhe.so, the B200-native (sm_100a) drop-in for the reference's
 * shortint keyswitch + programmable-bootstrap path.
 *
 * Reference: M-Bln/tfhe-rs-string (fork of tfhe-rs 0.5.0), paths relative to tfhe/src/.
 * The reference has no plugin interface for this path: callers use inherent methods of
 * shortint::ServerKey.  Each entry point below names the reference function(s) whose work it
 * replaces; INTEGRATION.md shows the Rust `extern "C"` shim a maintainer adds so that
 * ServerKey::apply_lookup_table[_assign] dispatches here.
 *
 * Conventions (mirroring the reference's own C API, c_api/utils.rs:3-27):
 *   - every function returns int: 0 = success, non-zero = error (never unwinds / aborts);
 *     b200tfhe_last_error() returns the message of the last failing call on that context;
 *   - only plain pointers and sizes cross the ABI; ciphertext metadata (degree, noise level)
 *     stays with the caller exactly as at shortint/server_key/mod.rs:855-856;
 *   - all ciphertext/key layouts are the reference's flat u64 layouts, verbatim:
 *       LWE ciphertext   [a_0 .. a_{n-1}, b]                 (entities/lwe_ciphertext.rs:598-609)
 *       GLWE accumulator [mask polys .., body poly]          (entities/glwe_ciphertext.rs:423-433)
 *       KSK              [in_dim][level (l first)][out_dim+1]  (entities/lwe_keyswitch_key.rs:77-108,
 *                                                             algorithms/lwe_keyswitch_key_generation.rs:109-111)
 *       standard BSK     [n][level 1..l][k+1 rows][k+1 polys][N] (entities/ggsw_ciphertext.rs:185-197)
 *   - there is NO CPU fallback: without a CUDA device every call fails with an error.
 *
 * "host" entry points borrow host buffers for the duration of the call (H2D, compute, D2H,
 * synchronous).  "_device" entry points take device pointers, enqueue on the context's stream
 * and return immediately; use b200tfhe_sync() or your own event on b200tfhe_stream().
 */
#ifndef B200TFHE_H
#define B200TFHE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200tfhe_ctx b200tfhe_ctx;

/* shortint/parameters/mod.rs:62-76 (ClassicPBSParameters), native modulus 2^64; either PBS order.
 * Every classic parameter set of the reference is accepted: glwe_dimension 1..8, polynomial_size 256..32768,
 * any PBS / KS decomposition with base_log*level < 64, lwe_dimension <= 4096.  Sets with glwe_dimension = 1,
 * polynomial_size = 2048, pbs_level = 1, pbs_base_log = 23 (PARAM_MESSAGE_1_CARRY_3, 2_2, 3_1, 4_0 _KS_PBS and
 * 2_2 _PBS_KS, shortint/parameters/mod.rs:688-747,1155-1169) run on the specialised kernels (pbs_kernel5 / pbs_lat4_kernel /
 * pbs_lat_kernel); all others (1_1 k=3 N=512 :613-627, 3_3 N=8192 l=2 :853-867, 4_4 N=32768 :1063-1077, ...) on the
 * generic one-CTA-per-ciphertext kernel (csrc/pbs_generic.cuh). */
typedef struct {
    uint32_t lwe_dimension;   /* n  */
    uint32_t glwe_dimension;  /* k  */
    uint32_t polynomial_size; /* N  */
    uint32_t pbs_base_log, pbs_level;
    uint32_t ks_base_log, ks_level;
    uint32_t message_modulus, carry_modulus;
} b200tfhe_params;

/* ---- lifetime ------------------------------------------------------------------------- */
/* Creates a context bound to CUDA device `device` (one context per GPU; one process per GPU
 * in multi-GPU runs).  Replaces ShortintEngine's thread-local scratch (shortint/engine/mod.rs:
 * 23-25,40-69): all scratch lives on the device and is owned by the context. */
int b200tfhe_ctx_create(const b200tfhe_params *params, int device, b200tfhe_ctx **out);
/* One context over several GPUs of one box (the reference's analogue is the rayon fan-out of
 * benches/core_crypto/pbs_bench.rs:517-531).  Keys are uploaded and converted once on devices[0] and copied GPU to
 * GPU (cudaMemcpyPeerAsync, NVLink); LUTs are registered on every GPU; every host-buffer batch call is cut into
 * contiguous shards, one per GPU, each pipelined by its own host thread and streams; named programs are split over
 * their independent units (integers, strings).  There is no per-PBS collective.  "_device" entry points act on
 * devices[0]; b200tfhe_ks_pbs_batch_device_multi takes one device-resident shard per GPU. */
int b200tfhe_ctx_create_multi(const b200tfhe_params *params, const int *devices, int n_devices, b200tfhe_ctx **out);
int b200tfhe_ctx_device_count(const b200tfhe_ctx *ctx, int *n_devices);
int b200tfhe_ctx_destroy(b200tfhe_ctx *ctx);
int b200tfhe_last_error(const b200tfhe_ctx *ctx, char *buf, size_t buf_len);
/* Last error of a failed b200tfhe_ctx_create (no context exists yet). */
int b200tfhe_last_global_error(char *buf, size_t buf_len);

/* ---- server key upload ---------------------------------------------------------------- */
/* key_switching_key.as_ref() of shortint::ServerKey (shortint/server_key/mod.rs:284-297). */
int b200tfhe_load_ksk(b200tfhe_ctx *ctx, const uint64_t *ksk, size_t n_u64);
/* Standard-domain bootstrap key, taken where the reference still has it
 * (shortint/engine/server_side.rs:63-86, shortint/server_key/mod.rs:951-966); converted on the
 * GPU to this library's Fourier layout (replaces par_convert_standard_lwe_bootstrap_key_to_fourier,
 * algorithms/lwe_bootstrap_key_conversion.rs:99+). */
int b200tfhe_load_bsk_standard(b200tfhe_ctx *ctx, const uint64_t *bsk, size_t n_u64);
/* Device-resident key arena of devices[0] (Fourier BSK || KSK || KSK byte limbs), contiguous, so a one-process-per-GPU
 * launcher can broadcast it once (NCCL) instead of re-uploading: rank 0 loads keys, every rank passes its arena
 * pointer to the collective, then non-root ranks call b200tfhe_keys_adopt() (which also replicates the arena to the
 * other GPUs of a multi-GPU context). */
int b200tfhe_key_arena(b200tfhe_ctx *ctx, void **device_ptr, size_t *bytes);
int b200tfhe_keys_adopt(b200tfhe_ctx *ctx);

/* ---- server key wire format (csrc/key_import.hpp; PARITY UNPINNED: no Rust toolchain in the build image) ------ */
/* Where the key material sits inside a serialised server key.  Offsets are in bytes from the start of the buffer;
 * serde does not align, so copy before reinterpreting. */
typedef struct {
    uint64_t ksk_offset, ksk_len;          /* ksk_len u64 words, reference KSK layout                            */
    uint64_t bsk_offset, bsk_len;          /* standard key: bsk_len u64 words; Fourier key: bsk_len complex f64  */
    uint64_t bsk_poly_stride_bytes;        /* distance between consecutive polynomials (Fourier: 8-byte length prefix each) */
    uint32_t bsk_is_fourier;               /* 1: bincode(shortint::ServerKey), Fourier key in concrete-fft's serialisation order */
    uint32_t pbs_order;                    /* 0 = KeyswitchBootstrap, 1 = BootstrapKeyswitch (commons/parameters.rs:234-245) */
    uint64_t max_degree, max_noise_level;  /* ServerKey metadata (0 for the standard-domain bundle)               */
} b200tfhe_key_view;
/* Parses bincode(shortint::ServerKey) (shortint/server_key/mod.rs:283-297) or the standard-domain bundle
 * (LweKeyswitchKey<Vec<u64>>, LweBootstrapKey<Vec<u64>>, MessageModulus, CarryModulus, PBSOrder) written by the Rust
 * shim; fills the parameter set and the view.  Needs no GPU. */
int b200tfhe_parse_server_key(const uint8_t *bytes, size_t n_bytes, b200tfhe_params *params, b200tfhe_key_view *view);
/* Parses and uploads both keys.  The context must have been created with the key's dimensions and decompositions. */
int b200tfhe_load_server_key_bytes(b200tfhe_ctx *ctx, const uint8_t *bytes, size_t n_bytes);

/* ---- lookup tables -------------------------------------------------------------------- */
/* Registers a GLWE accumulator ((k+1)*N u64, as produced by generate_lookup_table /
 * fill_accumulator, shortint/server_key/mod.rs:383-399, shortint/engine/mod.rs:72-128) and
 * returns its id.  Content-addressed: registering the same table twice returns the same id. */
int b200tfhe_register_lut(b200tfhe_ctx *ctx, const uint64_t *glwe_acc, uint32_t *id);
/* fill_accumulator (shortint/engine/mod.rs:72-128) for a function table f(0..message_modulus*carry_modulus-1),
 * then register_lut: what ServerKey::generate_lookup_table does, with the table kept on the device. */
int b200tfhe_register_lut_from_table(b200tfhe_ctx *ctx, const uint64_t *table, size_t table_len, uint32_t *id);

/* ---- hot path, host buffers ----------------------------------------------------------- */
/* keyswitch_lwe_ciphertext (core_crypto/algorithms/lwe_keyswitch.rs:96-170), batched.
 * in: batch x (k*N+1), out: batch x (n+1). */
int b200tfhe_keyswitch_batch(b200tfhe_ctx *ctx, const uint64_t *in, uint64_t *out, size_t batch);
/* programmable_bootstrap_lwe_ciphertext_mem_optimized (core_crypto/algorithms/
 * lwe_programmable_bootstrapping.rs:1067-1110 -> fft64/crypto/bootstrap.rs:333-364), batched.
 * in: batch x (n+1), lut_id: batch ids (NULL = id 0 for all), out: batch x (k*N+1). */
int b200tfhe_pbs_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch);
/* ServerKey::keyswitch_programmable_bootstrap_assign (shortint/server_key/mod.rs:783-857,
 * classic branch, non-trivial ciphertexts; the trivial shortcut :788-791 stays on the caller's
 * side) == apply_lookup_table_assign for PBSOrder::KeyswitchBootstrap (:465-476), batched.
 * in/out: batch x (k*N+1); in == out allowed. */
int b200tfhe_ks_pbs_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch);

/* ServerKey::programmable_bootstrap_keyswitch_assign (shortint/server_key/mod.rs:859-933) ==
 * apply_lookup_table_assign for PBSOrder::BootstrapKeyswitch (:471-474, e.g.
 * PARAM_MESSAGE_2_CARRY_2_PBS_KS, shortint/parameters/mod.rs:1155-1169), batched: bootstrap the small
 * ciphertext, then keyswitch the result back.  in/out: batch x (n+1); in == out allowed. */
int b200tfhe_pbs_ks_batch(b200tfhe_ctx *ctx, const uint64_t *in, const uint32_t *lut_id, uint64_t *out, size_t batch);

/* ---- hot path, device buffers (asynchronous on the context stream) -------------------- */
int b200tfhe_keyswitch_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, uint64_t *d_out, size_t batch);
int b200tfhe_pbs_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch);
int b200tfhe_ks_pbs_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch);
int b200tfhe_pbs_ks_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_in, const uint32_t *d_lut_id, uint64_t *d_out, size_t batch);
/* Multi-GPU contexts: element i of every array belongs to GPU i of the context (pointers resident on that GPU,
 * batch[i] ciphertexts; d_lut_id or d_lut_id[i] may be NULL).  Asynchronous on every GPU; b200tfhe_sync waits for all. */
int b200tfhe_ks_pbs_batch_device_multi(b200tfhe_ctx *ctx, const uint64_t *const *d_in, const uint32_t *const *d_lut_id,
                                       uint64_t *const *d_out, const size_t *batch);
/* Lookup-table ids of device-buffer calls cannot be checked on the host: the kernels check them, use table 0 for an
 * out-of-range id and raise a flag that the next b200tfhe_sync reports as an error. */
/* lwe_linear_algebra.rs (:68 add, :276 plaintext add, :556 cleartext mul, :703 sub) and the
 * bivariate pack (shortint/server_key/bivariate_pbs.rs:173-181), one launch:
 *   out[b] = ca[b] * x[ia[b]] + cb[b] * y[ib[b]];  out[b].body += pt[b]
 * All pointers are device pointers; ia/ib/cb/pt/y may be NULL (identity / none). lwe_size is
 * the number of u64 per ciphertext. */
int b200tfhe_lwe_linear_batch_device(b200tfhe_ctx *ctx, const uint64_t *d_x, const uint64_t *d_y,
                                     const int32_t *d_ia, const int32_t *d_ib, const int64_t *d_ca,
                                     const int64_t *d_cb, const uint64_t *d_pt, uint64_t *d_out,
                                     size_t batch, size_t lwe_size);
int b200tfhe_sync(b200tfhe_ctx *ctx);
/* cudaStream_t of the context, as void*, so callers can record their own events on it. */
int b200tfhe_stream(b200tfhe_ctx *ctx, void **stream);

/* ---- measurement ---------------------------------------------------------------------- */
/* When enabled, every ks / pbs kernel launch is bracketed by CUDA events on the context stream. */
int b200tfhe_set_profiling(b200tfhe_ctx *ctx, int enabled);
/* Accumulated device time (ms) and launch counts since the last reset; synchronises. */
int b200tfhe_get_kernel_times(b200tfhe_ctx *ctx, double *ks_ms, uint64_t *ks_launches, double *pbs_ms,
                              uint64_t *pbs_launches, int reset);
/* Kernels launched by this library on all GPUs of the context since it was created. */
int b200tfhe_kernel_launch_count(b200tfhe_ctx *ctx, uint64_t *count);

/* ---- batched call sites: level-synchronous programs ---------------------------------- */
/* The reference's integer / FheString / Trivium layers issue vectors of independent (ciphertext,
 * lookup table) pairs per dependency level (integer/server_key/radix_parallel/comparison.rs:22-28,
 * add.rs:529-535, comparator.rs:257-288; examples/fhe_strings/server_key/{comparisons,change_case,
 * contains,find}.rs; apps/trivium/src/trivium/trivium_bool.rs:143-227).  A program is that schedule
 * compiled once for a workload shape: per level ONE lwe-linear launch + ONE ks_pbs_batch launch, all
 * intermediate blocks resident in HBM.  Names and shapes: tfhe_rs_string_b200/csrc/programs.hpp
 * ("radix_eq|ne|gt|lt|ge|le|max|min|add|sub|bitand|bitor|bitxor|shl", "radix_scalar_gt|lt|le|ge|eq",
 * "string_eq|ne|starts_with|ends_with|to_uppercase|to_lowercase|contains|find", "trivium").  Inputs and outputs are
 * arrays of big-key LWE blocks (k*N+1 u64 each). */
typedef struct b200tfhe_program b200tfhe_program;
int b200tfhe_program_create(b200tfhe_ctx *ctx, const char *op, const uint64_t *shape, size_t n_shape,
                            b200tfhe_program **out);
/* A caller-built schedule: what the reference's integer layer would hand over instead of one apply_lookup_table call
 * per block.  Blocks 0..n_inputs-1 are the inputs; node j produces block n_inputs + j =
 *     LUT[node_lut[j]]( sum_{t in [node_term_begin[j], node_term_begin[j+1])} term_coeff[t] * block[term_block[t]]
 *                        + node_plaintext[j] * delta )
 * or just the linear combination when node_lut[j] < 0.  Nodes may only reference earlier blocks; the library levels
 * the DAG and runs one lwe-linear launch + one KS+PBS launch per dependency level, everything resident in HBM.
 * luts: n_luts tables of message_modulus*carry_modulus function values (as for b200tfhe_register_lut_from_table).
 * Degree / noise bookkeeping stays with the caller, as for apply_lookup_table itself. */
typedef struct {
    size_t n_inputs, n_nodes, n_luts, n_outputs;
    const uint32_t *node_term_begin;   /* n_nodes + 1 */
    const int32_t *term_block;         /* block ids */
    const int64_t *term_coeff;         /* small signed scalars */
    const uint64_t *node_plaintext;    /* n_nodes message-space constants, or NULL */
    const int32_t *node_lut;           /* n_nodes indices into luts, -1 = linear only */
    const uint64_t *luts;              /* n_luts * message_modulus * carry_modulus */
    const int32_t *outputs;            /* n_outputs block ids */
} b200tfhe_circuit_desc;
int b200tfhe_program_create_from_circuit(b200tfhe_ctx *ctx, const b200tfhe_circuit_desc *desc, b200tfhe_program **out);
/* info[0..5] = n_inputs, n_outputs, n_pbs, depth, n_stages, n_luts */
int b200tfhe_program_info(const b200tfhe_program *prog, uint64_t *info);
int b200tfhe_program_run(b200tfhe_program *prog, const uint64_t *in, uint64_t *out);              /* host buffers, synchronous */
int b200tfhe_program_run_device(b200tfhe_program *prog, const uint64_t *d_in, uint64_t *d_out);   /* device buffers, asynchronous */
int b200tfhe_program_destroy(b200tfhe_program *prog);

/* ---- boolean gates: 32-bit torus words (boolean/engine/{mod,bootstrapping}.rs) -------------------------------- */
/* The reference's boolean layer keeps u32 ciphertexts and bootstraps every binary gate with a constant test polynomial
 * (1/8).  `params`: BooleanParameters (boolean/parameters/mod.rs:123-192; message/carry modulus are ignored);
 * keyswitch_first != 0 for EncryptionKeyChoice::Big sets (ciphertexts of k*N+1 words, keyswitch then bootstrap), 0 for
 * ::Small (n+1 words, bootstrap then keyswitch).  One GPU per context; host buffers; synchronous. */
typedef struct b200tfhe_boolean_ctx b200tfhe_boolean_ctx;
int b200tfhe_boolean_ctx_create(const b200tfhe_params *params, int keyswitch_first, int device, b200tfhe_boolean_ctx **out);
int b200tfhe_boolean_ctx_destroy(b200tfhe_boolean_ctx *ctx);
int b200tfhe_boolean_last_error(const b200tfhe_boolean_ctx *ctx, char *buf, size_t buf_len);
int b200tfhe_boolean_load_ksk(b200tfhe_boolean_ctx *ctx, const uint32_t *ksk, size_t n_u32);
int b200tfhe_boolean_load_bsk_standard(b200tfhe_boolean_ctx *ctx, const uint32_t *bsk, size_t n_u32);
/* gate: 0 AND, 1 NAND, 2 OR, 3 NOR, 4 XOR, 5 XNOR (boolean/engine/mod.rs:606-850), batched over independent pairs. */
int b200tfhe_boolean_gate_batch(b200tfhe_boolean_ctx *ctx, int gate, const uint32_t *a, const uint32_t *b, uint32_t *out, size_t batch);

/* ---- unit-test hooks (exercise exactly the transforms the PBS kernel uses) ------------ */
/* out[i] += a_int[i] (x) b_torus[i] in Z[X]/(X^2048+1); host buffers, count x 2048 u64 each.
 * Mirrors the reference's FFT product test, fft_impl/fft64/math/fft/tests.rs:82-222. */
int b200tfhe_debug_negacyclic_mul(b200tfhe_ctx *ctx, const uint64_t *a_int, const uint64_t *b_torus,
                                  uint64_t *out, size_t count);
/* The production bootstrap kernel selected for `batch`, stopped after the first `steps` CMUX steps (in_small: batch x
 * (steps + 1) words = mask prefix and body).  With steps = 1 this is ONE external product on caller data: both
 * implementations see identical digits, so the GPU and the CPU oracle must agree to FFT rounding (tests assert
 * 2^42); mirrors fft_impl/common.rs:145-304, which tests the bootstrap against its definition. */
int b200tfhe_debug_pbs_steps(b200tfhe_ctx *ctx, const uint64_t *in_small, const uint32_t *lut_id, uint64_t *out,
                             size_t batch, uint32_t steps);
/* from_torus (core_crypto/commons/math/torus/mod.rs:72-78, round-half-even as in fft/x86.rs:864) of n doubles, |x| < 2^37,
 * by the two device routines: out_fp = the FP64-pipe conversion the bootstrap kernels use, out_cvt = x - rint(x), times
 * 2^64, cvt.rni.s64.  Host buffers. */
int b200tfhe_debug_from_torus(b200tfhe_ctx *ctx, const double *x, uint64_t *out_fp, uint64_t *out_cvt, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* B200TFHE_H */
