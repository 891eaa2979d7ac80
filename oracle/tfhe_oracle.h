/*
 * tfhe_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain C++ restatement of the reference's (M-Bln/tfhe-rs-string, a fork of tfhe-rs 0.5.0)
 * CPU algorithm for the shortint keyswitch + programmable-bootstrap hot path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product library (libb200tfhe.so) never links or calls it.
 *
 * Parity status (see DESIGN.md "Oracle"):
 *   - integer stages (decomposer, keyswitch, mod-switch, monomial rotations, sample extract,
 *     LUT fill, trivial PBS) are exact restatements: pinned against the reference's doc-test
 *     vectors (polynomial_algorithms.rs:313,373), its decomposer property tests and, end to end,
 *     the Trivium ECRYPT known-answer vectors (apps/trivium/src/trivium/test.rs).
 *   - the Fourier stage: the reference delegates the complex FFT to the third-party crate
 *     concrete-fft 0.3.0 (absent from /root/reference, plan chosen at run time), so post-PBS
 *     ciphertext *bits* are "parity unpinned"; decrypted values are pinned (KATs above).
 *
 * All citations are relative to /root/reference/tfhe/src/.
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* shortint/parameters/mod.rs:62-76 (ClassicPBSParameters), KS->PBS order, native 2^64 modulus */
typedef struct {
    uint32_t lwe_dimension;      /* n   (small key)            */
    uint32_t glwe_dimension;     /* k                           */
    uint32_t polynomial_size;    /* N                           */
    double lwe_modular_std_dev;  /* noise of KSK rows           */
    double glwe_modular_std_dev; /* noise of BSK rows / fresh big-key encryptions */
    uint32_t pbs_base_log, pbs_level;
    uint32_t ks_base_log, ks_level;
    uint32_t message_modulus, carry_modulus;
} orc_params;

/* PARAM_MESSAGE_2_CARRY_2_KS_PBS, shortint/parameters/mod.rs:703-717 */
void orc_params_message_2_carry_2(orc_params *p);
/* PARAM_MESSAGE_2_CARRY_2_TEST (insecure, fast), examples/fhe_strings/ciphertext.rs:76-90,
 * with lwe_dimension raised from 1 to `n` so the blind-rotation loop is exercised. */
void orc_params_toy(orc_params *p, uint32_t n, uint32_t N);

typedef struct orc_keyset orc_keyset;

/* Client + server key material generated from a seed (xoshiro256**; NOT the reference's
 * AES-CTR CSPRNG, which is client-side and out of scope). */
orc_keyset *orc_keyset_create(const orc_params *p, uint64_t seed, int n_threads);
void orc_keyset_destroy(orc_keyset *ks);
const orc_params *orc_keyset_params(const orc_keyset *ks);
const uint64_t *orc_keyset_small_sk(const orc_keyset *ks);      /* n                     */
const uint64_t *orc_keyset_big_sk(const orc_keyset *ks);        /* k*N                   */
const uint64_t *orc_keyset_ksk(const orc_keyset *ks);           /* [k*N][ks_level][n+1]  */
size_t orc_keyset_ksk_len(const orc_keyset *ks);
const uint64_t *orc_keyset_bsk_standard(const orc_keyset *ks);  /* [n][lvl][k+1][k+1][N] */
size_t orc_keyset_bsk_len(const orc_keyset *ks);
const double *orc_keyset_bsk_fourier(const orc_keyset *ks);     /* [n][lvl][k+1][k+1][N/2] c64 */

/* -- integer primitives ------------------------------------------------------------------ */
uint64_t orc_closest_representable(uint64_t x, uint32_t base_log, uint32_t level);
/* digits[0] is the term of level `level` (least significant), digits[level-1] of level 1;
 * same order as SignedDecompositionIter (iter.rs:101-117). */
void orc_decompose(uint64_t x, uint32_t base_log, uint32_t level, int64_t *digits);
uint64_t orc_modulus_switch(uint64_t x, uint32_t log2_poly_size);
void orc_monomial_div(uint64_t *out, const uint64_t *in, size_t N, size_t degree);
void orc_monomial_mul(uint64_t *out, const uint64_t *in, size_t N, size_t degree);
void orc_monomial_mul_and_subtract(uint64_t *out, const uint64_t *in, size_t N, size_t degree);
void orc_sample_extract0(uint64_t *lwe_out, const uint64_t *glwe, uint32_t k, uint32_t N);

/* -- Fourier primitives ------------------------------------------------------------------- */
/* negacyclic forward transforms: poly (N u64) -> N/2 complex (interleaved re,im) */
void orc_fft_forward_integer(double *fourier, const uint64_t *poly, uint32_t N);
void orc_fft_forward_torus(double *fourier, const uint64_t *poly, uint32_t N);
/* inverse + from_torus rounding, wrapping-added into poly */
void orc_fft_add_backward_torus(uint64_t *poly, const double *fourier, uint32_t N);
uint64_t orc_from_torus(double x);

/* -- hot path ----------------------------------------------------------------------------- */
void orc_keyswitch(const orc_keyset *ks, const uint64_t *in_big, uint64_t *out_small);
/* generic version on raw arrays; ksk layout [in_dim][level][out_dim+1] */
void orc_keyswitch_raw(const uint64_t *ksk, uint32_t in_dim, uint32_t out_dim, uint32_t base_log,
                       uint32_t level, const uint64_t *in, uint64_t *out);
void orc_blind_rotate(const orc_keyset *ks, const uint64_t *lwe_small, uint64_t *acc_glwe);
void orc_bootstrap(const orc_keyset *ks, const uint64_t *lwe_small, const uint64_t *lut_glwe,
                   uint64_t *out_big);
/* test hook: the bootstrap stopped after the first n_steps CMUX steps; lwe_prefix = n_steps mask words then the body */
void orc_bootstrap_steps(const orc_keyset *ks, const uint64_t *lwe_prefix, uint32_t n_steps, const uint64_t *lut_glwe,
                         uint64_t *out_big);
void orc_ks_pbs(const orc_keyset *ks, const uint64_t *in_big, const uint64_t *lut_glwe,
                uint64_t *out_big);
/* luts: n_luts GLWE accumulators ((k+1)*N u64 each); lut_idx[b] selects per ciphertext */
void orc_ks_pbs_batch(const orc_keyset *ks, const uint64_t *in_big, const uint64_t *luts,
                      const uint32_t *lut_idx, uint64_t *out_big, size_t batch, int n_threads);

/* -- shortint layer ----------------------------------------------------------------------- */
/* fill_accumulator: table[i] = f(i), i < message_modulus*carry_modulus; returns max f */
uint64_t orc_fill_accumulator(const orc_params *p, const uint64_t *table, uint64_t *glwe_out);
/* trivial_pbs_assign (shortint/server_key/mod.rs:763-781) on the body word */
uint64_t orc_trivial_pbs(const orc_params *p, uint64_t body, const uint64_t *lut_glwe);
void orc_encrypt(orc_keyset *ks, uint64_t message_and_carry, uint64_t *out_big);
/* same but with an explicit rng stream so tests can reproduce inputs */
void orc_encrypt_seeded(const orc_keyset *ks, uint64_t message_and_carry, uint64_t seed,
                        uint64_t *out_big);
void orc_encrypt_batch_seeded(const orc_keyset *ks, const uint64_t *messages, size_t batch,
                              uint64_t seed, uint64_t *out_big);
uint64_t orc_decrypt_phase(const orc_keyset *ks, const uint64_t *ct_big);
uint64_t orc_decrypt_small_phase(const orc_keyset *ks, const uint64_t *ct_small);
uint64_t orc_decrypt_message_and_carry(const orc_keyset *ks, const uint64_t *ct_big);
void orc_decrypt_batch(const orc_keyset *ks, const uint64_t *cts, size_t batch, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
