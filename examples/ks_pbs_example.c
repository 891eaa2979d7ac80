/* Minimal C client of libb200tfhe.so: the calls a non-Rust host makes for one batched apply_lookup_table.
 * Build: gcc -std=c99 -Iinclude examples/ks_pbs_example.c -Ltfhe_rs_string_b200 -lb200tfhe -o ks_pbs_example
 * (keys here are random words: the point is the call sequence and error handling, not decryption). */
#include <stdio.h>
#include <stdlib.h>

#include "b200tfhe.h"

static int fail(b200tfhe_ctx *ctx, const char *what) {
    char msg[512];
    if (ctx) b200tfhe_last_error(ctx, msg, sizeof msg); else b200tfhe_last_global_error(msg, sizeof msg);
    fprintf(stderr, "%s: %s\n", what, msg);
    if (ctx) b200tfhe_ctx_destroy(ctx);
    return 1;
}

int main(void) {
    const b200tfhe_params p = {742, 1, 2048, 23, 1, 3, 5, 4, 4};   /* PARAM_MESSAGE_2_CARRY_2_KS_PBS */
    const size_t big = 2049, small = 743, batch = 8;
    const size_t ksk_len = 2048 * 5 * small, bsk_len = (size_t)742 * 4 * 2048;
    b200tfhe_ctx *ctx = NULL;
    if (b200tfhe_ctx_create(&p, 0, &ctx)) return fail(NULL, "ctx_create");   /* no GPU -> error, never a CPU fallback */

    uint64_t *ksk = malloc(ksk_len * 8), *bsk = malloc(bsk_len * 8), *cts = malloc(batch * big * 8);
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < ksk_len; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; ksk[i] = x; }
    for (size_t i = 0; i < bsk_len; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; bsk[i] = x; }
    for (size_t i = 0; i < batch * big; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; cts[i] = x; }
    if (b200tfhe_load_ksk(ctx, ksk, ksk_len)) return fail(ctx, "load_ksk");
    if (b200tfhe_load_bsk_standard(ctx, bsk, bsk_len)) return fail(ctx, "load_bsk_standard");

    const uint64_t table[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};   /* identity */
    uint32_t lut = 0, ids[8];
    if (b200tfhe_register_lut_from_table(ctx, table, 16, &lut)) return fail(ctx, "register_lut");
    for (size_t i = 0; i < batch; i++) ids[i] = lut;
    if (b200tfhe_ks_pbs_batch(ctx, cts, ids, cts, batch)) return fail(ctx, "ks_pbs_batch");   /* in place */
    printf("bootstrapped %zu ciphertexts, first body word %016llx\n", batch, (unsigned long long)cts[big - 1]);
    free(ksk); free(bsk); free(cts);
    return b200tfhe_ctx_destroy(ctx);
}
