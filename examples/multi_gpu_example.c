/* C client of the multi-GPU scheduler: one context over N GPUs of the box (argv[1], default 2), one batch call.
 * Build: gcc -std=c99 -Iinclude examples/multi_gpu_example.c -Ltfhe_rs_string_b200 -lb200tfhe -o multi_gpu_example
 * The library uploads and converts the keys on the first GPU, copies them GPU to GPU, and cuts every batch into one
 * contiguous shard per GPU (keys here are random words: the point is the call sequence, not decryption). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200tfhe.h"

static int fail(b200tfhe_ctx *ctx, const char *what) {
    char msg[512];
    if (ctx) b200tfhe_last_error(ctx, msg, sizeof msg); else b200tfhe_last_global_error(msg, sizeof msg);
    fprintf(stderr, "%s: %s\n", what, msg);
    if (ctx) b200tfhe_ctx_destroy(ctx);
    return 1;
}

int main(int argc, char **argv) {
    const b200tfhe_params p = {742, 1, 2048, 23, 1, 3, 5, 4, 4};   /* PARAM_MESSAGE_2_CARRY_2_KS_PBS */
    const size_t big = 2049, small = 743, batch = 1500;
    const size_t ksk_len = 2048 * 5 * small, bsk_len = (size_t)742 * 4 * 2048;
    int n = argc > 1 ? atoi(argv[1]) : 2, devices[8], got = 0;
    if (n < 1 || n > 8) n = 2;
    for (int i = 0; i < n; i++) devices[i] = i;
    b200tfhe_ctx *ctx = NULL;
    if (b200tfhe_ctx_create_multi(&p, devices, n, &ctx)) return fail(NULL, "ctx_create_multi");
    if (b200tfhe_ctx_device_count(ctx, &got) || got != n) return fail(ctx, "ctx_device_count");

    uint64_t *ksk = malloc(ksk_len * 8), *bsk = malloc(bsk_len * 8), *cts = malloc(batch * big * 8), *ref = malloc(batch * big * 8);
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < ksk_len; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; ksk[i] = x; }
    for (size_t i = 0; i < bsk_len; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; bsk[i] = x; }
    for (size_t i = 0; i < batch * big; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; cts[i] = x; }
    if (b200tfhe_load_ksk(ctx, ksk, ksk_len)) return fail(ctx, "load_ksk");
    if (b200tfhe_load_bsk_standard(ctx, bsk, bsk_len)) return fail(ctx, "load_bsk_standard");
    const uint64_t table[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
    uint32_t lut = 0;
    if (b200tfhe_register_lut_from_table(ctx, table, 16, &lut)) return fail(ctx, "register_lut");
    uint32_t *ids = malloc(batch * sizeof *ids);
    for (size_t i = 0; i < batch; i++) ids[i] = lut;
    if (b200tfhe_ks_pbs_batch(ctx, cts, ids, ref, batch)) return fail(ctx, "ks_pbs_batch");
    /* the keyswitch is exact integer arithmetic: every GPU must give the same words for the same ciphertext */
    uint64_t *ks_a = malloc(batch * small * 8), *ks_b = malloc(batch * small * 8);
    if (b200tfhe_keyswitch_batch(ctx, cts, ks_a, batch)) return fail(ctx, "keyswitch_batch");
    memcpy(cts + (batch - 1) * big, cts, big * 8);              /* last ciphertext (last GPU's shard) := first ciphertext */
    if (b200tfhe_keyswitch_batch(ctx, cts, ks_b, batch)) return fail(ctx, "keyswitch_batch");
    if (memcmp(ks_b, ks_b + (batch - 1) * small, small * 8) != 0 || memcmp(ks_a, ks_b, small * 8) != 0) {
        fprintf(stderr, "GPUs disagree on the keyswitch of the same ciphertext\n");
        return 1;
    }
    uint64_t launches = 0;
    b200tfhe_kernel_launch_count(ctx, &launches);
    printf("%d GPU(s): bootstrapped %zu ciphertexts in contiguous shards, %llu kernel launches, first body word %016llx\n", n, batch,
           (unsigned long long)launches, (unsigned long long)ref[big - 1]);
    free(ksk); free(bsk); free(cts); free(ref); free(ids); free(ks_a); free(ks_b);
    return b200tfhe_ctx_destroy(ctx);
}
