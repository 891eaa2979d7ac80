// workloads.hpp -- the reference's batched call sites, re-expressed as leveled circuits.
//
// Everything here is host-side C++ over Circuit (circuit.hpp): radix integer operations
// (integer/server_key/radix_parallel/*), the FheString operations of examples/fhe_strings and
// Trivium (apps/trivium).  Operator decomposition follows the reference (same lookup tables,
// same block packing); what changes is the SCHEDULE: sequential boolean folds become
// sum-of-flags trees and the per-position index accumulation of `find` becomes a carry-save
// column adder, so that a whole batch of strings runs in a handful of KS+PBS launches.
// Paths below are relative to /root/reference/tfhe/.
#pragma once
#include <string>

#include "circuit.hpp"

namespace b200 {
namespace wl {

using Radix = std::vector<Lin>;   // little-endian blocks, message_modulus values per block

inline int lut_message_extract(Circuit &c) { return c.lut([&](uint64_t x) { return x % c.msg_mod; }); }   // shortint/server_key/mod.rs:619-623
inline int lut_carry_extract(Circuit &c) { return c.lut([&](uint64_t x) { return x / c.msg_mod; }); }      // shortint/server_key/mod.rs:539-545

// ---- boolean blocks (integer/server_key/radix/bitwise_op.rs:632-735 -> shortint bitand/bitor/bitxor,
// shortint/server_key/bitwise_op.rs:111-126,204-207: bivariate PBS with factor = rhs.degree + 1 = 2)
inline Lin bool_and(Circuit &c, const Lin &a, const Lin &b) {
    return c.pbs_bivariate(a, b, c.lut_bivariate([](uint64_t x, uint64_t y) { return x & y; }, 2), 2);
}
inline Lin bool_or(Circuit &c, const Lin &a, const Lin &b) {
    return c.pbs_bivariate(a, b, c.lut_bivariate([](uint64_t x, uint64_t y) { return x | y; }, 2), 2);
}
inline Lin bool_xor(Circuit &c, const Lin &a, const Lin &b) {
    return c.pbs_bivariate(a, b, c.lut_bivariate([](uint64_t x, uint64_t y) { return x ^ y; }, 2), 2);
}
inline Lin bool_not(Circuit &c, const Lin &a) { return c.add_const(c.scale(a, -1), 1); }   // boolean_bitnot: 1 - x, leveled

// are_all_comparisons_block_true, src/integer/server_key/radix_parallel/scalar_comparison.rs:147-192.  The reference sums
// modulus_sup - 1 = 15 flags per lookup (x == 15).  A full chunk here takes modulus_sup = 16: the sum 0..16 reaches the
// padding bit exactly at "all true", where the negacyclic lookup returns -LUT(0); with the constant table -1/2 the PBS yields
// -delta/2 below 16 and +delta/2 at 16, and adding delta/2 (leveled) gives the flag.  256 block comparisons reduce in two
// levels (256 -> 16 -> 1) instead of three, which is one bootstrap depth less for every string comparison.
inline Lin all_true(Circuit &c, std::vector<Lin> flags) {
    if (flags.empty()) return c.constant(1);
    const size_t ms = c.modulus_sup();
    while (flags.size() > 1) {
        std::vector<Lin> next;
        for (size_t i = 0; i < flags.size(); i += ms) {
            const size_t n = std::min(ms, flags.size() - i);
            Lin sum = flags[i];
            for (size_t j = 1; j < n; j++) sum = c.add(sum, flags[i + j]);
            if (n == ms)
                next.push_back(c.add_half(c.pbs_unchecked(sum, c.lut_half([](uint64_t) { return (int64_t)-1; }), 1), 1, 1));
            else
                next.push_back(c.pbs(sum, c.lut([n](uint64_t x) { return (uint64_t)(x == n); })));
        }
        flags.swap(next);
    }
    return flags[0];
}
// is_at_least_one_comparisons_block_true, scalar_comparison.rs:194-228; full chunks of 16 with the table
// (-1/2 at 0, +1/2 elsewhere): the sum 16 reads -LUT(0) = +1/2 as it must
inline Lin any_true(Circuit &c, std::vector<Lin> flags) {
    if (flags.empty()) return c.constant(0);
    const size_t ms = c.modulus_sup();
    bool first = true;
    while (flags.size() > 1 || first) {
        first = false;
        std::vector<Lin> next;
        for (size_t i = 0; i < flags.size(); i += ms) {
            const size_t n = std::min(ms, flags.size() - i);
            Lin sum = flags[i];
            for (size_t j = 1; j < n; j++) sum = c.add(sum, flags[i + j]);
            if (n == ms)
                next.push_back(c.add_half(c.pbs_unchecked(sum, c.lut_half([](uint64_t x) { return (int64_t)(x == 0 ? -1 : 1); }), 1), 1, 1));
            else
                next.push_back(n == 1 ? sum : c.pbs(sum, c.lut([](uint64_t x) { return (uint64_t)(x != 0); })));
        }
        flags.swap(next);
    }
    return flags[0];
}

// ---- carries: full_propagate_parallelized, sequential branch (radix_parallel/mod.rs:65-86,109-113):
// per block message_extract and carry_extract of the SAME ciphertext (two LUTs), carry added to the next block
inline void full_propagate(Circuit &c, Radix &x) {
    const int lm = lut_message_extract(c), lc = lut_carry_extract(c);
    for (size_t i = 0; i < x.size(); i++) {
        const Lin cur = x[i];
        if (i + 1 < x.size()) x[i + 1] = c.add(x[i + 1], c.pbs(cur, lc));   // the last carry is dropped by the reference as well
        x[i] = c.pbs(cur, lm);
    }
}
inline bool carries_empty(const Circuit &c, const Radix &x) {
    for (const Lin &b : x)
        if (b.degree >= c.msg_mod) return false;
    return true;
}

// add_assign_parallelized, radix_parallel/add.rs:206-242 (for <= 8 blocks the thread-count heuristic
// :44-76 never picks the prefix-sum variant): unchecked add, then full propagation
inline Radix radix_add(Circuit &c, Radix a, Radix b) {
    if (!carries_empty(c, a)) full_propagate(c, a);
    if (!carries_empty(c, b)) full_propagate(c, b);
    Radix r(a.size());
    for (size_t i = 0; i < a.size(); i++) r[i] = c.add(a[i], b[i]);
    full_propagate(c, r);
    return r;
}

// unchecked_neg_assign, src/integer/server_key/radix/neg.rs:56-73 + shortint/server_key/neg.rs:223-246
inline Radix radix_neg(Circuit &c, Radix x) {
    uint64_t z_b = 0;
    for (Lin &blk : x) {
        if (z_b) blk = c.add_const(blk, (int64_t)z_b);
        uint64_t z = std::max<uint64_t>(1, (blk.degree + c.msg_mod - 1) / c.msg_mod) * c.msg_mod;
        Lin n = c.add_const(c.scale(blk, -1), (int64_t)z);
        n.degree = (uint32_t)(z - z_b);
        blk = n;
        z_b = z / c.msg_mod;
    }
    return x;
}
// sub_parallelized: unchecked_sub (= add of the negation) + full propagation
inline Radix radix_sub(Circuit &c, Radix a, Radix b) {
    if (!carries_empty(c, a)) full_propagate(c, a);
    if (!carries_empty(c, b)) full_propagate(c, b);
    Radix nb = radix_neg(c, b), r(a.size());
    for (size_t i = 0; i < a.size(); i++) r[i] = c.add(a[i], nb[i]);
    full_propagate(c, r);
    return r;
}

// unchecked_eq_parallelized, radix_parallel/comparison.rs:10-33: block `==` flags, then all-true.  The reference spends one
// bivariate lookup per block; with carry_modulus >= message_modulus the flags are formed here per PAIR of blocks with the
// packing of the reference's own comparator (pack_block_chunk + TRUE LWE subtraction, comparator.rs:191-220,430-443): the
// difference of two packed blocks lies in (-modulus_sup, modulus_sup), and the table (x == 0) needs no sign fix-up because it
// is 0 on every non-zero entry, so the negacyclic half (-LUT) reads -0.  Same flags, half the bootstraps; the noise of the
// operand is that of the reference's `gt` / `lt` (two packed blocks).
inline void radix_eq_flags(Circuit &c, const Radix &a, const Radix &b, std::vector<Lin> &out) {
    if (a.size() != b.size()) throw std::invalid_argument("radix_eq_flags: block counts differ");
    if (c.carry_mod >= c.msg_mod) {
        const int lut = c.lut([](uint64_t x) { return (uint64_t)(x == 0); });
        for (size_t i = 0; i < a.size(); i += 2) {
            const bool pair = i + 1 < a.size();
            const Lin pa = pair ? c.axpy(a[i + 1], c.msg_mod, a[i], 1) : a[i];
            const Lin pb = pair ? c.axpy(b[i + 1], c.msg_mod, b[i], 1) : b[i];
            out.push_back(c.pbs_unchecked(c.sub(pa, pb), lut, 1));   // deliberately through the padding bit
        }
        return;
    }
    const int lut = c.lut_bivariate([](uint64_t x, uint64_t y) { return (uint64_t)(x == y); }, c.msg_mod);
    for (size_t i = 0; i < a.size(); i++) out.push_back(c.pbs_bivariate(a[i], b[i], lut, c.msg_mod));
}
inline Lin radix_eq(Circuit &c, const Radix &a, const Radix &b) {
    std::vector<Lin> cmp;
    radix_eq_flags(c, a, b, cmp);
    return all_true(c, cmp);
}
// ne_parallelized: radix_parallel/comparison.rs:39-62 (block `!=` flags, then any-true) = the leveled negation of eq
inline Lin radix_ne(Circuit &c, const Radix &a, const Radix &b) { return bool_not(c, radix_eq(c, a, b)); }

// bitand / bitor / bitxor_parallelized (integer/server_key/radix_parallel/bitwise_op.rs -> per block
// shortint unchecked_bitand/bitor/bitxor, shortint/server_key/bitwise_op.rs:204-207): one bivariate PBS per block
inline Radix radix_bitop(Circuit &c, const Radix &a, const Radix &b, char op) {
    const int lut = c.lut_bivariate([op](uint64_t x, uint64_t y) { return op == '&' ? (x & y) : op == '|' ? (x | y) : (x ^ y); }, c.msg_mod);
    Radix r(a.size());
    for (size_t i = 0; i < a.size(); i++) r[i] = c.pbs_bivariate(a[i], b[i], lut, c.msg_mod);
    return r;
}

// ---- scalar comparisons (src/integer/server_key/comparator.rs)
constexpr uint64_t kInf = 0, kEq = 1, kSup = 2;
inline std::vector<uint8_t> scalar_blocks_early_stop(const Circuit &c, uint64_t s) {   // BlockDecomposer::with_early_stop_at_zero
    std::vector<uint8_t> v;
    while (s) { v.push_back((uint8_t)(s % c.msg_mod)); s /= c.msg_mod; }
    return v;
}
// reduce_signs_parallelized, comparator.rs:240-288 (pairwise, msb*4 + lsb through the reduction table :82-101)
inline Lin reduce_signs(Circuit &c, std::vector<Lin> signs) {
    const int lut = c.lut([](uint64_t x) {
        static const uint64_t t[11] = {kInf, kInf, kInf, kInf, kInf, kEq, kSup, kSup, kSup, kSup, kSup};
        return x < 11 ? t[x] : 0;
    });
    while (signs.size() > 1) {
        std::vector<Lin> next;
        for (size_t i = 0; i + 1 < signs.size(); i += 2) next.push_back(c.pbs(c.axpy(signs[i + 1], 4, signs[i], 1), lut));
        if (signs.size() % 2) next.push_back(signs.back());
        signs.swap(next);
    }
    return signs[0];
}
// compare_blocks_with_zero(Equality), radix_parallel/scalar_comparison.rs:16-110: pack pairs, sum up to
// (total_modulus-1)/(packed max) of them, one PBS x == 0 per chunk
inline std::vector<Lin> blocks_eq_zero(Circuit &c, const std::vector<Lin> &blocks) {
    std::vector<Lin> packed;
    for (size_t i = 0; i < blocks.size(); i += 2)
        packed.push_back(i + 1 < blocks.size() ? c.axpy(blocks[i + 1], c.msg_mod, blocks[i], 1) : blocks[i]);
    const uint64_t per_packed_max = (uint64_t)c.msg_mod * c.msg_mod - 1;
    const size_t chunk = std::max<uint64_t>(1, (c.modulus_sup() - 1) / per_packed_max);
    const int lut = c.lut([](uint64_t x) { return (uint64_t)(x == 0); });
    std::vector<Lin> out;
    for (size_t i = 0; i < packed.size(); i += chunk) {
        Lin s = packed[i];
        for (size_t j = 1; j < chunk && i + j < packed.size(); j++) s = c.add(s, packed[i + j]);
        out.push_back(c.pbs(s, lut));
    }
    return out;
}
// unsigned_unchecked_scalar_compare_blocks_parallelized, comparator.rs:677-783 -> sign block in {0,1,2}
inline Lin scalar_sign(Circuit &c, const Radix &a, uint64_t scalar) {
    std::vector<uint8_t> sb = scalar_blocks_early_stop(c, scalar);
    for (size_t i = a.size(); i < sb.size(); i++)
        if (sb[i]) return c.constant(kInf);   // scalar obviously bigger
    if (sb.size() > a.size()) sb.resize(a.size());
    const size_t n_lsb = sb.size();
    bool has_lsb = n_lsb > 0, has_msb = n_lsb < a.size();
    Lin lsb_sign, msb_sign;
    if (has_lsb) {
        // unchecked_scalar_block_slice_compare_parallelized, comparator.rs:474-498 + scalar_compare_block_assign :222-238
        const int sign_lut = c.lut([](uint64_t x) { return (uint64_t)(x != 0); });
        std::vector<Lin> signs;
        for (size_t i = 0; i < n_lsb; i += 2) {
            const bool pair = i + 1 < n_lsb;
            Lin packed = pair ? c.axpy(a[i + 1], c.msg_mod, a[i], 1) : a[i];
            const int64_t ps = sb[i] + (pair ? sb[i + 1] * (int64_t)c.msg_mod : 0);
            // lhs - scalar: negative results set the padding bit, the sign table then yields -1 (mod), so +1 gives 0/1/2
            signs.push_back(c.add_const(c.pbs(c.add_const(packed, -ps), sign_lut), 1));
        }
        lsb_sign = reduce_signs(c, signs);
    }
    if (has_msb) {
        std::vector<Lin> msb(a.begin() + n_lsb, a.end());
        Lin all_zero = all_true(c, blocks_eq_zero(c, msb));
        msb_sign = c.pbs(all_zero, c.lut([](uint64_t x) { return x == 1 ? kEq : kSup; }));
    }
    if (has_lsb && has_msb) return reduce_signs(c, {lsb_sign, msb_sign});
    return has_lsb ? lsb_sign : msb_sign;
}
// map_sign_result, comparator.rs:957-971,1640-1664
inline Lin scalar_cmp(Circuit &c, const Radix &a, uint64_t scalar, const std::function<bool(uint64_t)> &pred) {
    Lin sign = scalar_sign(c, a, scalar);
    return c.pbs(sign, c.lut([pred](uint64_t x) { return (uint64_t)pred(x); }));
}
inline Lin scalar_gt(Circuit &c, const Radix &a, uint64_t s) { return scalar_cmp(c, a, s, [](uint64_t x) { return x == kSup; }); }
inline Lin scalar_lt(Circuit &c, const Radix &a, uint64_t s) { return scalar_cmp(c, a, s, [](uint64_t x) { return x == kInf; }); }
inline Lin scalar_le(Circuit &c, const Radix &a, uint64_t s) { return scalar_cmp(c, a, s, [](uint64_t x) { return x != kSup; }); }
inline Lin scalar_ge(Circuit &c, const Radix &a, uint64_t s) { return scalar_cmp(c, a, s, [](uint64_t x) { return x != kInf; }); }
// ---- encrypted / encrypted ordering (integer/server_key/comparator.rs)
// unchecked_compare_parallelized for unsigned radix integers, comparator.rs:383-463: blocks packed in pairs
// (pack_block_chunk, needs carry_modulus >= message_modulus), compare_block_assign :191-220 = TRUE LWE subtraction
// (a negative difference sets the padding bit, the sign table then yields -1 mod), sign table, +1 -> {0, 1, 2};
// then reduce_signs_parallelized
inline Lin radix_sign(Circuit &c, const Radix &a, const Radix &b) {
    if (a.size() != b.size() || a.empty()) throw std::invalid_argument("radix_sign: operands need the same, non-zero block count");
    if (c.carry_mod < c.msg_mod) throw std::invalid_argument("radix_sign: block packing needs carry_modulus >= message_modulus");
    const int sign_lut = c.lut([](uint64_t x) { return (uint64_t)(x != 0); });
    std::vector<Lin> signs;
    for (size_t i = 0; i < a.size(); i += 2) {
        const bool pair = i + 1 < a.size();
        const Lin pa = pair ? c.axpy(a[i + 1], c.msg_mod, a[i], 1) : a[i];
        const Lin pb = pair ? c.axpy(b[i + 1], c.msg_mod, b[i], 1) : b[i];
        // the difference lies in (-modulus_sup, modulus_sup): deliberately through the padding bit, hence pbs_unchecked
        signs.push_back(c.add_const(c.pbs_unchecked(c.sub(pa, pb), sign_lut, 1), 1));
        signs.back().degree = 2;
    }
    return reduce_signs(c, signs);
}
// map_sign_result (comparator.rs:957-971): gt / lt / ge / le as one more lookup on the sign block
inline Lin radix_cmp(Circuit &c, const Radix &a, const Radix &b, const std::function<bool(uint64_t)> &pred) {
    return c.pbs(radix_sign(c, a, b), c.lut([pred](uint64_t x) { return (uint64_t)pred(x); }));
}
inline Lin radix_gt(Circuit &c, const Radix &a, const Radix &b) { return radix_cmp(c, a, b, [](uint64_t x) { return x == kSup; }); }
inline Lin radix_lt(Circuit &c, const Radix &a, const Radix &b) { return radix_cmp(c, a, b, [](uint64_t x) { return x == kInf; }); }
inline Lin radix_ge(Circuit &c, const Radix &a, const Radix &b) { return radix_cmp(c, a, b, [](uint64_t x) { return x != kInf; }); }
inline Lin radix_le(Circuit &c, const Radix &a, const Radix &b) { return radix_cmp(c, a, b, [](uint64_t x) { return x != kSup; }); }
// unchecked_min_or_max_parallelized, comparator.rs:849-875 -> unchecked_programmable_if_then_else_parallelized
// (radix_parallel/cmux.rs:194-248): both operands are zeroed block by block with a bivariate lookup on (block, sign)
// under the predicate / its negation, added, and cleaned with message_extract
inline Radix radix_min_max(Circuit &c, const Radix &a, const Radix &b, bool want_max) {
    const Lin sign = radix_sign(c, a, b);
    const uint32_t factor = 3;   // sign block in {0, 1, 2}: factor = degree + 1
    const uint64_t keep_a = want_max ? kSup : kInf;   // a is kept when it is the strict winner, b otherwise (ties: b = a)
    const int lut_a = c.lut_bivariate([keep_a](uint64_t blk, uint64_t sg) { return sg == keep_a ? blk : (uint64_t)0; }, factor);
    const int lut_b = c.lut_bivariate([keep_a](uint64_t blk, uint64_t sg) { return sg == keep_a ? (uint64_t)0 : blk; }, factor);
    const int clean = lut_message_extract(c);
    Radix r(a.size());
    for (size_t i = 0; i < a.size(); i++) {
        const Lin ta = c.pbs_bivariate(a[i], sign, lut_a, factor), tb = c.pbs_bivariate(b[i], sign, lut_b, factor);
        r[i] = c.pbs(c.add(ta, tb), clean);
    }
    return r;
}

// unchecked_scalar_eq_parallelized, radix_parallel/scalar_comparison.rs:230-330: per pair of blocks one PBS
// "packed == packed scalar", then all-true
inline Lin scalar_eq(Circuit &c, const Radix &a, uint64_t scalar) {
    std::vector<Lin> cmp;
    uint64_t s = scalar;
    for (size_t i = 0; i < a.size(); i += 2) {
        const bool pair = i + 1 < a.size();
        const uint64_t d0 = s % c.msg_mod; s /= c.msg_mod;
        const uint64_t d1 = pair ? s % c.msg_mod : 0; if (pair) s /= c.msg_mod;
        const uint64_t ps = d0 + d1 * c.msg_mod;
        Lin packed = pair ? c.axpy(a[i + 1], c.msg_mod, a[i], 1) : a[i];
        cmp.push_back(c.pbs(packed, c.lut([ps](uint64_t x) { return (uint64_t)(x == ps); })));
    }
    if (s != 0) return c.constant(0);
    return all_true(c, cmp);
}

// scalar_mul by a power of two = scalar_left_shift (radix_parallel/scalar_mul.rs:346-351 ->
// scalar_shift.rs): whole-block rotation plus one bivariate PBS per block for the sub-block bits
inline Radix scalar_left_shift(Circuit &c, const Radix &a, unsigned bits) {
    unsigned log_mod = 0;
    while ((1u << log_mod) < c.msg_mod) log_mod++;
    const size_t rot = bits / log_mod;
    const unsigned s = bits % log_mod;
    const size_t n = a.size();
    Radix r(n, c.constant(0));
    for (size_t i = rot; i < n; i++) r[i] = a[i - rot];
    if (s == 0) return r;
    const uint32_t mm = c.msg_mod;
    const int lut = c.lut_bivariate([=](uint64_t cur, uint64_t prev) { return ((cur << s) % mm) | (prev >> (log_mod - s)); }, mm);
    Radix out(n);
    for (size_t i = 0; i < n; i++) {
        const Lin prev = i > 0 ? r[i - 1] : c.constant(0);
        out[i] = c.pbs_bivariate(r[i], prev, lut, mm);
    }
    return out;
}
inline Radix bool_to_radix(Circuit &c, const Lin &b, size_t n_blocks) {   // examples/fhe_strings: bool_to_radix
    Radix r(n_blocks, c.constant(0));
    r[0] = b;
    return r;
}

// ---- FheString operations (examples/fhe_strings/server_key/*)
using FheChars = std::vector<Radix>;   // one radix (4 blocks) per character, no padding, clear length

// to_uppercase_char, change_case.rs:53-67: (c > 96 & c < 123) -> c - 32 * flag; the reference's own decomposition
// (two scalar comparisons, a boolean and, a shift and a subtraction with carry propagation: 18 bootstraps, depth 7)
inline Radix to_uppercase_char_reference(Circuit &c, const Radix &ch) {
    Lin flag = bool_and(c, scalar_gt(c, ch, 96), scalar_lt(c, ch, 123));
    Radix delta = scalar_left_shift(c, bool_to_radix(c, flag, ch.size()), 5);   // scalar_mul_parallelized(.., 32)
    return radix_sub(c, ch, delta);
}
// to_lowercase_char, change_case.rs:69-82: (c > 64 & c < 91) -> c + 32 * flag
inline Radix to_lowercase_char_reference(Circuit &c, const Radix &ch) {
    Lin flag = bool_and(c, scalar_gt(c, ch, 64), scalar_lt(c, ch, 91));
    Radix delta = scalar_left_shift(c, bool_to_radix(c, flag, ch.size()), 5);
    return radix_add(c, ch, delta);
}
// The same function on an 8-bit character held in four 2-bit blocks, scheduled for the batch (3 bootstraps, depth 2).
// With hi = packed blocks 3,2 and lo = packed blocks 1,0 (pack_block_chunk), the letters to change are hi = h0 with
// lo >= 1 or hi = h0 + 1 with lo <= 10 (h0 = 6 for a..z, 4 for A..Z), and adding / subtracting 32 only changes block 2
// by 2, never with a carry or a borrow (block 2 is 2 or 3 for a..z, 0 or 1 for A..Z).  So: one lookup on hi (which row),
// one on lo (which of the two bounds hold), one on their packing (2 * flag), and a leveled +- on block 2.  Decrypted
// results are those of the reference's decomposition for all 256 byte values (tests); the other three blocks are the inputs.
inline Radix case_change_char(Circuit &c, const Radix &ch, bool to_upper) {
    if (!(c.msg_mod == 4 && c.carry_mod >= 4 && ch.size() == 4))
        return to_upper ? to_uppercase_char_reference(c, ch) : to_lowercase_char_reference(c, ch);
    const uint64_t h0 = to_upper ? 6 : 4;
    const Lin hi = c.axpy(ch[3], 4, ch[2], 1), lo = c.axpy(ch[1], 4, ch[0], 1);
    const Lin row = c.pbs(hi, c.lut([h0](uint64_t x) { return x == h0 ? (uint64_t)1 : x == h0 + 1 ? (uint64_t)2 : (uint64_t)0; }));
    const Lin bounds = c.pbs(lo, c.lut([](uint64_t x) { return (uint64_t)(x >= 1) + 2 * (uint64_t)(x <= 10); }));
    const Lin twice_flag = c.pbs(c.axpy(row, 4, bounds, 1), c.lut([](uint64_t x) {
        const uint64_t r = x / 4, b = x % 4;
        return (uint64_t)(((r == 1 && (b & 1)) || (r == 2 && (b & 2))) ? 2 : 0);
    }));
    Radix out = ch;
    out[2] = to_upper ? c.sub(ch[2], twice_flag) : c.add(ch[2], twice_flag);
    out[2].degree = c.msg_mod - 1;   // the value stays a clean 2-bit block (no borrow / carry by construction)
    return out;
}
inline Radix to_uppercase_char(Circuit &c, const Radix &ch) { return case_change_char(c, ch, true); }
inline Radix to_lowercase_char(Circuit &c, const Radix &ch) { return case_change_char(c, ch, false); }
// eq_no_init_padding for two unpadded strings (comparisons.rs:184-215); the serial `&=` fold of the
// reference becomes one sum-of-flags tree over all character comparisons
inline Lin string_eq(Circuit &c, const FheChars &a, const FheChars &b) {
    const size_t n = std::min(a.size(), b.size());
    std::vector<Lin> flags;   // block comparisons of ALL characters in one reduction tree
    for (size_t i = 0; i < n; i++) radix_eq_flags(c, a[i], b[i], flags);
    if (a.size() != b.size()) {   // the longer string must continue with a padding zero (:195-212)
        const FheChars &longer = a.size() > b.size() ? a : b;
        flags.push_back(scalar_eq(c, longer[n], 0));
    }
    return all_true(c, flags);
}
// is_prefix_of_slice with Padding::None (pattern.rs:246-274): AND of the character comparisons
inline Lin match_at(Circuit &c, const FheChars &hay, const FheChars &pat, size_t pos) {
    if (pat.size() > hay.size() - pos) return c.constant(0);
    std::vector<Lin> flags;
    for (size_t j = 0; j < pat.size(); j++) radix_eq_flags(c, hay[pos + j], pat[j], flags);
    return all_true(c, flags);
}
// starts_with_encrypted_vec with Padding::None (contains.rs:96-134): the overlapping characters agree and,
// when the prefix is longer than the string, its next character is a padding zero
inline Lin string_starts_with(Circuit &c, const FheChars &s, const FheChars &prefix) {
    std::vector<Lin> flags;
    for (size_t j = 0; j < std::min(s.size(), prefix.size()); j++) radix_eq_flags(c, s[j], prefix[j], flags);
    if (prefix.size() > s.size()) flags.push_back(scalar_eq(c, prefix[s.size()], 0));
    return all_true(c, flags);
}
// is_suffix_of_string for two unpadded strings (pattern.rs: the pattern must match at position len - pat_len)
inline Lin string_ends_with(Circuit &c, const FheChars &s, const FheChars &suffix) {
    if (suffix.size() > s.size()) return c.constant(0);
    return match_at(c, s, suffix, s.size() - suffix.size());
}
// contains_unpadded_string, contains.rs:77-92: OR over all start positions
inline Lin string_contains(Circuit &c, const FheChars &hay, const FheChars &pat) {
    if (hay.empty()) return pat.empty() ? c.constant(1) : scalar_eq(c, pat[0], 0);
    std::vector<Lin> m;
    for (size_t n = 0; n < hay.size(); n++) {
        Lin f = match_at(c, hay, pat, n);
        if (!f.is_const() || f.cst) m.push_back(f);
    }
    return any_true(c, m);
}
// prefix OR of boolean flags with depth O(log_15): found[n] = OR_{m<=n} flags[m]
inline std::vector<Lin> prefix_or(Circuit &c, const std::vector<Lin> &flags) {
    const size_t G = c.modulus_sup() - 1;
    const size_t n = flags.size();
    if (n == 0) return {};
    const int nz = c.lut([](uint64_t x) { return (uint64_t)(x != 0); });
    std::vector<Lin> within(n), group_any;
    for (size_t g = 0; g * G < n; g++) {
        Lin run = c.constant(0);
        for (size_t i = g * G; i < std::min(n, (g + 1) * G); i++) {
            run = c.add(run, flags[i]);
            within[i] = (i == g * G) ? run : c.pbs(run, nz);
        }
        group_any.push_back(within[std::min(n, (g + 1) * G) - 1]);
    }
    if (group_any.size() == 1) return within;
    std::vector<Lin> gp = prefix_or(c, group_any);
    std::vector<Lin> out(n);
    for (size_t i = 0; i < n; i++) {
        const size_t g = i / G;
        out[i] = g == 0 ? within[i] : c.pbs(c.add(within[i], gp[g - 1]), nz);
    }
    return out;
}
// Sum of boolean flags as a radix integer modulo msg_mod^n_blocks: carry-save column compression
// (replaces the per-position add_assign_parallelized chain of find.rs:150-160)
inline Radix sum_flags(Circuit &c, const std::vector<Lin> &flags, size_t n_blocks) {
    const int lm = lut_message_extract(c), lc = lut_carry_extract(c);
    std::vector<std::vector<Lin>> col(n_blocks);
    col[0] = flags;
    const uint32_t cap = c.modulus_sup() - 1;
    for (;;) {
        bool busy = false;
        std::vector<std::vector<Lin>> next(n_blocks);
        for (size_t k = 0; k < n_blocks; k++) {
            if (col[k].size() <= 1) { for (auto &v : col[k]) next[k].push_back(v); continue; }
            busy = true;
            size_t i = 0;
            while (i < col[k].size()) {
                Lin s = col[k][i++];
                while (i < col[k].size() && s.degree + col[k][i].degree <= cap) s = c.add(s, col[k][i++]);
                if (s.degree < c.msg_mod) { next[k].push_back(s); continue; }
                next[k].push_back(c.pbs(s, lm));
                if (k + 1 < n_blocks) next[k + 1].push_back(c.pbs(s, lc));
            }
        }
        col.swap(next);
        if (!busy) break;
    }
    Radix r(n_blocks, c.constant(0));
    for (size_t k = 0; k < n_blocks; k++)
        if (!col[k].empty()) r[k] = col[k][0];
    return r;
}
// connected_find_unpadded_string, find.rs:139-160: (found, index) with index = number of positions
// before the first match (= haystack length when there is none), as an n_blocks radix
inline std::pair<Lin, Radix> string_find(Circuit &c, const FheChars &hay, const FheChars &pat, size_t n_blocks) {
    std::vector<Lin> m;
    for (size_t n = 0; n < hay.size(); n++) m.push_back(match_at(c, hay, pat, n));
    std::vector<Lin> found = prefix_or(c, m);
    std::vector<Lin> not_found;
    for (const Lin &f : found) not_found.push_back(bool_not(c, f));   // increment_index, find.rs:594-605
    Radix index = sum_flags(c, not_found, n_blocks);
    return {found.empty() ? c.constant(0) : found.back(), index};
}

// ---- Trivium (apps/trivium/src/trivium/trivium_bool.rs): 64 steps per round, each step 3 AND + 11 XOR
struct Trivium {
    std::vector<Lin> a, b, c;   // registers, index 0 = most recently pushed bit (StaticDeque order)
};
inline Trivium trivium_init(Circuit &cir, const std::vector<Lin> &key80, const std::vector<bool> &iv80) {
    // TriviumStream::new, trivium_bool.rs:63-85: a <- key, b <- iv, c <- 111 with three ones
    Trivium t;
    t.a.assign(93, cir.constant(0)); t.b.assign(84, cir.constant(0)); t.c.assign(111, cir.constant(0));
    for (int i = 0; i < 80; i++) {
        t.a[93 - 80 + i] = key80[i];
        t.b[84 - 80 + i] = cir.constant(iv80[i] ? 1 : 0);
    }
    t.c[0] = t.c[1] = t.c[2] = cir.constant(1);
    // StaticDeque indexing: reg[k] is the k-th most recent element; the constructor array is stored oldest-first,
    // so element j of the array is reg[len-1-j]
    std::reverse(t.a.begin(), t.a.end()); std::reverse(t.b.begin(), t.b.end()); std::reverse(t.c.begin(), t.c.end());
    return t;
}
// next_64, trivium_bool.rs:143-227; returns the 64 output bits, oldest first
inline std::vector<Lin> trivium_next64(Circuit &cir, Trivium &t) {
    auto X = [&](const Lin &x, const Lin &y) { return bool_xor(cir, x, y); };
    auto A = [&](const Lin &x, const Lin &y) { return bool_and(cir, x, y); };
    std::vector<Lin> out(64), na(64), nb(64), nc(64);
    for (int n = 0; n < 64; n++) {
        const Lin ta = X(t.a[65 - n], t.a[92 - n]), tb = X(t.b[68 - n], t.b[83 - n]), tc = X(t.c[65 - n], t.c[110 - n]);
        const Lin a_and = A(t.a[91 - n], t.a[90 - n]), b_and = A(t.b[82 - n], t.b[81 - n]), c_and = A(t.c[109 - n], t.c[108 - n]);
        out[n] = X(X(ta, tb), tc);
        na[n] = X(tc, X(c_and, t.a[68 - n]));
        nb[n] = X(ta, X(a_and, t.b[77 - n]));
        nc[n] = X(tb, X(b_and, t.c[86 - n]));
    }
    // values.pop() order: step 0 first; each push makes the new bit index 0
    for (int n = 0; n < 64; n++) {
        t.a.insert(t.a.begin(), na[n]); t.a.pop_back();
        t.b.insert(t.b.begin(), nb[n]); t.b.pop_back();
        t.c.insert(t.c.begin(), nc[n]); t.c.pop_back();
    }
    return out;
}

}  // namespace wl
}  // namespace b200
