// programs.hpp -- named batch programs (the BASELINE.json workload shapes) built from workloads.hpp.
// Shared by the CUDA executor (product) and the CPU test executors (oracle/); host-only code.
#pragma once
#include <memory>
#include <string>

#include "workloads.hpp"

namespace b200 {

// I/O convention: all inputs / outputs are arrays of big-key LWE blocks in the order stated below.
//   radix_eq        shape {n_ints, n_blocks}            in: lhs[n_ints][n_blocks], rhs[n_ints][n_blocks]     out: n_ints booleans
//   radix_add       shape {n_ints, n_blocks}            in: lhs, rhs                                         out: sum[n_ints][n_blocks]
//   radix_sub       shape {n_ints, n_blocks}            in: lhs, rhs                                         out: diff[n_ints][n_blocks]
//   radix_ne        shape {n_ints, n_blocks}            in: lhs, rhs                                         out: n_ints booleans
//   radix_bitand / radix_bitor / radix_bitxor
//                   shape {n_ints, n_blocks}            in: lhs, rhs                                         out: result[n_ints][n_blocks]
//   radix_gt / radix_lt / radix_ge / radix_le (encrypted vs encrypted, unsigned)
//                   shape {n_ints, n_blocks}            in: lhs, rhs                                         out: n_ints booleans
//   radix_max / radix_min
//                   shape {n_ints, n_blocks}            in: lhs, rhs                                         out: result[n_ints][n_blocks]
//   radix_shl       shape {n_ints, n_blocks, bits}      in: lhs                                              out: (lhs << bits)[n_ints][n_blocks]
//   radix_scalar_gt / radix_scalar_lt / radix_scalar_le / radix_scalar_ge / radix_scalar_eq
//                   shape {n_ints, n_blocks, scalar}    in: lhs                                              out: n_ints booleans
//   string_eq       shape {n_str, len_a, len_b, nb}     in: a[n_str][len_a][nb], b[n_str][len_b][nb]         out: n_str booleans
//   string_ne / string_starts_with / string_ends_with: same shape and inputs as string_eq (b is the pattern)
//   string_to_lowercase: as string_to_uppercase; string_to_uppercase_reference / string_to_lowercase_reference: the same
//                   functions through the reference's own operator decomposition (18 bootstraps per character instead of 3)
//   string_to_uppercase shape {n_str, len, nb}          in: s[n_str][len][nb]                                out: s'[n_str][len][nb]
//   string_contains shape {n_str, hay_len, pat_len, nb} in: hay[n_str][hay_len][nb], pat[n_str][pat_len][nb] out: n_str booleans
//   string_find     shape {n_str, hay_len, pat_len, nb} in: hay, pat                                         out: per string: found, index[nb]
//   trivium         shape {n_rounds, iv_lo, iv_hi}      in: 80 key bits (booleans)                           out: 64*n_rounds keystream bits
//                   (iv bit i = bit i of iv_lo for i < 64, bit i-64 of iv_hi otherwise; 18 warm-up rounds are part of the program)
inline std::unique_ptr<Circuit> build_program(const std::string &op, const std::vector<uint64_t> &shape,
                                              uint32_t msg_mod, uint32_t carry_mod) {
    using namespace wl;
    auto need = [&](size_t n) {
        if (shape.size() != n) throw std::invalid_argument("program '" + op + "': wrong shape length");
    };
    std::unique_ptr<Circuit> cp;
    auto radix_at = [&](Circuit &c, size_t base, size_t nb) {
        Radix r(nb);
        for (size_t k = 0; k < nb; k++) r[k] = c.input(base + k);
        return r;
    };
    if (op == "radix_eq" || op == "radix_ne" || op == "radix_add" || op == "radix_sub" || op == "radix_bitand" ||
        op == "radix_bitor" || op == "radix_bitxor" || op == "radix_gt" || op == "radix_lt" || op == "radix_ge" ||
        op == "radix_le" || op == "radix_max" || op == "radix_min") {
        need(2);
        const size_t n = shape[0], nb = shape[1];
        cp.reset(new Circuit(msg_mod, carry_mod, 2 * n * nb));
        Circuit &c = *cp;
        for (size_t i = 0; i < n; i++) {
            Radix a = radix_at(c, i * nb, nb), b = radix_at(c, (n + i) * nb, nb);
            if (op == "radix_eq") c.output(radix_eq(c, a, b));
            else if (op == "radix_ne") c.output(radix_ne(c, a, b));
            else if (op == "radix_gt") c.output(radix_gt(c, a, b));
            else if (op == "radix_lt") c.output(radix_lt(c, a, b));
            else if (op == "radix_ge") c.output(radix_ge(c, a, b));
            else if (op == "radix_le") c.output(radix_le(c, a, b));
            else if (op == "radix_max" || op == "radix_min") {
                for (const Lin &blk : radix_min_max(c, a, b, op == "radix_max")) c.output(blk);
            } else {
                Radix r = op == "radix_add" ? radix_add(c, a, b) : op == "radix_sub" ? radix_sub(c, a, b)
                        : radix_bitop(c, a, b, op == "radix_bitand" ? '&' : op == "radix_bitor" ? '|' : '^');
                for (const Lin &blk : r) c.output(blk);
            }
        }
    } else if (op == "radix_shl") {
        need(3);
        const size_t n = shape[0], nb = shape[1];
        cp.reset(new Circuit(msg_mod, carry_mod, n * nb));
        Circuit &c = *cp;
        for (size_t i = 0; i < n; i++)
            for (const Lin &blk : scalar_left_shift(c, radix_at(c, i * nb, nb), (unsigned)shape[2])) c.output(blk);
    } else if (op == "radix_scalar_gt" || op == "radix_scalar_lt" || op == "radix_scalar_eq" || op == "radix_scalar_le" ||
               op == "radix_scalar_ge") {
        need(3);
        const size_t n = shape[0], nb = shape[1];
        cp.reset(new Circuit(msg_mod, carry_mod, n * nb));
        Circuit &c = *cp;
        for (size_t i = 0; i < n; i++) {
            Radix a = radix_at(c, i * nb, nb);
            c.output(op == "radix_scalar_gt" ? scalar_gt(c, a, shape[2]) : op == "radix_scalar_lt" ? scalar_lt(c, a, shape[2])
                     : op == "radix_scalar_le" ? scalar_le(c, a, shape[2]) : op == "radix_scalar_ge" ? scalar_ge(c, a, shape[2])
                     : scalar_eq(c, a, shape[2]));
        }
    } else if (op == "string_eq" || op == "string_ne" || op == "string_starts_with" || op == "string_ends_with") {
        need(4);
        const size_t n = shape[0], la = shape[1], lb = shape[2], nb = shape[3];
        cp.reset(new Circuit(msg_mod, carry_mod, n * (la + lb) * nb));
        Circuit &c = *cp;
        for (size_t s = 0; s < n; s++) {
            FheChars a(la), b(lb);
            for (size_t i = 0; i < la; i++) a[i] = radix_at(c, (s * la + i) * nb, nb);
            for (size_t i = 0; i < lb; i++) b[i] = radix_at(c, (n * la + s * lb + i) * nb, nb);
            c.output(op == "string_eq" ? string_eq(c, a, b) : op == "string_ne" ? bool_not(c, string_eq(c, a, b))
                     : op == "string_starts_with" ? string_starts_with(c, a, b) : string_ends_with(c, a, b));
        }
    } else if (op == "string_to_uppercase" || op == "string_to_lowercase" || op == "string_to_uppercase_reference" ||
               op == "string_to_lowercase_reference") {
        need(3);
        const size_t n = shape[0], len = shape[1], nb = shape[2];
        cp.reset(new Circuit(msg_mod, carry_mod, n * len * nb));
        Circuit &c = *cp;
        for (size_t s = 0; s < n; s++)
            for (size_t i = 0; i < len; i++) {
                Radix ch = radix_at(c, (s * len + i) * nb, nb);
                Radix r = op == "string_to_uppercase" ? to_uppercase_char(c, ch) : op == "string_to_lowercase" ? to_lowercase_char(c, ch)
                        : op == "string_to_uppercase_reference" ? to_uppercase_char_reference(c, ch) : to_lowercase_char_reference(c, ch);
                for (const Lin &blk : r) c.output(blk);
            }
    } else if (op == "string_contains" || op == "string_find") {
        need(4);
        const size_t n = shape[0], hl = shape[1], pl = shape[2], nb = shape[3];
        if (op == "string_find") {
            // the index (number of positions before the first match, hay_len - pat_len + 1 when absent) is a radix
            // integer of nb blocks: it must not wrap
            double cap = 1;
            for (size_t k = 0; k < nb; k++) cap *= msg_mod;
            if (pl <= hl && (double)(hl - pl + 1) >= cap)
                throw std::invalid_argument("string_find: hay_len - pat_len + 1 does not fit the index radix (message_modulus^nb)");
        }
        cp.reset(new Circuit(msg_mod, carry_mod, n * (hl + pl) * nb));
        Circuit &c = *cp;
        for (size_t s = 0; s < n; s++) {
            FheChars h(hl), p(pl);
            for (size_t i = 0; i < hl; i++) h[i] = radix_at(c, (s * hl + i) * nb, nb);
            for (size_t i = 0; i < pl; i++) p[i] = radix_at(c, (n * hl + s * pl + i) * nb, nb);
            if (op == "string_contains") c.output(string_contains(c, h, p));
            else {
                auto fr = string_find(c, h, p, nb);
                c.output(fr.first);
                for (const Lin &blk : fr.second) c.output(blk);
            }
        }
    } else if (op == "trivium") {
        need(3);
        const size_t rounds = shape[0];
        cp.reset(new Circuit(msg_mod, carry_mod, 80));
        Circuit &c = *cp;
        std::vector<Lin> key(80);
        std::vector<bool> iv(80);
        for (int i = 0; i < 80; i++) {
            key[i] = c.input(i, 1);
            iv[i] = i < 64 ? ((shape[1] >> i) & 1) : ((shape[2] >> (i - 64)) & 1);
        }
        Trivium t = trivium_init(c, key, iv);
        for (int r = 0; r < 18; r++) trivium_next64(c, t);   // TriviumStream::init, trivium_bool.rs:117-121
        for (size_t r = 0; r < rounds; r++)
            for (const Lin &bit : trivium_next64(c, t)) c.output(bit);
    } else {
        throw std::invalid_argument("unknown program '" + op + "'");
    }
    cp->finalize();
    return cp;
}

// How a named program's inputs / outputs decompose into independent units (shape[0] of them): used to shard a
// program over several GPUs with no exchange between them (every integer / string of a batch is
// independent).  Inputs are `in_seg.size()` consecutive arrays, array s holding units * in_seg[s] blocks.
struct ProgramLayout {
    bool splittable = false;
    size_t units = 1;
    std::vector<size_t> in_seg;
    size_t out_per_unit = 0;
};
inline ProgramLayout program_layout(const std::string &op, const std::vector<uint64_t> &shape) {
    ProgramLayout l;
    auto at = [&](size_t i) { return i < shape.size() ? (size_t)shape[i] : (size_t)0; };
    if (op == "radix_eq" || op == "radix_ne" || op == "radix_gt" || op == "radix_lt" || op == "radix_ge" || op == "radix_le") { l = {true, at(0), {at(1), at(1)}, 1}; }
    else if (op == "radix_add" || op == "radix_sub" || op == "radix_bitand" || op == "radix_bitor" || op == "radix_bitxor" || op == "radix_max" || op == "radix_min") { l = {true, at(0), {at(1), at(1)}, at(1)}; }
    else if (op == "radix_shl") { l = {true, at(0), {at(1)}, at(1)}; }
    else if (op.rfind("radix_scalar_", 0) == 0) { l = {true, at(0), {at(1)}, 1}; }
    else if (op == "string_eq" || op == "string_ne" || op == "string_starts_with" || op == "string_ends_with") { l = {true, at(0), {at(1) * at(3), at(2) * at(3)}, 1}; }
    else if (op.rfind("string_to_uppercase", 0) == 0 || op.rfind("string_to_lowercase", 0) == 0) { l = {true, at(0), {at(1) * at(2)}, at(1) * at(2)}; }
    else if (op == "string_contains") { l = {true, at(0), {at(1) * at(3), at(2) * at(3)}, 1}; }
    else if (op == "string_find") { l = {true, at(0), {at(1) * at(3), at(2) * at(3)}, 1 + at(3)}; }
    return l;   // anything else (trivium, custom circuits): one unit, runs on the first GPU
}

}  // namespace b200
