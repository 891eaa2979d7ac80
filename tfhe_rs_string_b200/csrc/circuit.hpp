// circuit.hpp -- level-synchronous description of a batch of shortint operations (host side, no CUDA).
//
// The reference issues one KS+PBS per call of shortint::ServerKey::apply_lookup_table and gets
// its parallelism from rayon (integer/server_key/radix_parallel/*: par_iter over blocks,
// rayon::join over independent sub-expressions).  On a GPU the same work is expressed as a DAG
// whose nodes are "linear combination of earlier blocks (+ plaintext), then one KS+PBS with a
// lookup table"; all nodes of one dependency level become ONE lwe-linear launch and ONE
// ks_pbs_batch launch.  This file only builds and levels the DAG; executors
// (CUDA: executor.cuh; tests: cleartext / CPU-oracle executors under oracle/) run it.
//
// Semantics mirrored from the reference (paths relative to tfhe/src/):
//   * leveled ops between bootstraps are plain LWE linear algebra (core_crypto/algorithms/
//     lwe_linear_algebra.rs:68,276,556,703; shortint/server_key/add.rs:520-524, scalar_add.rs:211-218,
//     scalar_mul.rs) -> Lin values below, never materialised unless needed;
//   * a bootstrap of a trivial (constant) value is evaluated on the host exactly like
//     ServerKey::trivial_pbs_assign (shortint/server_key/mod.rs:763-781), including the negated
//     output when the padding bit is set;
//   * bivariate functions pack lhs * factor + rhs before one PBS (shortint/server_key/
//     bivariate_pbs.rs:71-98,167-181).
#pragma once
#include <algorithm>
#include <cstdint>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace b200 {

struct Term {
    int32_t block;   // id of an earlier block (inputs first, then node outputs)
    int64_t coeff;   // small signed scalar
};

// A not-yet-materialised linear combination of blocks plus a constant (in message units, i.e.
// multiples of delta, kept modulo 2 * message_modulus * carry_modulus: the padding bit counts).
struct Lin {
    std::vector<Term> terms;
    uint64_t cst = 0;
    int64_t half = 0;      // additional constant in HALF message units (delta / 2): only the 16-input reductions of
                           // workloads.hpp produce it
    int64_t odd = 0;       // sum of the coefficients of blocks that hold an odd number of half units (outputs of half-unit
                           // tables); the value is a whole number of units iff half + odd is even
    uint32_t level = 0;    // dependency level of the deepest block referenced
    uint32_t degree = 0;   // upper bound of the encrypted value (reference: Degree bookkeeping)
    bool is_const() const { return terms.empty(); }
    bool is_plain_block() const { return terms.size() == 1 && terms[0].coeff == 1 && cst == 0; }
};

struct Node {
    uint32_t term_begin = 0, term_end = 0;  // into Circuit::terms
    uint64_t plaintext = 0;                 // message units, added to the body
    uint64_t plaintext_half = 0;            // plus this many half units (delta / 2), modulo 4 * modulus_sup
    int32_t lut = -1;                       // index into Circuit::luts, -1 = linear only (no PBS)
    uint32_t level = 0;
};

class Circuit {
  public:
    Circuit(uint32_t message_modulus, uint32_t carry_modulus, size_t n_inputs)
        : msg_mod(message_modulus), carry_mod(carry_modulus), n_inputs_(n_inputs) {}

    const uint32_t msg_mod, carry_mod;
    uint32_t modulus_sup() const { return msg_mod * carry_mod; }

    // ---- values
    Lin input(size_t i, uint32_t degree) const {
        if (i >= n_inputs_) throw std::out_of_range("circuit input index");
        Lin l;
        l.terms.push_back({(int32_t)i, 1});
        l.degree = degree;
        return l;
    }
    Lin input(size_t i) const { return input(i, msg_mod - 1); }
    Lin constant(uint64_t m) const {
        Lin l;
        l.cst = m % (2 * (uint64_t)modulus_sup());
        l.degree = (uint32_t)l.cst;
        return l;
    }

    // ---- leveled (linear) operations: no node is created
    Lin add(const Lin &a, const Lin &b) const { return axpy(a, 1, b, 1); }
    Lin sub(const Lin &a, const Lin &b) const { return axpy(a, 1, b, -1); }   // true LWE subtraction
    Lin scale(const Lin &a, int64_t c) const { return axpy(a, c, Lin{}, 0); }
    Lin add_const(const Lin &a, int64_t m) const {
        Lin r = a;
        const int64_t mod = 2 * (int64_t)modulus_sup();
        r.cst = (uint64_t)((((int64_t)a.cst + m) % mod + mod) % mod);
        r.degree = a.degree + (uint32_t)(m > 0 ? m : 0);
        return r;
    }
    // a + h * delta / 2 (leveled): undoes the -1/2 offset of a half-unit lookup table (lut_half)
    Lin add_half(const Lin &a, int64_t h, uint32_t degree) const {
        Lin r = a;
        r.half += h;
        r.degree = degree;
        return r;
    }
    Lin axpy(const Lin &a, int64_t ca, const Lin &b, int64_t cb) const {
        std::map<int32_t, int64_t> acc;
        for (const Term &t : a.terms) acc[t.block] += t.coeff * ca;
        for (const Term &t : b.terms) acc[t.block] += t.coeff * cb;
        Lin r;
        for (auto &kv : acc)
            if (kv.second != 0) r.terms.push_back({kv.first, kv.second});
        const int64_t mod = 2 * (int64_t)modulus_sup();
        int64_t c = ((int64_t)a.cst * ca + (int64_t)b.cst * cb) % mod;
        r.cst = (uint64_t)((c + mod) % mod);
        r.half = a.half * ca + b.half * cb;
        r.odd = a.odd * ca + b.odd * cb;
        r.level = std::max(a.level, b.level);
        r.degree = a.degree * (uint32_t)(ca < 0 ? -ca : ca) + b.degree * (uint32_t)(cb < 0 ? -cb : cb);
        return r;
    }

    // ---- lookup tables (generate_lookup_table, shortint/server_key/mod.rs:383-399)
    int lut(const std::function<uint64_t(uint64_t)> &f) {
        std::vector<uint64_t> table(modulus_sup());
        for (uint32_t x = 0; x < modulus_sup(); x++) table[x] = f(x);
        auto it = lut_index_.find(table);
        if (it != lut_index_.end()) return it->second;
        const int id = (int)luts.size();
        luts.push_back(table);
        lut_index_[table] = id;
        return id;
    }
    // A table whose entries are in HALF message units (delta / 2), possibly negative.  The negacyclic rule LUT(x + modulus_sup)
    // = -LUT(x) makes a lookup on the 17 values 0..modulus_sup well defined only for outputs symmetric around 0; a table
    // of +-1/2 followed by add_half(+1) is how a PBS tests a sum of modulus_sup boolean flags (one more than the
    // reference's are_all_comparisons_block_true can take, scalar_comparison.rs:161-183), see workloads.hpp all_true.
    int lut_half(const std::function<int64_t(uint64_t)> &f) {
        std::vector<uint64_t> table(modulus_sup());
        const uint64_t mod = 4 * (uint64_t)modulus_sup();
        for (uint32_t x = 0; x < modulus_sup(); x++) table[x] = (uint64_t)(((f(x) % (int64_t)mod) + (int64_t)mod) % (int64_t)mod) | kHalfTag;
        auto it = lut_index_.find(table);
        if (it != lut_index_.end()) return it->second;
        const int id = (int)luts.size();
        luts.push_back(table);
        lut_index_[table] = id;
        return id;
    }
    // entries of a half-unit table carry this tag bit so that equal numbers in different units stay different tables
    static constexpr uint64_t kHalfTag = (uint64_t)1 << 62;
    bool lut_is_half(int id) const { return !luts[id].empty() && (luts[id][0] & kHalfTag) != 0; }
    // entry x of table id in half units modulo 4 * modulus_sup (whole-unit tables: twice the entry)
    uint64_t lut_entry_half_units(int id, uint64_t x) const {
        const uint64_t e = luts[id][x], mod = 4 * (uint64_t)modulus_sup();
        return lut_is_half(id) ? (e & ~kHalfTag) % mod : (2 * e) % mod;
    }
    // GLWE body of table id (fill_accumulator, shortint/engine/mod.rs:92-127) with entries scaled by delta or delta / 2
    std::vector<uint64_t> lut_body(int id, uint32_t polynomial_size) const {
        const size_t ms = modulus_sup(), box = polynomial_size / ms, half_box = box / 2;
        const uint64_t half_delta = ((uint64_t)1 << 62) / ms;
        std::vector<uint64_t> body(polynomial_size);
        for (size_t i = 0; i < ms; i++)
            for (size_t j = 0; j < box; j++) body[i * box + j] = lut_entry_half_units(id, i) * half_delta;
        for (size_t j = 0; j < half_box; j++) body[j] = 0 - body[j];
        std::rotate(body.begin(), body.begin() + half_box, body.end());
        return body;
    }
    // generate_lookup_table_bivariate_with_factor, shortint/server_key/bivariate_pbs.rs:71-98
    int lut_bivariate(const std::function<uint64_t(uint64_t, uint64_t)> &f, uint32_t factor) {
        const uint64_t mm = msg_mod;
        return lut([=](uint64_t x) { return f((x / factor) % mm, (x % factor) % mm); });
    }

    // ---- bootstrap: one node, or host evaluation when the input is a constant
    // The reference guards every lookup with its Degree / MaxDegree bookkeeping (checked and smart operations,
    // shortint/server_key/mod.rs:1040-1100): a value that can reach the padding bit would be bootstrapped through the
    // negacyclic half of the table.  Same rule here: the tracked upper bound must stay below
    // message_modulus * carry_modulus (pbs_unchecked is the explicit opt-out, e.g. for caller-built schedules).
    Lin pbs(const Lin &x, int lut_id, uint32_t out_degree) {
        if (!x.is_const() && x.degree >= modulus_sup())
            throw std::logic_error("pbs: operand degree " + std::to_string(x.degree) + " can reach the padding bit (>= " +
                                   std::to_string(modulus_sup()) + "); propagate carries first");
        return pbs_unchecked(x, lut_id, out_degree);
    }
    Lin pbs_unchecked(const Lin &x, int lut_id) {
        uint64_t mx = 0;
        if (!lut_is_half(lut_id))
            for (uint64_t v : luts.at(lut_id)) mx = std::max(mx, v);
        return pbs_unchecked(x, lut_id, (uint32_t)mx);
    }
    Lin pbs_unchecked(const Lin &x, int lut_id, uint32_t out_degree) {
        if (lut_id < 0 || lut_id >= (int)luts.size()) throw std::out_of_range("lut id");
        if ((x.half + x.odd) & 1) throw std::logic_error("pbs: operand is off the message grid by half a unit");
        if (x.is_const() && x.half != 0) {   // an even number of half units is a whole number of units
            Lin y = add_const(x, x.half / 2);
            y.half = 0;
            y.degree = x.degree;
            return pbs_unchecked(y, lut_id, out_degree);
        }
        if (x.is_const()) {
            if (lut_is_half(lut_id)) throw std::logic_error("pbs: half-unit table on a constant operand");
            // trivial_pbs_assign: value = body / delta; negate the table entry when the padding bit is set
            const uint64_t ms = modulus_sup();
            const uint64_t v = x.cst % (2 * ms);
            const uint64_t y = luts[lut_id][v % ms] % (2 * ms);
            return constant(v >= ms ? (2 * ms - y) % (2 * ms) : y);
        }
        Lin r;
        r.terms.push_back({new_node(x, lut_id), 1});
        r.level = x.level + 1;
        r.degree = out_degree;
        r.odd = lut_is_half(lut_id) ? 1 : 0;   // +-1/2 tables leave an odd number of half units in the block
        return r;
    }
    Lin pbs(const Lin &x, int lut_id) {
        uint64_t mx = 0;
        for (uint64_t v : luts.at(lut_id)) mx = std::max(mx, v);
        return pbs(x, lut_id, (uint32_t)mx);
    }
    // unchecked_apply_lookup_table_bivariate: lhs * factor + rhs, then PBS
    Lin pbs_bivariate(const Lin &lhs, const Lin &rhs, int lut_id, uint32_t factor) {
        return pbs(axpy(lhs, (int64_t)factor, rhs, 1), lut_id);
    }

    // A block that physically holds x (needed for outputs); no PBS.
    Lin materialize(const Lin &x) {
        if (x.is_plain_block() && x.half == 0 && x.odd == 0) return x;
        Lin r;
        r.terms.push_back({new_node(x, -1), 1});
        // the node itself runs right after the level that produced its operands (stage key 2L+1);
        // anything that consumes the materialised block is scheduled from the next level on
        r.level = x.level + 1;
        r.degree = x.degree;
        r.odd = (x.half + x.odd) & 1;
        return r;
    }
    // like materialize, but always a fresh block (a caller-built schedule addresses nodes by position)
    Lin materialize_always(const Lin &x) {
        Lin r;
        r.terms.push_back({new_node(x, -1), 1});
        r.level = x.level + 1;
        r.degree = x.degree;
        return r;
    }
    void output(const Lin &x) {
        Lin m = materialize(x);
        outputs.push_back(m.terms[0].block);
    }

    // ---- leveling: stable sort of nodes by level, block ids renumbered; call once before running
    void finalize() {
        if (finalized_) return;
        const size_t n = nodes.size();
        // linear-only nodes depend on blocks of their own level: give them a half step
        std::vector<uint64_t> key(n);
        for (size_t i = 0; i < n; i++) key[i] = (uint64_t)nodes[i].level * 2 + (nodes[i].lut < 0 ? 1 : 0);
        std::vector<uint32_t> order(n);
        for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return key[a] < key[b]; });
        std::vector<int32_t> remap(n);
        for (size_t pos = 0; pos < n; pos++) remap[order[pos]] = (int32_t)(n_inputs_ + pos);
        std::vector<Node> nn(n);
        std::vector<Term> nt;
        nt.reserve(terms.size());
        for (size_t pos = 0; pos < n; pos++) {
            Node nd = nodes[order[pos]];
            const uint32_t b = (uint32_t)nt.size();
            for (uint32_t t = nd.term_begin; t < nd.term_end; t++) {
                Term tm = terms[t];
                if (tm.block >= (int32_t)n_inputs_) tm.block = remap[tm.block - n_inputs_];
                nt.push_back(tm);
            }
            nd.term_begin = b;
            nd.term_end = (uint32_t)nt.size();
            nn[pos] = nd;
        }
        for (int32_t &o : outputs)
            if (o >= (int32_t)n_inputs_) o = remap[o - n_inputs_];
        nodes.swap(nn);
        terms.swap(nt);
        // stage boundaries: maximal runs of equal (level, kind)
        stages.clear();
        size_t i = 0;
        while (i < n) {
            size_t j = i;
            const uint64_t k = (uint64_t)nodes[i].level * 2 + (nodes[i].lut < 0 ? 1 : 0);
            while (j < n && ((uint64_t)nodes[j].level * 2 + (nodes[j].lut < 0 ? 1 : 0)) == k) j++;
            stages.push_back({(uint32_t)i, (uint32_t)j, nodes[i].lut >= 0});
            i = j;
        }
        finalized_ = true;
    }

    struct Stage {
        uint32_t begin, end;   // node range
        bool bootstrap;        // false: linear-only stage
    };

    size_t n_inputs() const { return n_inputs_; }
    size_t n_blocks() const { return n_inputs_ + nodes.size(); }
    size_t n_pbs() const {
        size_t c = 0;
        for (const Node &n : nodes) c += n.lut >= 0;
        return c;
    }
    size_t depth() const {
        uint32_t d = 0;
        for (const Node &n : nodes) d = std::max(d, n.level);
        return d;
    }
    bool finalized() const { return finalized_; }

    std::vector<Node> nodes;
    std::vector<Term> terms;
    std::vector<std::vector<uint64_t>> luts;   // function tables, one entry per message value
    std::vector<int32_t> outputs;              // block ids, in output order
    std::vector<Stage> stages;

  private:
    int32_t new_node(const Lin &x, int lut_id) {
        if (finalized_) throw std::logic_error("circuit already finalized");
        Node nd;
        nd.term_begin = (uint32_t)terms.size();
        for (const Term &t : x.terms) terms.push_back(t);
        nd.term_end = (uint32_t)terms.size();
        nd.plaintext = x.cst;
        const int64_t hm = 4 * (int64_t)modulus_sup();
        nd.plaintext_half = (uint64_t)(((x.half % hm) + hm) % hm);
        nd.lut = lut_id;
        nd.level = lut_id >= 0 ? x.level + 1 : x.level;
        nodes.push_back(nd);
        return (int32_t)(n_inputs_ + nodes.size() - 1);
    }

    size_t n_inputs_;
    bool finalized_ = false;
    std::map<std::vector<uint64_t>, int> lut_index_;
};

}  // namespace b200
