/*
 * circuit_exec.cpp -- CPU executors for the leveled circuits (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Compiles the product's host-side circuit builder (tfhe_rs_string_b200/csrc/{circuit,workloads,
 * programs}.hpp, header-only, no CUDA) together with two executors so the host logic can be
 * tested without a GPU:
 *   - cleartext: runs a program on message values (mod 2*modulus_sup, padding bit included) with the
 *     exact lookup-table / negacyclic semantics of a PBS;
 *   - encrypted: runs it on real ciphertexts with the CPU oracle's keyswitch + bootstrap.
 */
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../tfhe_rs_string_b200/csrc/programs.hpp"
#include "tfhe_oracle.h"

using namespace b200;

static thread_local std::string g_err;

static std::unique_ptr<Circuit> build(const char *op, const uint64_t *shape, size_t n_shape, uint32_t mm, uint32_t cm) {
    return build_program(op, std::vector<uint64_t>(shape, shape + n_shape), mm, cm);
}

extern "C" {

const char *orc_circuit_last_error() { return g_err.c_str(); }

/* info[0..5] = n_inputs, n_outputs, n_pbs, depth, n_stages, n_luts */
int orc_circuit_info(const char *op, const uint64_t *shape, size_t n_shape, uint32_t mm, uint32_t cm, uint64_t *info) {
    try {
        auto c = build(op, shape, n_shape, mm, cm);
        info[0] = c->n_inputs(); info[1] = c->outputs.size(); info[2] = c->n_pbs(); info[3] = c->depth();
        info[4] = c->stages.size(); info[5] = c->luts.size();
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

/* Messages in, messages out (values modulo 2*modulus_sup). */
int orc_circuit_run_cleartext(const char *op, const uint64_t *shape, size_t n_shape, uint32_t mm, uint32_t cm,
                              const uint64_t *in_msgs, uint64_t *out_msgs) {
    try {
        auto c = build(op, shape, n_shape, mm, cm);
        /* values are tracked in HALF message units modulo 4 * modulus_sup (the 16-input reductions of workloads.hpp use
         * tables of +-1/2); a value entering a lookup, and every output, must be a whole number of units */
        const uint64_t ms = (uint64_t)mm * cm, full = 4 * ms;
        std::vector<uint64_t> val(c->n_blocks());
        for (size_t i = 0; i < c->n_inputs(); i++) val[i] = (2 * in_msgs[i]) % full;
        for (size_t k = 0; k < c->nodes.size(); k++) {
            const Node &nd = c->nodes[k];
            int64_t acc = 2 * (int64_t)nd.plaintext + (int64_t)nd.plaintext_half;
            for (uint32_t t = nd.term_begin; t < nd.term_end; t++) {
                if (c->terms[t].block >= (int32_t)(c->n_inputs() + k)) throw std::logic_error("circuit not topologically ordered");
                acc += c->terms[t].coeff * (int64_t)val[c->terms[t].block];
            }
            uint64_t v = (uint64_t)(((acc % (int64_t)full) + (int64_t)full) % (int64_t)full);
            if (nd.lut >= 0) {
                if (v & 1) throw std::logic_error("lookup on a value that is off the message grid");
                const uint64_t m = v / 2;                                  /* message incl. padding bit, < 2 * ms */
                const uint64_t y = c->lut_entry_half_units(nd.lut, m % ms);
                v = m >= ms ? (full - y) % full : y;   /* negacyclic: padding bit set => -LUT */
            }
            val[c->n_inputs() + k] = v;
        }
        for (size_t i = 0; i < c->outputs.size(); i++) {
            if (val[c->outputs[i]] & 1) throw std::logic_error("program output is off the message grid");
            out_msgs[i] = val[c->outputs[i]] / 2;
        }
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

/* Ciphertexts in, ciphertexts out, stage by stage with the oracle's KS+PBS (threads over nodes). */
int orc_circuit_run_encrypted(const orc_keyset *ks, const char *op, const uint64_t *shape, size_t n_shape,
                              const uint64_t *in_cts, uint64_t *out_cts, int n_threads) {
    try {
        const orc_params *p = orc_keyset_params(ks);
        auto c = build(op, shape, n_shape, p->message_modulus, p->carry_modulus);
        const size_t big = (size_t)p->glwe_dimension * p->polynomial_size + 1;
        const uint64_t ms = (uint64_t)p->message_modulus * p->carry_modulus;
        const uint64_t delta = ((uint64_t)1 << 63) / ms;
        const size_t glwe_len = (size_t)(p->glwe_dimension + 1) * p->polynomial_size;
        std::vector<uint64_t> luts(c->luts.size() * glwe_len, 0);
        for (size_t l = 0; l < c->luts.size(); l++) {   /* mask polynomials zero, body as fill_accumulator builds it */
            const std::vector<uint64_t> body = c->lut_body((int)l, p->polynomial_size);
            std::memcpy(&luts[l * glwe_len + (size_t)p->glwe_dimension * p->polynomial_size], body.data(), body.size() * sizeof(uint64_t));
        }
        std::vector<uint64_t> pool(c->n_blocks() * big);
        std::memcpy(pool.data(), in_cts, c->n_inputs() * big * sizeof(uint64_t));
        if (n_threads < 1) n_threads = 1;
        for (const Circuit::Stage &st : c->stages) {
            auto work = [&](size_t t) {
                std::vector<uint64_t> tmp(big);
                for (size_t k = st.begin + t; k < st.end; k += (size_t)n_threads) {
                    const Node &nd = c->nodes[k];
                    std::fill(tmp.begin(), tmp.end(), 0);
                    for (uint32_t q = nd.term_begin; q < nd.term_end; q++) {
                        const uint64_t *src = &pool[(size_t)c->terms[q].block * big];
                        const uint64_t cf = (uint64_t)c->terms[q].coeff;
                        for (size_t j = 0; j < big; j++) tmp[j] += cf * src[j];
                    }
                    tmp[big - 1] += nd.plaintext * delta + nd.plaintext_half * (delta / 2);
                    uint64_t *dst = &pool[(c->n_inputs() + k) * big];
                    if (nd.lut >= 0) orc_ks_pbs(ks, tmp.data(), &luts[(size_t)nd.lut * glwe_len], dst);
                    else std::memcpy(dst, tmp.data(), big * sizeof(uint64_t));
                }
            };
            std::vector<std::thread> th;
            for (int t = 0; t < n_threads; t++) th.emplace_back(work, (size_t)t);
            for (auto &x : th) x.join();
        }
        for (size_t i = 0; i < c->outputs.size(); i++)
            std::memcpy(out_cts + i * big, &pool[(size_t)c->outputs[i] * big], big * sizeof(uint64_t));
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

}  // extern "C"
