// key_import.hpp -- reader for the reference's server-key wire format (bincode 1.x default options: little endian,
// fixed-width integers, usize / sequence lengths as u64, enum variant index as u32, u128 as 16 bytes).
//
// PARITY STATUS: UNPINNED.  No Rust toolchain exists in the build image, so no byte stream produced by the reference
// has ever been fed to this parser; it follows the serde field order of the reference's structs, and the test fixture
// (tests/golden/make_server_key_fixture.py) is built from the same reading of those structs.
//
// Two layouts are recognised:
//  (A) bincode(shortint::ServerKey), shortint/server_key/mod.rs:283-297:
//        key_switching_key  LweKeyswitchKey<Vec<u64>> { data: Vec<u64>, decomp_base_log, decomp_level_count,
//                                                       output_lwe_size, ciphertext_modulus }   (entities/lwe_keyswitch_key.rs:76-86)
//        bootstrapping_key  enum tag u32 (0 = Classic; MultiBit is rejected), then FourierLweBootstrapKey { fourier,
//                           input_lwe_dimension, glwe_size, decomposition_base_log, decomposition_level_count }
//                           (fft64/crypto/bootstrap.rs:24-32); `fourier` is FourierPolynomialList's hand-written
//                           sequence: len = 2 + count, polynomial_size, count, then per polynomial a sequence of N/2
//                           c64 (fft/mod.rs:588-630) in concrete-fft's serialisation order
//        message_modulus, carry_modulus, max_degree, max_noise_level (usize newtypes), ciphertext_modulus, pbs_order
//      CiphertextModulus serialises as { modulus: u128 (0 = native), scalar_bits: usize } (commons/ciphertext_modulus.rs:41-63).
//  (B) the standard-domain bundle the Rust shim writes where the reference still holds the standard key
//      (shortint/engine/server_side.rs:63-86): bincode of the tuple
//        (LweKeyswitchKey<Vec<u64>>, LweBootstrapKey<Vec<u64>>, MessageModulus, CarryModulus, PBSOrder)
//      with LweBootstrapKey = GgswCiphertextList { data, glwe_size, polynomial_size, decomp_base_log,
//      decomp_level_count, ciphertext_modulus } (entities/ggsw_ciphertext_list.rs:9-20).
// Compressed (seeded) keys are out of scope: expanding them needs the reference's AES-CTR CSPRNG.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/b200tfhe.h"

namespace b200 {

class BincodeReader {
  public:
    BincodeReader(const uint8_t *p, size_t n) : p_(p), n_(n) {}
    bool ok() const { return ok_; }
    size_t pos() const { return pos_; }
    uint64_t u64() {
        uint64_t v = 0;
        if (!need(8)) return 0;
        std::memcpy(&v, p_ + pos_, 8);   // the build targets little-endian hosts only (x86-64, aarch64)
        pos_ += 8;
        return v;
    }
    uint32_t u32() {
        uint32_t v = 0;
        if (!need(4)) return 0;
        std::memcpy(&v, p_ + pos_, 4);
        pos_ += 4;
        return v;
    }
    void skip(size_t bytes) {
        if (need(bytes)) pos_ += bytes;
    }
    // CiphertextModulus<u64>: only the native modulus 2^64 is supported by the kernels
    bool native_modulus(std::string *err) {
        const uint64_t lo = u64(), hi = u64(), bits = u64();
        if (!ok_) return false;
        if (lo != 0 || hi != 0) { *err = "non-native ciphertext modulus"; return false; }
        if (bits != 64) { *err = "ciphertext modulus carries " + std::to_string(bits) + " scalar bits, expected 64"; return false; }
        return true;
    }

  private:
    bool need(size_t bytes) {
        if (!ok_ || bytes > n_ - pos_) { ok_ = false; return false; }
        return true;
    }
    const uint8_t *p_;
    size_t n_, pos_ = 0;
    bool ok_ = true;
};

inline bool parse_shortint_server_key(const uint8_t *bytes, size_t n_bytes, b200tfhe_params *params, b200tfhe_key_view *view,
                                      std::string *err) {
    BincodeReader r(bytes, n_bytes);
    std::memset(params, 0, sizeof(*params));
    std::memset(view, 0, sizeof(*view));
    // ---- LweKeyswitchKey<Vec<u64>>
    view->ksk_len = r.u64();
    view->ksk_offset = r.pos();
    if (!r.ok() || view->ksk_len > (n_bytes - r.pos()) / 8) { *err = "truncated keyswitch key"; return false; }
    r.skip(view->ksk_len * 8);
    params->ks_base_log = (uint32_t)r.u64();
    params->ks_level = (uint32_t)r.u64();
    const uint64_t out_lwe_size = r.u64();
    if (!r.native_modulus(err)) { if (err->empty()) *err = "truncated keyswitch key header"; return false; }
    if (out_lwe_size < 2 || params->ks_level == 0 || view->ksk_len % (out_lwe_size * params->ks_level) != 0) {
        *err = "inconsistent keyswitch key dimensions";
        return false;
    }
    params->lwe_dimension = (uint32_t)(out_lwe_size - 1);
    const uint64_t ks_in_dim = view->ksk_len / (out_lwe_size * params->ks_level);
    // ---- bootstrap key: layout (A) starts with the enum tag and the sequence header, layout (B) with a Vec length
    const size_t mark = r.pos();
    {
        BincodeReader probe(bytes + mark, n_bytes - mark);
        const uint32_t tag = probe.u32();
        const uint64_t seq_len = probe.u64(), poly_size = probe.u64(), count = probe.u64();
        const bool pow2 = poly_size >= 256 && (poly_size & (poly_size - 1)) == 0;
        view->bsk_is_fourier = probe.ok() && tag <= 1 && seq_len == count + 2 && pow2 ? 1 : 0;
        if (view->bsk_is_fourier && tag == 1) { *err = "multi-bit bootstrap keys are not supported"; return false; }
    }
    uint64_t glwe_size = 0, poly_size = 0;
    if (view->bsk_is_fourier) {
        r.u32();
        r.u64();
        poly_size = r.u64();
        const uint64_t count = r.u64();
        view->bsk_offset = r.pos() + 8;                       // first polynomial's data (after its own length prefix)
        view->bsk_poly_stride_bytes = 8 + (poly_size / 2) * 16;
        view->bsk_len = count * (poly_size / 2);              // complex elements
        for (uint64_t c = 0; c < count && r.ok(); c++) {
            if (r.u64() != poly_size / 2) { *err = "Fourier polynomial of unexpected length"; return false; }
            r.skip((poly_size / 2) * 16);
        }
        const uint64_t in_dim = r.u64();
        glwe_size = r.u64();
        params->pbs_base_log = (uint32_t)r.u64();
        params->pbs_level = (uint32_t)r.u64();
        if (!r.ok() || glwe_size < 2 || count != in_dim * params->pbs_level * glwe_size * glwe_size || in_dim != params->lwe_dimension) {
            *err = "inconsistent Fourier bootstrap key dimensions";
            return false;
        }
        params->message_modulus = (uint32_t)r.u64();
        params->carry_modulus = (uint32_t)r.u64();
        view->max_degree = r.u64();
        view->max_noise_level = r.u64();
        if (!r.native_modulus(err)) { if (err->empty()) *err = "truncated server key trailer"; return false; }
        view->pbs_order = r.u32();
    } else {
        view->bsk_len = r.u64();
        view->bsk_offset = r.pos();
        if (!r.ok() || view->bsk_len > (n_bytes - r.pos()) / 8) { *err = "truncated bootstrap key"; return false; }
        r.skip(view->bsk_len * 8);
        glwe_size = r.u64();
        poly_size = r.u64();
        params->pbs_base_log = (uint32_t)r.u64();
        params->pbs_level = (uint32_t)r.u64();
        if (!r.native_modulus(err)) { if (err->empty()) *err = "truncated bootstrap key header"; return false; }
        const uint64_t per_ggsw = (uint64_t)params->pbs_level * glwe_size * glwe_size * poly_size;
        if (glwe_size < 2 || poly_size == 0 || per_ggsw == 0 || view->bsk_len != per_ggsw * params->lwe_dimension) {
            *err = "inconsistent bootstrap key dimensions";
            return false;
        }
        view->bsk_poly_stride_bytes = poly_size * 8;
        params->message_modulus = (uint32_t)r.u64();
        params->carry_modulus = (uint32_t)r.u64();
        view->pbs_order = r.u32();
    }
    if (!r.ok()) { *err = "truncated server key"; return false; }
    if (r.pos() != n_bytes) { *err = "trailing bytes after the server key"; return false; }
    if (view->pbs_order > 1) { *err = "unknown PBS order"; return false; }
    params->glwe_dimension = (uint32_t)(glwe_size - 1);
    params->polynomial_size = (uint32_t)poly_size;
    if (ks_in_dim != (uint64_t)params->glwe_dimension * params->polynomial_size) {
        *err = "keyswitch key input dimension does not match the GLWE dimensions";
        return false;
    }
    return true;
}

}  // namespace b200
