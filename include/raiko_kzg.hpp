// raiko_kzg.hpp -- C++ host-side mirror of /root/reference/lib/src/primitives/eip4844.rs over the
// C ABI of raiko_kzg.h (header only).  Same function names, argument meaning and error
// behaviour as the Rust wrapper; Result<_, Eip4844Error> becomes an exception of the same name.
// The reference is compiled (Rust) code and no Rust toolchain exists in this image, so this is
// the compiled-language host side; bindings/rust/ carries the equivalent crate as source.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "raiko_kzg.h"

namespace raiko {
namespace eip4844 {

constexpr uint8_t VERSIONED_HASH_VERSION_KZG = 0x01;            // eip4844.rs:26
using KzgGroup = std::array<uint8_t, 48>;                        // eip4844.rs:28
using KzgField = std::array<uint8_t, 32>;                        // eip4844.rs:29
using KzgCommitment = KzgGroup;                                  // eip4844.rs:30
using B256 = std::array<uint8_t, 32>;

// eip4844.rs:32-42
struct Eip4844Error : std::runtime_error {
    enum Kind { DeserializeBlob, EvaluatePolynomial, ComputeKzgProof, KzgDataPoison } kind;
    Eip4844Error(Kind k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

namespace detail {
inline void check(rk_status st, Eip4844Error::Kind other, const char* other_msg) {
    if (st == RK_OK) return;
    if (st == RK_ERR_BAD_LENGTH || st == RK_ERR_NONCANONICAL_FE)
        throw Eip4844Error(Eip4844Error::DeserializeBlob, "Failed to deserialize blob to field elements");
    throw Eip4844Error(other, std::string(other_msg) + ": " + rk_last_error());
}
}  // namespace detail

// KZGSettings (eip4844.rs:13) / KZG_SETTINGS (eip4844.rs:21-24): owns the GPU-resident tables.
class KZGSettings {
  public:
    explicit KZGSettings(const std::vector<uint8_t>& image, const std::vector<int>& devices = {}, int window_bits = 0) {
        rk_kzg_ctx* c = nullptr;
        rk_status st = rk_kzg_ctx_create_ex(image.data(), image.size(), devices.empty() ? nullptr : devices.data(),
                                            (int)devices.size(), window_bits, &c);
        if (st != RK_OK)
            throw std::runtime_error(std::string("failed to load trusted setup, please run `cargo run --bin gen_kzg_settings`: ") + rk_last_error());
        ctx_.reset(c, rk_kzg_ctx_destroy);
    }
    rk_kzg_ctx* get() const { return ctx_.get(); }

  private:
    std::shared_ptr<rk_kzg_ctx> ctx_;
};
using KzgSettings = KZGSettings;                                 // north-star spelling

// eip4844.rs:91-95
inline B256 commitment_to_version_hash(const KzgCommitment& commitment) {
    B256 h;
    rk_kzg_to_versioned_hash(commitment.data(), h.data());
    return h;
}
// eip4844.rs:80-89
inline KzgGroup calc_kzg_proof_commitment(const KZGSettings& s, const std::vector<uint8_t>& blob) {
    KzgGroup out;
    detail::check(rk_blob_to_kzg_commitment(s.get(), blob.data(), blob.size(), out.data()), Eip4844Error::ComputeKzgProof,
                  "Failed to compute KZG proof");
    return out;
}
// eip4844.rs:44-48 (the reference returns ZFr; here its 32-byte big-endian form)
inline KzgField get_evaluation_point(const KZGSettings& s, const std::vector<uint8_t>& blob, const B256& versioned_hash) {
    KzgField x;
    detail::check(rk_get_evaluation_point(s.get(), blob.data(), blob.size(), versioned_hash.data(), x.data()),
                  Eip4844Error::EvaluatePolynomial, "Failed to evaluate polynomial at hashed point");
    return x;
}
// eip4844.rs:50-65
inline std::pair<KzgField, KzgField> proof_of_equivalence(const KZGSettings& s, const std::vector<uint8_t>& blob,
                                                           const B256& versioned_hash) {
    KzgField x, y;
    detail::check(rk_proof_of_equivalence(s.get(), blob.data(), blob.size(), versioned_hash.data(), x.data(), y.data()),
                  Eip4844Error::EvaluatePolynomial, "Failed to evaluate polynomial at hashed point");
    return {x, y};
}
// eip4844.rs:71-78 (ZG1 travels as its 48-byte encoding)
inline KzgGroup calc_kzg_proof_with_point(const KZGSettings& s, const std::vector<uint8_t>& blob, const KzgField& z) {
    KzgGroup proof;
    detail::check(rk_compute_kzg_proof(s.get(), blob.data(), blob.size(), z.data(), proof.data(), nullptr),
                  Eip4844Error::ComputeKzgProof, "Failed to compute KZG proof");
    return proof;
}
// eip4844.rs:67-69
inline KzgGroup calc_kzg_proof(const KZGSettings& s, const std::vector<uint8_t>& blob, const B256& versioned_hash) {
    KzgGroup proof;
    detail::check(rk_calc_kzg_proof(s.get(), blob.data(), blob.size(), versioned_hash.data(), proof.data()),
                  Eip4844Error::ComputeKzgProof, "Failed to compute KZG proof");
    return proof;
}
// eip4844.rs:97-99
inline KzgGroup kzg_proof_to_bytes(const KzgGroup& proof) { return proof; }
// verify_kzg_proof_rust as called in the reference's tests (eip4844.rs:176-183)
inline bool verify_kzg_proof(const KZGSettings& s, const KzgCommitment& c, const KzgField& z, const KzgField& y, const KzgGroup& proof) {
    int ok = 0;
    rk_status st = rk_verify_kzg_proof(s.get(), c.data(), z.data(), y.data(), proof.data(), &ok);
    if (st != RK_OK) throw std::runtime_error(rk_last_error());
    return ok != 0;
}
// c-kzg era spellings used by BASELINE.json's north_star
inline KzgGroup blob_to_kzg_commitment(const KZGSettings& s, const std::vector<uint8_t>& blob) { return calc_kzg_proof_commitment(s, blob); }
inline B256 kzg_to_versioned_hash(const KzgCommitment& c) { return commitment_to_version_hash(c); }

}  // namespace eip4844
}  // namespace raiko
