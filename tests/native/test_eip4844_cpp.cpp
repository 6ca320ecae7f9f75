// The reference's own unit tests (lib/src/primitives/eip4844.rs:147-214), re-stated in C++ against
// include/raiko_kzg.hpp.  Built and run by tests/test_gpu_cpp_host.py on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include "../../include/raiko_kzg.hpp"
using namespace raiko::eip4844;

static std::string hex(const uint8_t* p, size_t n) { static const char* d = "0123456789abcdef"; std::string s; for (size_t i = 0; i < n; i++) { s += d[p[i] >> 4]; s += d[p[i] & 15]; } return s; }
#define CHECK(c) do { if (!(c)) { printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<uint8_t> image((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    KZGSettings settings(image, {}, argc > 2 ? atoi(argv[2]) : 8);
    // test_blob_to_kzg_commitment (eip4844.rs:147-160)
    std::vector<uint8_t> zero(131072, 0);
    auto c0 = calc_kzg_proof_commitment(settings, zero);
    CHECK("0x" + hex(commitment_to_version_hash(c0).data(), 32) == "0x010657f37554c781402a22917dee2f75def7ab966d7b770905398eba3c444014");
    // test_verify_kzg_proof (eip4844.rs:162-184)
    std::vector<uint8_t> data(131072);
    for (size_t v = 0; v < data.size(); v++) data[v] = (uint8_t)(v % 64);
    auto commitment = calc_kzg_proof_commitment(settings, data);
    KzgField x; x.fill(5);                                   // hash_to_bls_field(&[5; 32])
    KzgGroup proof; KzgField y;
    CHECK(rk_compute_kzg_proof(settings.get(), data.data(), data.size(), x.data(), proof.data(), y.data()) == RK_OK);
    CHECK(hex(proof.data(), 48) == hex(calc_kzg_proof_with_point(settings, data, x).data(), 48));
    CHECK(verify_kzg_proof(settings, commitment, x, y, proof));
    // test_verify_kzg_proof_in_precompile negatives (eip4844.rs:202-213)
    KzgField x6; x6.fill(6);
    auto proof6 = calc_kzg_proof_with_point(settings, data, x6);
    CHECK(!verify_kzg_proof(settings, commitment, x6, y, proof6));
    KzgField y1 = y; for (int i = 31; i >= 0; i--) if (++y1[i] != 0) break;
    CHECK(!verify_kzg_proof(settings, commitment, x, y1, proof));
    // wrapper semantics
    auto vh = commitment_to_version_hash(commitment);
    auto xy = proof_of_equivalence(settings, data, vh);
    CHECK(xy.first == get_evaluation_point(settings, data, vh));
    CHECK(calc_kzg_proof(settings, data, vh) == calc_kzg_proof_with_point(settings, data, xy.first));
    std::vector<uint8_t> bad = data; for (int i = 0; i < 32; i++) bad[i] = 0xff;
    try { calc_kzg_proof_commitment(settings, bad); CHECK(false); } catch (const Eip4844Error& e) { CHECK(e.kind == Eip4844Error::DeserializeBlob); }
    try { calc_kzg_proof_commitment(settings, std::vector<uint8_t>(100)); CHECK(false); } catch (const Eip4844Error& e) { CHECK(e.kind == Eip4844Error::DeserializeBlob); }
    printf("CPP_HOST_OK %s\n", hex(commitment.data(), 48).c_str());
    return 0;
}
