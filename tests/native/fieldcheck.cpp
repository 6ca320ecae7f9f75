// Host-side build of the product's arithmetic headers (field30.cuh, g1.cuh), so the
// limb algorithms can be checked against Python big integers without a GPU.
// TEST INFRASTRUCTURE: never linked into libraiko_kzg.so.
#include <cstring>
#include "../../raiko_b200/csrc/g1.cuh"
using namespace rk;

extern "C" {
// raw 30-bit-limb Montgomery product / square (N = 13 or 9)
void fc_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y, z; memcpy(x.v, a, 52); memcpy(y.v, b, 52); fe_mul(z, x, y); memcpy(r, z.v, 52); }
void fc_fp_sqr(const uint32_t* a, uint32_t* r) { Fp x, z; memcpy(x.v, a, 52); fe_sqr(z, x); memcpy(r, z.v, 52); }
void fc_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fr x, y, z; memcpy(x.v, a, 36); memcpy(y.v, b, 36); fe_mul(z, x, y); memcpy(r, z.v, 36); }
void fc_fr_sqr(const uint32_t* a, uint32_t* r) { Fr x, z; memcpy(x.v, a, 36); fe_sqr(z, x); memcpy(r, z.v, 36); }
void fc_fp_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y, z; memcpy(x.v, a, 52); memcpy(y.v, b, 52); fe_add(z, x, y); memcpy(r, z.v, 52); }
void fc_fp_sub6(const uint32_t* a, const uint32_t* b, uint32_t* r) { Fp x, y, z; memcpy(x.v, a, 52); memcpy(y.v, b, 52); fe_sub<FpTag, 6>(z, x, y); memcpy(r, z.v, 52); }
int fc_fp_is_zero_mod(const uint32_t* a) { Fp x; memcpy(x.v, a, 52); return fe_is_zero_mod(x); }
void fc_fp_pack(const uint32_t* a, uint32_t* w) { Fp x; memcpy(x.v, a, 52); fe_pack<FpTag>(w, x); }
void fc_fp_unpack(const uint32_t* w, uint32_t* a) { Fp x; fe_unpack<FpTag>(x, w); memcpy(a, x.v, 52); }
void fc_fr_pack(const uint32_t* a, uint32_t* w) { Fr x; memcpy(x.v, a, 36); fe_pack<FrTag>(w, x); }
void fc_fr_unpack(const uint32_t* w, uint32_t* a) { Fr x; fe_unpack<FrTag>(x, w); memcpy(a, x.v, 36); }
// canonical words in, canonical words out
void fc_fp_inv(const uint32_t* w, uint32_t* out) { Fp x, m, i, c; fe_unpack<FpTag>(x, w); fe_to_mont(m, x); fe_inv(i, m); fe_from_mont(c, i); fe_pack<FpTag>(out, c); }
void fc_fr_inv(const uint32_t* w, uint32_t* out) { Fr x, m, i, c; fe_unpack<FrTag>(x, w); fe_to_mont(m, x); fe_inv(i, m); fe_from_mont(c, i); fe_pack<FrTag>(out, c); }

// G1: points travel as 48-byte compressed encodings
static int load(G1Xyzz& p, const uint8_t* c) {
    G1Affine a; int rc = g1_decompress(a, c);
    if (rc == 1) { g1_set_inf(p); return 0; }
    if (rc) return rc;
    g1_from_affine(p, a); return 0;
}
int fc_g1_roundtrip(const uint8_t* in, uint8_t* out) { G1Xyzz p; int rc = load(p, in); if (rc) return rc; g1_compress(out, p); return 0; }
int fc_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out) { G1Xyzz p, q; if (load(p, a) || load(q, b)) return -1; g1_add(p, q); g1_compress(out, p); return 0; }
int fc_g1_madd(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    G1Xyzz p; G1Affine q; if (load(p, a)) return -1; if (g1_decompress(q, b)) return -1;
    g1_madd(p, q.x, q.y); g1_compress(out, p); return 0; }
int fc_g1_dbl(const uint8_t* a, uint8_t* out) { G1Xyzz p, r; if (load(p, a)) return -1; g1_dbl(r, p); g1_compress(out, r); return 0; }
// k*P by double-and-add through madd/dbl (exercises long chains with loose bounds)
int fc_g1_mul_u64(const uint8_t* a, uint64_t k, uint8_t* out) {
    G1Affine q; if (g1_decompress(q, a)) return -1;
    G1Xyzz acc; g1_set_inf(acc);
    for (int bit = 63; bit >= 0; bit--) { G1Xyzz t; g1_dbl(t, acc); acc = t; if ((k >> bit) & 1) g1_madd(acc, q.x, q.y); }
    g1_compress(out, acc); return 0; }
// sum of n points by repeated madd (long accumulation chain)
int fc_g1_sum(const uint8_t* pts, int n, uint8_t* out) {
    G1Xyzz acc; g1_set_inf(acc);
    for (int i = 0; i < n; i++) { G1Affine q; int rc = g1_decompress(q, pts + 48 * i); if (rc == 1) continue; if (rc) return rc; g1_madd(acc, q.x, q.y); }
    g1_compress(out, acc); return 0; }
}
