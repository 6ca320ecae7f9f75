/* kzg_ref.c -- plain-C CPU restatement of raiko's blob-KZG path.
 *
 * TEST INFRASTRUCTURE ONLY (the checker and the timed CPU baseline): nothing under
 * raiko_b200/ may link or call this.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Restates, single-threaded like the reference (workspace dependency declared with
 * default-features = false, /root/reference/Cargo.toml:145-146):
 *   - lib/src/primitives/eip4844.rs:44-48   get_evaluation_point
 *   - lib/src/primitives/eip4844.rs:50-65   proof_of_equivalence
 *   - lib/src/primitives/eip4844.rs:67-78   calc_kzg_proof[_with_point]
 *   - lib/src/primitives/eip4844.rs:80-89   calc_kzg_proof_commitment
 *   - lib/src/primitives/eip4844.rs:91-95   commitment_to_version_hash
 * and the rust-kzg algorithms behind them (un-vendored crate brechtpd/rust-kzg
 * @cbbfafdd, Cargo.lock:4167-4176,7499-7513; algorithms per SURVEY.md Appendix B =
 * Deneb polynomial-commitments spec): blob deserialisation with the canonical check,
 * Pippenger MSM without precomputation, TWO Montgomery batch inversions
 * (evaluation, quotient), the in-domain quotient case, zcash point compression.
 *
 * Parity pinning: validated byte-for-byte against tests/golden/kzg_golden.json
 * (pairing-validated by oracle/kzg_oracle.py; rows C1..C7 = SURVEY.md Appendix C,
 * which include the reference's only byte-level KAT, eip4844.rs:147-160).
 *
 * Arithmetic: 64-bit limbs with unsigned __int128, Montgomery R = 2^384 / 2^256 -- the
 * same representation the reference's settings images store, and
 * deliberately unlike the GPU code's 30-bit limbs so the two are independent.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

#define NPTS 4096
#define BLOB_BYTES 131072

/* ------------------------------------------------------------------ Fp (6 limbs) */
typedef struct { u64 l[6]; } fp;
static const fp FP_P = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                         0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static const u64 FP_INV = 0x89f3fffcfffcfffdULL; /* -p^-1 mod 2^64 */
static const fp FP_R2 = {{0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL,
                          0x67eb88a9939d83c0ULL, 0x9a793e85b519952dULL, 0x11988fe592cae3aaULL}};
static const fp FP_ONE = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                           0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};

static int fp_is_zero(const fp* a) { u64 o = 0; for (int i = 0; i < 6; i++) o |= a->l[i]; return o == 0; }
static int fp_eq(const fp* a, const fp* b) { u64 o = 0; for (int i = 0; i < 6; i++) o |= a->l[i] ^ b->l[i]; return o == 0; }
static int fp_geq(const fp* a, const fp* b) {
    for (int i = 5; i >= 0; i--) { if (a->l[i] > b->l[i]) return 1; if (a->l[i] < b->l[i]) return 0; }
    return 1;
}
static void fp_sub_nored(fp* r, const fp* a, const fp* b) {
    u64 borrow = 0;
    for (int i = 0; i < 6; i++) { u128 t = (u128)a->l[i] - b->l[i] - borrow; r->l[i] = (u64)t; borrow = (u64)(t >> 64) & 1; }
}
static void fp_add(fp* r, const fp* a, const fp* b) {
    u64 c = 0;
    for (int i = 0; i < 6; i++) { u128 t = (u128)a->l[i] + b->l[i] + c; r->l[i] = (u64)t; c = (u64)(t >> 64); }
    if (c || fp_geq(r, &FP_P)) fp_sub_nored(r, r, &FP_P);
}
static void fp_sub(fp* r, const fp* a, const fp* b) {
    u64 borrow = 0; fp t;
    for (int i = 0; i < 6; i++) { u128 d = (u128)a->l[i] - b->l[i] - borrow; t.l[i] = (u64)d; borrow = (u64)(d >> 64) & 1; }
    if (borrow) { u64 c = 0; for (int i = 0; i < 6; i++) { u128 s = (u128)t.l[i] + FP_P.l[i] + c; t.l[i] = (u64)s; c = (u64)(s >> 64); } }
    *r = t;
}
static void fp_neg(fp* r, const fp* a) { if (fp_is_zero(a)) *r = *a; else fp_sub_nored(r, &FP_P, a); }
static void fp_mul(fp* r, const fp* a, const fp* b) {
    u64 t[8] = {0};
    for (int i = 0; i < 6; i++) {
        u64 c = 0;
        for (int j = 0; j < 6; j++) { u128 s = (u128)a->l[j] * b->l[i] + t[j] + c; t[j] = (u64)s; c = (u64)(s >> 64); }
        u128 s = (u128)t[6] + c; t[6] = (u64)s; t[7] = (u64)(s >> 64);
        u64 m = t[0] * FP_INV;
        s = (u128)m * FP_P.l[0] + t[0]; c = (u64)(s >> 64);
        for (int j = 1; j < 6; j++) { s = (u128)m * FP_P.l[j] + t[j] + c; t[j - 1] = (u64)s; c = (u64)(s >> 64); }
        s = (u128)t[6] + c; t[5] = (u64)s; t[6] = t[7] + (u64)(s >> 64);
    }
    fp o; memcpy(o.l, t, 48);
    if (t[6] || fp_geq(&o, &FP_P)) fp_sub_nored(&o, &o, &FP_P);
    *r = o;
}
static void fp_sqr(fp* r, const fp* a) { fp_mul(r, a, a); }
static void fp_pow(fp* r, const fp* a, const u64* e, int nlimbs) {
    fp acc = FP_ONE;
    for (int i = nlimbs - 1; i >= 0; i--)
        for (int b = 63; b >= 0; b--) { fp_sqr(&acc, &acc); if ((e[i] >> b) & 1) fp_mul(&acc, &acc, a); }
    *r = acc;
}
static void fp_inv(fp* r, const fp* a) {
    u64 e[6]; memcpy(e, FP_P.l, 48); e[0] -= 2;
    fp_pow(r, a, e, 6);
}
static void fp_from_mont(fp* r, const fp* a) { fp one = {{1, 0, 0, 0, 0, 0}}; fp_mul(r, a, &one); }
static void fp_to_mont(fp* r, const fp* a) { fp_mul(r, a, &FP_R2); }
static void fp_from_be48(fp* r, const uint8_t* b) { /* canonical integer, not Montgomery */
    for (int i = 0; i < 6; i++) { u64 v = 0; for (int k = 0; k < 8; k++) v = (v << 8) | b[8 * (5 - i) + k]; r->l[i] = v; }
}
static void fp_to_be48(uint8_t* b, const fp* a) {
    for (int i = 0; i < 6; i++) for (int k = 0; k < 8; k++) b[8 * (5 - i) + k] = (uint8_t)(a->l[i] >> (56 - 8 * k));
}

/* ------------------------------------------------------------------ Fr (4 limbs) */
typedef struct { u64 l[4]; } fr;
static const fr FR_R = {{0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL}};
static const u64 FR_INV = 0xfffffffeffffffffULL;
static const fr FR_R2 = {{0xc999e990f3f29c6dULL, 0x2b6cedcb87925c23ULL, 0x05d314967254398fULL, 0x0748d9d99f59ff11ULL}};
static const fr FR_ONE = {{0x00000001fffffffeULL, 0x5884b7fa00034802ULL, 0x998c4fefecbc4ff5ULL, 0x1824b159acc5056fULL}};

static int fr_is_zero(const fr* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int fr_eq(const fr* a, const fr* b) { return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0; }
static int fr_geq(const fr* a, const fr* b) {
    for (int i = 3; i >= 0; i--) { if (a->l[i] > b->l[i]) return 1; if (a->l[i] < b->l[i]) return 0; }
    return 1;
}
static void fr_sub_nored(fr* r, const fr* a, const fr* b) {
    u64 borrow = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a->l[i] - b->l[i] - borrow; r->l[i] = (u64)t; borrow = (u64)(t >> 64) & 1; }
}
static void fr_add(fr* r, const fr* a, const fr* b) {
    u64 c = 0;
    for (int i = 0; i < 4; i++) { u128 t = (u128)a->l[i] + b->l[i] + c; r->l[i] = (u64)t; c = (u64)(t >> 64); }
    if (c || fr_geq(r, &FR_R)) fr_sub_nored(r, r, &FR_R);
}
static void fr_sub(fr* r, const fr* a, const fr* b) {
    u64 borrow = 0; fr t;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a->l[i] - b->l[i] - borrow; t.l[i] = (u64)d; borrow = (u64)(d >> 64) & 1; }
    if (borrow) { u64 c = 0; for (int i = 0; i < 4; i++) { u128 s = (u128)t.l[i] + FR_R.l[i] + c; t.l[i] = (u64)s; c = (u64)(s >> 64); } }
    *r = t;
}
static void fr_mul(fr* r, const fr* a, const fr* b) {
    u64 t[6] = {0};
    for (int i = 0; i < 4; i++) {
        u64 c = 0;
        for (int j = 0; j < 4; j++) { u128 s = (u128)a->l[j] * b->l[i] + t[j] + c; t[j] = (u64)s; c = (u64)(s >> 64); }
        u128 s = (u128)t[4] + c; t[4] = (u64)s; t[5] = (u64)(s >> 64);
        u64 m = t[0] * FR_INV;
        s = (u128)m * FR_R.l[0] + t[0]; c = (u64)(s >> 64);
        for (int j = 1; j < 4; j++) { s = (u128)m * FR_R.l[j] + t[j] + c; t[j - 1] = (u64)s; c = (u64)(s >> 64); }
        s = (u128)t[4] + c; t[3] = (u64)s; t[4] = t[5] + (u64)(s >> 64);
    }
    fr o; memcpy(o.l, t, 32);
    if (t[4] || fr_geq(&o, &FR_R)) fr_sub_nored(&o, &o, &FR_R);
    *r = o;
}
static void fr_pow(fr* r, const fr* a, const u64* e, int nlimbs) {
    fr acc = FR_ONE;
    for (int i = nlimbs - 1; i >= 0; i--)
        for (int b = 63; b >= 0; b--) { fr_mul(&acc, &acc, &acc); if ((e[i] >> b) & 1) fr_mul(&acc, &acc, a); }
    *r = acc;
}
static void fr_inv(fr* r, const fr* a) { u64 e[4]; memcpy(e, FR_R.l, 32); e[0] -= 2; fr_pow(r, a, e, 4); }
static void fr_from_mont(fr* r, const fr* a) { fr one = {{1, 0, 0, 0}}; fr_mul(r, a, &one); }
static void fr_to_mont(fr* r, const fr* a) { fr_mul(r, a, &FR_R2); }
/* 32 big-endian bytes -> canonical integer limbs; returns 1 when the value is >= r */
static int fr_from_be32(fr* r, const uint8_t* b) {
    for (int i = 0; i < 4; i++) { u64 v = 0; for (int k = 0; k < 8; k++) v = (v << 8) | b[8 * (3 - i) + k]; r->l[i] = v; }
    return fr_geq(r, &FR_R);
}
static void fr_to_be32(uint8_t* b, const fr* canon) {
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) b[8 * (3 - i) + k] = (uint8_t)(canon->l[i] >> (56 - 8 * k));
}
/* Montgomery's trick (upstream fr_batch_inv); inputs must be non-zero */
static void fr_batch_inv(fr* out, const fr* in, int n, fr* scratch) {
    fr acc = FR_ONE;
    for (int i = 0; i < n; i++) { scratch[i] = acc; fr_mul(&acc, &acc, &in[i]); }
    fr inv; fr_inv(&inv, &acc);
    for (int i = n - 1; i >= 0; i--) { fr t; fr_mul(&t, &inv, &scratch[i]); fr_mul(&inv, &inv, &in[i]); out[i] = t; }
}

/* ------------------------------------------------------------------ SHA-256 */
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha_block(uint32_t h[8], const uint8_t* p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = hh + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[i] + w[i];
        uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
void kzgref_sha256(const uint8_t* msg, size_t len, uint8_t out[32]) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t off = 0;
    for (; off + 64 <= len; off += 64) sha_block(h, msg + off);
    uint8_t tail[128] = {0};
    size_t rem = len - off;
    memcpy(tail, msg + off, rem);
    tail[rem] = 0x80;
    size_t tl = (rem + 9 <= 64) ? 64 : 128;
    u64 bits = (u64)len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha_block(h, tail);
    if (tl == 128) sha_block(h, tail + 64);
    for (int i = 0; i < 8; i++) { out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16); out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i]; }
}

/* ------------------------------------------------------------------ G1 (Jacobian) */
typedef struct { fp x, y; } g1a;            /* affine Montgomery, never infinity */
typedef struct { fp x, y, z; } g1j;         /* z == 0: infinity */

static void g1j_set_inf(g1j* r) { memset(r, 0, sizeof *r); }
static void g1j_dbl(g1j* r, const g1j* p) {
    if (fp_is_zero(&p->z)) { *r = *p; return; }
    fp a, b, c, d, e, f, t;
    fp_sqr(&a, &p->x); fp_sqr(&b, &p->y); fp_sqr(&c, &b);
    fp_add(&t, &p->x, &b); fp_sqr(&t, &t); fp_sub(&t, &t, &a); fp_sub(&t, &t, &c); fp_add(&d, &t, &t);
    fp_add(&e, &a, &a); fp_add(&e, &e, &a);
    fp_sqr(&f, &e);
    g1j o;
    fp_sub(&o.x, &f, &d); fp_sub(&o.x, &o.x, &d);
    fp_mul(&o.z, &p->y, &p->z); fp_add(&o.z, &o.z, &o.z);
    fp_sub(&t, &d, &o.x); fp_mul(&t, &e, &t);
    fp c8; fp_add(&c8, &c, &c); fp_add(&c8, &c8, &c8); fp_add(&c8, &c8, &c8);
    fp_sub(&o.y, &t, &c8);
    *r = o;
}
static void g1j_add_affine(g1j* r, const g1j* p, const g1a* q) {
    if (fp_is_zero(&p->z)) { r->x = q->x; r->y = q->y; r->z = FP_ONE; return; }
    fp z1z1, u2, s2, h, hh, hhh, rr, v, t;
    fp_sqr(&z1z1, &p->z); fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&s2, &q->y, &p->z); fp_mul(&s2, &s2, &z1z1);
    fp_sub(&h, &u2, &p->x); fp_sub(&rr, &s2, &p->y);
    if (fp_is_zero(&h)) { if (fp_is_zero(&rr)) { g1j_dbl(r, p); } else g1j_set_inf(r); return; }
    fp_sqr(&hh, &h); fp_mul(&hhh, &h, &hh); fp_mul(&v, &p->x, &hh);
    g1j o;
    fp_sqr(&o.x, &rr); fp_sub(&o.x, &o.x, &hhh); fp_sub(&o.x, &o.x, &v); fp_sub(&o.x, &o.x, &v);
    fp_sub(&t, &v, &o.x); fp_mul(&t, &rr, &t); fp_mul(&v, &p->y, &hhh); fp_sub(&o.y, &t, &v);
    fp_mul(&o.z, &p->z, &h);
    *r = o;
}
static void g1j_add(g1j* r, const g1j* p, const g1j* q) {
    if (fp_is_zero(&p->z)) { *r = *q; return; }
    if (fp_is_zero(&q->z)) { *r = *p; return; }
    fp z1z1, z2z2, u1, u2, s1, s2, h, hh, hhh, rr, v, t;
    fp_sqr(&z1z1, &p->z); fp_sqr(&z2z2, &q->z);
    fp_mul(&u1, &p->x, &z2z2); fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&s1, &p->y, &q->z); fp_mul(&s1, &s1, &z2z2);
    fp_mul(&s2, &q->y, &p->z); fp_mul(&s2, &s2, &z1z1);
    fp_sub(&h, &u2, &u1); fp_sub(&rr, &s2, &s1);
    if (fp_is_zero(&h)) { if (fp_is_zero(&rr)) g1j_dbl(r, p); else g1j_set_inf(r); return; }
    fp_sqr(&hh, &h); fp_mul(&hhh, &h, &hh); fp_mul(&v, &u1, &hh);
    g1j o;
    fp_sqr(&o.x, &rr); fp_sub(&o.x, &o.x, &hhh); fp_sub(&o.x, &o.x, &v); fp_sub(&o.x, &o.x, &v);
    fp_sub(&t, &v, &o.x); fp_mul(&t, &rr, &t); fp_mul(&v, &s1, &hhh); fp_sub(&o.y, &t, &v);
    fp_mul(&o.z, &p->z, &q->z); fp_mul(&o.z, &o.z, &h);
    *r = o;
}
/* ZG1::to_bytes */
static void g1j_compress(uint8_t out[48], const g1j* p) {
    if (fp_is_zero(&p->z)) { memset(out, 0, 48); out[0] = 0xc0; return; }
    fp zi, zi2, x, y;
    fp_inv(&zi, &p->z); fp_sqr(&zi2, &zi);
    fp_mul(&x, &p->x, &zi2); fp_mul(&y, &p->y, &zi2); fp_mul(&y, &y, &zi);
    fp_from_mont(&x, &x); fp_from_mont(&y, &y);
    fp_to_be48(out, &x);
    out[0] |= 0x80;
    /* y > (p-1)/2  <=>  2y > p - 1  <=>  2y >= p (p odd) */
    fp y2; u64 c = 0;
    for (int i = 0; i < 6; i++) { u128 t = (u128)y.l[i] + y.l[i] + c; y2.l[i] = (u64)t; c = (u64)(t >> 64); }
    if (c || fp_geq(&y2, &FP_P)) out[0] |= 0x20;
}
static int g1a_decompress(g1a* r, const uint8_t in[48]) {
    if (!(in[0] & 0x80) || (in[0] & 0x40)) return -1;
    uint8_t b[48]; memcpy(b, in, 48); b[0] &= 0x1f;
    fp xc; fp_from_be48(&xc, b);
    if (fp_geq(&xc, &FP_P)) return -2;
    fp x, rhs, y, four = {{4, 0, 0, 0, 0, 0}}, t;
    fp_to_mont(&x, &xc); fp_to_mont(&four, &four);
    fp_sqr(&rhs, &x); fp_mul(&rhs, &rhs, &x); fp_add(&rhs, &rhs, &four);
    /* sqrt = rhs^((p+1)/4) */
    u64 e[6]; u64 c = 1;
    for (int i = 0; i < 6; i++) { u128 s = (u128)FP_P.l[i] + c; e[i] = (u64)s; c = (u64)(s >> 64); }
    for (int i = 0; i < 6; i++) e[i] = (e[i] >> 2) | (i < 5 ? e[i + 1] << 62 : 0);
    fp_pow(&y, &rhs, e, 6);
    fp_sqr(&t, &y);
    if (!fp_eq(&t, &rhs)) return -3;
    fp yc, y2; fp_from_mont(&yc, &y);
    c = 0;
    for (int i = 0; i < 6; i++) { u128 s = (u128)yc.l[i] + yc.l[i] + c; y2.l[i] = (u64)s; c = (u64)(s >> 64); }
    int big = c || fp_geq(&y2, &FP_P);
    if (big != ((in[0] & 0x20) != 0)) fp_neg(&y, &y);
    r->x = x; r->y = y;
    return 0;
}

/* ------------------------------------------------------------------ settings */
typedef struct {
    g1a g1[NPTS];        /* Lagrange setup, bit-reversal order */
    fr roots[NPTS];      /* brp roots of unity, Montgomery */
    fr inv_n;            /* 1/4096 */
} kzgref_settings;

static void compute_roots(kzgref_settings* s) {
    /* w = 7^((r-1)/4096) */
    fr seven = {{7, 0, 0, 0}}, w;
    fr_to_mont(&seven, &seven);
    u64 e[4]; memcpy(e, FR_R.l, 32); e[0] -= 1;
    for (int i = 0; i < 4; i++) e[i] = (e[i] >> 12) | (i < 3 ? e[i + 1] << 52 : 0);
    fr_pow(&w, &seven, e, 4);
    fr* nat = (fr*)malloc(sizeof(fr) * NPTS);
    nat[0] = FR_ONE;
    for (int i = 1; i < NPTS; i++) fr_mul(&nat[i], &nat[i - 1], &w);
    for (int i = 0; i < NPTS; i++) {
        int rev = 0;
        for (int b = 0; b < 12; b++) rev |= ((i >> b) & 1) << (11 - b);
        s->roots[i] = nat[rev];
    }
    free(nat);
    fr n = {{4096, 0, 0, 0}};
    fr_to_mont(&n, &n);
    fr_inv(&s->inv_n, &n);
}

void* kzgref_load(const uint8_t* data, size_t len) {
    kzgref_settings* s = (kzgref_settings*)malloc(sizeof *s);
    if (!s) return NULL;
    size_t g1_off = 0; int ref = 0;
    if (len == 739624) { g1_off = 131080; ref = 1; }
    else if (len == 1001905) { g1_off = 393344 + 8; ref = 1; }
    else if (len >= 16 && memcmp(data, "RKZGTS02", 8) == 0) g1_off = 16;
    else { free(s); return NULL; }
    for (int i = 0; i < NPTS; i++) {
        if (ref) { memcpy(s->g1[i].x.l, data + g1_off + 144 * (size_t)i, 48); memcpy(s->g1[i].y.l, data + g1_off + 144 * (size_t)i + 48, 48); }
        else if (g1a_decompress(&s->g1[i], data + g1_off + 48 * (size_t)i)) { free(s); return NULL; }
    }
    compute_roots(s);
    if (ref) { /* the images carry the roots too: they must agree with the derived ones */
        size_t roff = len == 739624 ? 8 : 262264 + 8;
        if (memcmp(s->roots, data + roff, 32 * NPTS) != 0) { free(s); return NULL; }
    }
    return s;
}
void kzgref_free(void* s) { free(s); }

/* ------------------------------------------------------------------ MSM */
/* g1_lincomb: unsigned-window Pippenger, no precomputation, single thread */
static void g1_lincomb(g1j* out, const g1a* pts, const fr* scalars_canon, int n) {
    enum { C = 10, NB = (1 << C) - 1, NW = (255 + C - 1) / C };
    g1j* buckets = (g1j*)malloc(sizeof(g1j) * NB);
    g1j total; g1j_set_inf(&total);
    for (int w = NW - 1; w >= 0; w--) {
        for (int k = 0; k < C; k++) g1j_dbl(&total, &total);
        for (int b = 0; b < NB; b++) g1j_set_inf(&buckets[b]);
        int sh = w * C;
        for (int i = 0; i < n; i++) {
            const u64* l = scalars_canon[i].l;
            int li = sh >> 6, bo = sh & 63;
            u64 d = l[li] >> bo;
            if (bo + C > 64 && li + 1 < 4) d |= l[li + 1] << (64 - bo);
            d &= NB;
            if (d) g1j_add_affine(&buckets[d - 1], &buckets[d - 1], &pts[i]);
        }
        g1j run, acc; g1j_set_inf(&run); g1j_set_inf(&acc);
        for (int b = NB - 1; b >= 0; b--) { g1j_add(&run, &run, &buckets[b]); g1j_add(&acc, &acc, &run); }
        g1j_add(&total, &total, &acc);
    }
    free(buckets);
    *out = total;
}

/* ------------------------------------------------------------------ the path */
static int deserialize_blob(fr* out_canon, const uint8_t* blob) {   /* Blob::from_bytes + deserialize_blob_rust */
    for (int i = 0; i < NPTS; i++) if (fr_from_be32(&out_canon[i], blob + 32 * i)) return 2;
    return 0;
}
static void hash_to_bls_field(fr* canon, const uint8_t b[32]) {
    fr_from_be32(canon, b);
    while (fr_geq(canon, &FR_R)) fr_sub_nored(canon, canon, &FR_R);
}

void kzgref_versioned_hash(const uint8_t c[48], uint8_t out[32]) { kzgref_sha256(c, 48, out); out[0] = 0x01; }

void kzgref_evaluation_point(const uint8_t* blob, const uint8_t vh[32], uint8_t out_x[32]) {
    uint8_t buf[64], h[32];
    kzgref_sha256(blob, BLOB_BYTES, buf);
    memcpy(buf + 32, vh, 32);
    kzgref_sha256(buf, 64, h);
    fr x; hash_to_bls_field(&x, h);
    fr_to_be32(out_x, &x);
}

int kzgref_commit(const void* sv, const uint8_t* blob, size_t blob_len, uint8_t out[48]) {
    const kzgref_settings* s = (const kzgref_settings*)sv;
    if (blob_len != BLOB_BYTES) return 1;
    fr* p = (fr*)malloc(sizeof(fr) * NPTS);
    int rc = deserialize_blob(p, blob);
    if (!rc) { g1j c; g1_lincomb(&c, s->g1, p, NPTS); g1j_compress(out, &c); }
    free(p);
    return rc;
}

/* evaluate_polynomial_in_evaluation_form (App. B.3); p, z Montgomery */
static void eval_poly(fr* y, const kzgref_settings* s, const fr* pm, const fr* z, fr* tmp /* 3*NPTS */) {
    for (int i = 0; i < NPTS; i++) if (fr_eq(z, &s->roots[i])) { *y = pm[i]; return; }
    fr *d = tmp, *inv = tmp + NPTS, *scr = tmp + 2 * NPTS;
    for (int i = 0; i < NPTS; i++) fr_sub(&d[i], z, &s->roots[i]);
    fr_batch_inv(inv, d, NPTS, scr);
    fr acc = {{0, 0, 0, 0}};
    for (int i = 0; i < NPTS; i++) { fr t; fr_mul(&t, &inv[i], &s->roots[i]); fr_mul(&t, &t, &pm[i]); fr_add(&acc, &acc, &t); }
    fr zn = *z;
    for (int k = 0; k < 12; k++) fr_mul(&zn, &zn, &zn);
    fr_sub(&zn, &zn, &FR_ONE);
    fr_mul(&acc, &acc, &s->inv_n);
    fr_mul(y, &acc, &zn);
}

int kzgref_proof_of_equivalence(const void* sv, const uint8_t* blob, size_t blob_len, const uint8_t vh[32],
                                uint8_t out_x[32], uint8_t out_y[32]) {
    const kzgref_settings* s = (const kzgref_settings*)sv;
    if (blob_len != BLOB_BYTES) return 1;
    fr* buf = (fr*)malloc(sizeof(fr) * NPTS * 4);
    fr* p = buf;
    int rc = deserialize_blob(p, blob);
    if (!rc) {
        for (int i = 0; i < NPTS; i++) fr_to_mont(&p[i], &p[i]);
        kzgref_evaluation_point(blob, vh, out_x);
        fr z, y; fr_from_be32(&z, out_x); fr_to_mont(&z, &z);
        eval_poly(&y, s, p, &z, buf + NPTS);
        fr_from_mont(&y, &y); fr_to_be32(out_y, &y);
    }
    free(buf);
    return rc;
}

/* compute_kzg_proof_rust (App. B.4) */
int kzgref_compute_proof(const void* sv, const uint8_t* blob, size_t blob_len, const uint8_t z_be[32],
                         uint8_t out_proof[48], uint8_t out_y[32]) {
    const kzgref_settings* s = (const kzgref_settings*)sv;
    if (blob_len != BLOB_BYTES) return 1;
    fr* buf = (fr*)malloc(sizeof(fr) * NPTS * 5);
    fr *p = buf, *tmp = buf + NPTS, *q = buf + 4 * NPTS;
    int rc = deserialize_blob(p, blob);
    if (!rc) {
        for (int i = 0; i < NPTS; i++) fr_to_mont(&p[i], &p[i]);
        fr z, y; hash_to_bls_field(&z, z_be); fr_to_mont(&z, &z);
        eval_poly(&y, s, p, &z, tmp);
        /* second batch inversion: (w_i - z), slot m gets 1 */
        fr *d = tmp, *inv = tmp + NPTS, *scr = tmp + 2 * NPTS;
        int m = -1;
        for (int i = 0; i < NPTS; i++) {
            if (fr_eq(&z, &s->roots[i])) { m = i; d[i] = FR_ONE; } else fr_sub(&d[i], &s->roots[i], &z);
        }
        fr_batch_inv(inv, d, NPTS, scr);
        for (int i = 0; i < NPTS; i++) { fr t; fr_sub(&t, &p[i], &y); fr_mul(&q[i], &t, &inv[i]); }
        if (m >= 0) {
            for (int i = 0; i < NPTS; i++) {
                if (i == m) { d[i] = FR_ONE; continue; }
                fr t; fr_sub(&t, &z, &s->roots[i]); fr_mul(&d[i], &t, &z);
            }
            fr_batch_inv(inv, d, NPTS, scr);
            fr acc = {{0, 0, 0, 0}};
            for (int i = 0; i < NPTS; i++) {
                if (i == m) continue;
                fr t; fr_sub(&t, &p[i], &y); fr_mul(&t, &t, &s->roots[i]); fr_mul(&t, &t, &inv[i]); fr_add(&acc, &acc, &t);
            }
            q[m] = acc;
        }
        for (int i = 0; i < NPTS; i++) fr_from_mont(&q[i], &q[i]);
        g1j pi; g1_lincomb(&pi, s->g1, q, NPTS);
        g1j_compress(out_proof, &pi);
        if (out_y) { fr_from_mont(&y, &y); fr_to_be32(out_y, &y); }
    }
    free(buf);
    return rc;
}

/* commitment + versioned hash + raiko challenge + (x, y) + proof: the per-blob work of
 * preflight + run_prover (core/src/preflight.rs:260, core/src/interfaces.rs:207-219) */
int kzgref_commit_prove(const void* sv, const uint8_t* blob, size_t blob_len, uint8_t c[48], uint8_t vh[32],
                        uint8_t x[32], uint8_t y[32], uint8_t proof[48]) {
    int rc = kzgref_commit(sv, blob, blob_len, c);
    if (rc) return rc;
    kzgref_versioned_hash(c, vh);
    kzgref_evaluation_point(blob, vh, x);
    return kzgref_compute_proof(sv, blob, blob_len, x, proof, y);
}
