// Host-side build of the product's arithmetic headers (field30.cuh, g1.cuh), so the
// limb algorithms can be checked against Python big integers without a GPU.
// TEST INFRASTRUCTURE: never linked into libraiko_kzg.so.
#include <cstring>
#include "../../raiko_b200/csrc/pairing.cuh"
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
void fc_fp_inv(const uint32_t* w, uint32_t* out) { Fp x, m, i, c; fe_unpack<FpTag>(x, w); fe_to_mont(m, x); fe_inv_fermat(i, m); fe_from_mont(c, i); fe_pack<FpTag>(out, c); }
void fc_fr_inv(const uint32_t* w, uint32_t* out) { Fr x, m, i, c; fe_unpack<FrTag>(x, w); fe_to_mont(m, x); fe_inv_fermat(i, m); fe_from_mont(c, i); fe_pack<FrTag>(out, c); }

// safegcd inversion: raw Montgomery limbs in (< 8 mod), raw Montgomery limbs out
void fc_fp_inv_safegcd(const uint32_t* a, uint32_t* r) { Fp x, z; memcpy(x.v, a, 52); fe_inv_safegcd(z, x); memcpy(r, z.v, 52); }
void fc_fr_inv_safegcd(const uint32_t* a, uint32_t* r) { Fr x, z; memcpy(x.v, a, 36); fe_inv_safegcd(z, x); memcpy(r, z.v, 36); }

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
// the batched-affine MSM's pieces: the slow-path addition, and one batch of affine additions
// sharing one safegcd inversion exactly as k_msm_affine's forward / backward passes do
// (acc[i] += pts[i], all distinct x), n <= 64
int fc_g1_affine_add_slow(const uint8_t* a, const uint8_t* b, uint8_t* out) {
    G1Affine p, q; if (g1_decompress(p, a) || g1_decompress(q, b)) return -1;
    if (!g1_affine_add_slow(p.x, p.y, q.x, q.y)) { g1_compress_inf(out); return 0; }
    g1_compress_affine(out, p); return 0; }
int fc_g1_affine_batch_add(const uint8_t* accs, const uint8_t* pts, int n, uint8_t* out) {
    G1Affine A[64], B[64]; Fp pre[64];
    if (n > 64) return -1;
    for (int i = 0; i < n; i++) if (g1_decompress(A[i], accs + 48 * i) || g1_decompress(B[i], pts + 48 * i)) return -1;
    Fp P; fe_const<FpTag, FP_ONE>(P);
    for (int k = 0; k < n; k++) { Fp d; fe_sub<FpTag, 2>(d, B[k].x, A[k].x); if (fe_is_zero_mod(d)) return -2; pre[k] = P; fe_mul(P, P, d); }
    Fp I; fe_inv_safegcd(I, P);
    for (int k = n - 1; k >= 0; k--) {
        Fp inv, d, lam, t, u;
        fe_mul(inv, I, pre[k]);
        fe_sub<FpTag, 2>(d, B[k].x, A[k].x); fe_mul(I, I, d);
        fe_add(u, A[k].x, B[k].x);
        fe_sub<FpTag, 2>(t, B[k].y, A[k].y);
        fe_mul(lam, t, inv);
        fe_sqr(t, lam);
        fe_sub<FpTag, 3>(t, t, u); fe_reduce_loose<FpTag>(t);
        fe_sub<FpTag, 2>(u, A[k].x, t); fe_mul(u, lam, u);
        fe_sub<FpTag, 2>(u, u, A[k].y); fe_reduce_loose<FpTag>(u);
        A[k].x = t; A[k].y = u;
        g1_compress_affine(out + 48 * k, A[k]);
    }
    return 0; }
// the fused limb passes of k_msm_affine's backward loop (RK_AFF_FUSE)
void fc_fp_sub_cneg4(const uint32_t* a, const uint32_t* b, int neg, uint32_t* r) { Fp x, y, z; memcpy(x.v, a, 52); memcpy(y.v, b, 52); fe_sub_cneg<FpTag, 4>(z, x, y, neg != 0); memcpy(r, z.v, 52); }
void fc_fp_sub_reduce3_rawsum(const uint32_t* a, const uint32_t* b1, const uint32_t* b2, uint32_t* r) {
    uint32_t tab[9 * FP_N]; fe_fill_multiples<FpTag>(tab);
    Fp x, y1, y2, u, z; memcpy(x.v, a, 52); memcpy(y1.v, b1, 52); memcpy(y2.v, b2, 52);
    fe_add_raw(u, y1, y2); fe_sub_reduce<FpTag, 3>(z, x, u, tab); memcpy(r, z.v, 52); }
void fc_fp_sub_reduce2(const uint32_t* a, const uint32_t* b, uint32_t* r) {
    uint32_t tab[9 * FP_N]; fe_fill_multiples<FpTag>(tab);
    Fp x, y, z; memcpy(x.v, a, 52); memcpy(y.v, b, 52); fe_sub_reduce<FpTag, 2>(z, x, y, tab); memcpy(r, z.v, 52); }
// one batch of affine additions exactly as the fused backward loop does them; negs[i] != 0 adds -pts[i]
int fc_g1_affine_batch_add_fused(const uint8_t* accs, const uint8_t* pts, const uint8_t* negs, int n, uint8_t* out) {
    G1Affine A[64], B[64]; Fp pre[64];
    uint32_t tab[9 * FP_N]; fe_fill_multiples<FpTag>(tab);
    if (n > 64) return -1;
    for (int i = 0; i < n; i++) if (g1_decompress(A[i], accs + 48 * i) || g1_decompress(B[i], pts + 48 * i)) return -1;
    Fp P; fe_const<FpTag, FP_ONE>(P);
    for (int k = 0; k < n; k++) { Fp d; fe_sub<FpTag, 2>(d, B[k].x, A[k].x); if (fe_is_zero_mod(d)) return -2; pre[k] = P; fe_mul(P, P, d); }
    Fp I; fe_inv_safegcd(I, P);
    for (int k = n - 1; k >= 0; k--) {
        Fp inv, d, lam, t, u;
        fe_mul(inv, I, pre[k]);
        fe_sub<FpTag, 2>(d, B[k].x, A[k].x); fe_mul(I, I, d);
        fe_add_raw(u, A[k].x, B[k].x);
        fe_sub_cneg<FpTag, 4>(t, B[k].y, A[k].y, negs[k] != 0);
        fe_mul(lam, t, inv);
        fe_sqr(t, lam);
        fe_sub_reduce<FpTag, 3>(t, t, u, tab);
        fe_sub<FpTag, 2>(u, A[k].x, t); fe_mul(u, lam, u);
        fe_sub_reduce<FpTag, 2>(u, u, A[k].y, tab);
        A[k].x = t; A[k].y = u;
        g1_compress_affine(out + 48 * k, A[k]);
    }
    return 0; }
void fc_fp_reduce_loose(const uint32_t* a, uint32_t* r) { Fp x; memcpy(x.v, a, 52); fe_reduce_loose<FpTag>(x); memcpy(r, x.v, 52); }
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

// ---- pairing (raiko_b200/csrc/pairing.cuh) -------------------------------------------------
static void be32_to_words(uint32_t k[8], const uint8_t* b) {
    for (int i = 0; i < 8; i++) k[7 - i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
}
// e(P1, Q1) * e(P2, Q2) == 1 ?  G1 compressed 48 B, G2 as 192 B (x.c0|x.c1|y.c0|y.c1 big-endian)
int fc_pairing_check2(const uint8_t* p1, const uint8_t* q1, const uint8_t* p2, const uint8_t* q2) {
    G1Affine ps[2]; int inf[2]; G2Affine qs[2];
    int rc = g1_decompress(ps[0], p1); if (rc < 0) return -1; inf[0] = rc == 1;
    rc = g1_decompress(ps[1], p2); if (rc < 0) return -1; inf[1] = rc == 1;
    g2_from_be192(qs[0], q1); g2_from_be192(qs[1], q2);
    if (!g2_on_curve(qs[0]) || !g2_on_curve(qs[1])) return -2;
    return pairing_product_is_one<2>(ps, inf, qs) ? 1 : 0;
}
// verify_kzg_proof: e(C - [y]G1 + [z]proof, G2) * e(-proof, [s]G2) == 1
int fc_verify_kzg_proof(const uint8_t* c48, const uint8_t* z32, const uint8_t* y32, const uint8_t* proof48,
                        const uint8_t* g2_gen, const uint8_t* g2_s) {
    G1Affine C, PI, G; int rc;
    rc = g1_decompress(C, c48); if (rc < 0) return -1; bool c_inf = rc == 1;
    rc = g1_decompress(PI, proof48); if (rc < 0) return -1; bool pi_inf = rc == 1;
    fe_const<FpTag, FP_GEN_X>(G.x); fe_const<FpTag, FP_GEN_Y>(G.y);
    uint32_t z[8], y[8];
    be32_to_words(z, z32); be32_to_words(y, y32);
    G1Xyzz acc, t;
    g1_set_inf(acc);
    if (!c_inf) g1_madd(acc, C.x, C.y);
    g1_scalar_mul(t, G, y);
    fe_neg<FpTag, 6>(t.y, t.y);                      // -[y]G  (t.y < 6p)
    g1_add(acc, t);
    if (!pi_inf) { g1_scalar_mul(t, PI, z); g1_add(acc, t); }
    G1Affine ps[2]; int inf[2]; G2Affine qs[2];
    inf[0] = !g1_xyzz_to_affine(ps[0], acc);
    ps[1] = PI; fe_neg<FpTag, 2>(ps[1].y, PI.y); inf[1] = pi_inf;
    g2_from_be192(qs[0], g2_gen); g2_from_be192(qs[1], g2_s);
    return pairing_product_is_one<2>(ps, inf, qs) ? 1 : 0;
}
// ---- lane-parallel pairing (pairing.cuh, second half) against the scalar implementation -------
static void rand_fp12(Fp12& a, uint32_t seed) {
    uint64_t s = seed * 0x9E3779B97F4A7C15ull + 12345;
    for (int k = 0; k < 12; k++) {
        Fp c;
        for (int i = 0; i < FP_N; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; c.v[i] = (uint32_t)(s >> 33) & LIMB_MASK; }
        c.v[FP_N - 1] &= 0x7ffff;                        // < 2^379 < p
        a.c[k] = c;
    }
}
static bool fp12_same(const Fp12& a, const Fp12& b) { for (int k = 0; k < 12; k++) if (!fq_eq(a.c[k], b.c[k])) return false; return true; }
}  // extern "C" (templates below)
template <class X> static void to_lanes(LaneFp12<X>& l, const Fp12& a) { for (int i = 0; i < X::LANES; i++) l.c[i] = a.c[i % 12]; }
// every sub-lane of a lane must hold the same coefficient; returns false when they differ
template <class X> static bool from_lanes(Fp12& a, const LaneFp12<X>& l) {
    for (int k = 0; k < 12; k++) a.c[k] = l.c[k];
    for (int i = 12; i < X::LANES; i++) if (!fq_eq(l.c[i], l.c[i % 12])) return false;
    return true;
}
// returns 0 when every lane-parallel Fp12 operation equals the scalar one on `iters` random inputs,
// else a code naming the first operation that differs (+100 when sub-lanes of one lane disagree)
template <int S> static int lane_fp12_ops(int iters) {
    using X = HostLanesT<S>;
    static X x;
    for (int it = 0; it < iters; it++) {
        Fp12 a, b, want, got;
        rand_fp12(a, 2 * it + 1); rand_fp12(b, 2 * it + 2);
        static LaneFp12<X> la, lb, lr;
        to_lanes(la, a); to_lanes(lb, b);
#define RK_SAME(code) do { if (!from_lanes(got, lr)) return 100 + code; if (!fp12_same(want, got)) return code; } while (0)
        fp12_mul(want, a, b); lp_mul(x, lr, la, lb); RK_SAME(1);
        fp12_sqr(want, a); lp_sqr(x, lr, la); RK_SAME(2);
        LineCoeffs l; l.l0 = b.c[0]; l.l2 = b.c[1]; l.l3 = b.c[2]; l.l6 = b.c[3]; l.l8 = b.c[4];
        fp12_mul_line(want, a, l); lp_line(x, lr, la, l); RK_SAME(3);
        fp12_frob<FP12_FROB1>(want, a); lp_frob<FP12_FROB1>(x, lr, la); RK_SAME(4);
        fp12_frob<FP12_FROB2>(want, a); lp_frob<FP12_FROB2>(x, lr, la); RK_SAME(5);
        fp12_conj(want, a); lp_conj(x, lr, la); RK_SAME(6);
        if (!fp12_inv(want, a)) return 7;
        if (!lp_inv(x, lr, la)) return 8;
        RK_SAME(9);
        if (it < 2) { fp12_pow_x(want, a); lp_pow_x(x, lr, la); RK_SAME(10); }
#undef RK_SAME
    }
    Fp12 z; for (int k = 0; k < 12; k++) fe_zero(z.c[k]);
    static LaneFp12<X> lz, lr; to_lanes(lz, z);
    if (lp_inv(x, lr, lz)) return 11;                    // 0 has no inverse
    return 0;
}
extern "C" {
// sub = sub-lanes per lane: 1 (one thread per coefficient), 5 and 12 (what the CTA-wide device kernel runs)
int fc_lane_fp12_ops(int iters, int sub) {
    return sub == 1 ? lane_fp12_ops<1>(iters) : sub == 5 ? lane_fp12_ops<5>(iters) : sub == 12 ? lane_fp12_ops<12>(iters) : -1;
}
// the same check as fc_pairing_check2 through the lane-parallel path with precomputed lines
int fc_pairing_check2_lanes(const uint8_t* p1, const uint8_t* q1, const uint8_t* p2, const uint8_t* q2, int sub) {
    G1Affine ps[2]; int inf[2]; G2Affine qs[2];
    int rc = g1_decompress(ps[0], p1); if (rc < 0) return -1; inf[0] = rc == 1;
    rc = g1_decompress(ps[1], p2); if (rc < 0) return -1; inf[1] = rc == 1;
    g2_from_be192(qs[0], q1); g2_from_be192(qs[1], q2);
    if (!g2_on_curve(qs[0]) || !g2_on_curve(qs[1])) return -2;
    static LineStep l0[PAIRING_STEPS], l1[PAIRING_STEPS];
    pairing_precompute_lines(l0, qs[0]);
    pairing_precompute_lines(l1, qs[1]);
    if (sub == 12) { static HostLanesT<12> x12; return lane_pairing_product_is_one(x12, ps, inf, l0, l1) ? 1 : 0; }
    HostLanes x;
    return lane_pairing_product_is_one(x, ps, inf, l0, l1) ? 1 : 0;
}
// subgroup membership of an affine point given as 96 bytes (x | y big-endian canonical), which need
// not be in G1: 1 / 0 by the endomorphism test, and by [r]P == infinity; -1 when not on the curve
static int load_xy(G1Affine& a, const uint8_t* xy) {
    fp_from_be48_mont(a.x, xy); fp_from_be48_mont(a.y, xy + 48);
    Fp l, r3, b4, t;
    fe_sqr(l, a.y); fe_sqr(r3, a.x); fe_mul(r3, r3, a.x);
    fe_const<FpTag, FP_B_COEFF>(b4); fe_add(r3, r3, b4);
    fe_sub<FpTag, 4>(t, l, r3);
    return fe_is_zero_mod(t) ? 0 : -1;
}
// k * P for a 256-bit k (8 little-endian words): the bitwise ladder and the 4-bit-window one of k_verify_terms
int fc_g1_scalar_mul(const uint8_t* a, const uint32_t* k, int windowed, uint8_t* out) {
    G1Affine q; if (g1_decompress(q, a)) return -1;
    G1Xyzz t;
    if (windowed) g1_scalar_mul_w4(t, q, k); else g1_scalar_mul(t, q, k);
    g1_compress(out, t); return 0; }
int fc_g1_in_subgroup_fast(const uint8_t* xy) { G1Affine a; if (load_xy(a, xy)) return -1; return g1_in_subgroup(a) ? 1 : 0; }
int fc_g1_in_subgroup_slow(const uint8_t* xy) {
    G1Affine a; if (load_xy(a, xy)) return -1;
    uint32_t k[8]; for (int w = 0; w < 8; w++) k[w] = FR_MOD_W32::at(w);
    G1Xyzz t; g1_scalar_mul(t, a, k); return g1_is_inf(t) ? 1 : 0;
}
}
