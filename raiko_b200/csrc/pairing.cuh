// BLS12-381 pairing product check (device + host), for verify_kzg_proof /
// verify_blob_kzg_proof_batch (the equation the
// reference tests at lib/src/primitives/eip4844.rs:176-183).
//
// Deliberately simple and obviously correct rather than fast -- two pairings close a whole
// batch, so this is not on the throughput path:
//   * Fq  : Fp kept strictly below 2p ("strict" wrappers over field30.cuh)
//   * Fp2 = Fq[i]/(i^2 + 1)
//   * Fp12 = Fq[w]/(w^12 - 2 w^6 + 2)  (i = w^6 - 1); schoolbook products
//   * G2 on the twist y^2 = x^3 + 4(1 + i), affine; lines scaled by w^3 (killed by the
//     final exponentiation because w^3 lies in Fp4)
//   * Miller loop over |x| = 0xd201000000010000, shared squarings for all pairs
//   * final exponentiation: easy part, then 3 * hard = (x-1)^2 (x+p)(x^2+p^2-1) + 3
// Only "product == 1" is needed, so the sign of x (a conjugation) is irrelevant.
#pragma once
#include "g1.cuh"

namespace rk {

// ---------------------------------------------------------------------------
// strict Fq: every value <= 2p
// ---------------------------------------------------------------------------
template <int K>
RK_HD void fe_cond_sub_k(Fp& a) {          // a >= K p  ?  a -= K p
    uint32_t t[FP_N];
    int32_t carry = 0;
#pragma unroll
    for (int i = 0; i < FP_N; i++) {
        int32_t d = (int32_t)a.v[i] - (int32_t)FpTag::modx(K, i) + carry;
        carry = d >> LIMB_BITS;
        t[i] = (uint32_t)d & LIMB_MASK;
    }
    if (carry == 0) {
#pragma unroll
        for (int i = 0; i < FP_N; i++) a.v[i] = t[i];
    }
}
RK_HD void fq_add(Fp& r, const Fp& a, const Fp& b) { fe_add(r, a, b); fe_cond_sub_k<2>(r); }
RK_HD void fq_sub(Fp& r, const Fp& a, const Fp& b) { fe_sub<FpTag, 2>(r, a, b); fe_cond_sub_k<2>(r); }
RK_HD void fq_neg(Fp& r, const Fp& a) { fe_neg<FpTag, 2>(r, a); }
RK_HD void fq_dbl(Fp& r, const Fp& a) { fq_add(r, a, a); }
RK_HD bool fq_is_zero(const Fp& a) { return fe_is_zero_mod(a); }
RK_HD bool fq_eq(const Fp& a, const Fp& b) { Fp t; fe_sub<FpTag, 2>(t, a, b); return fe_is_zero_mod(t); }

// ---------------------------------------------------------------------------
// Fp2
// ---------------------------------------------------------------------------
struct Fp2 { Fp c0, c1; };
RK_HD void fp2_add(Fp2& r, const Fp2& a, const Fp2& b) { fq_add(r.c0, a.c0, b.c0); fq_add(r.c1, a.c1, b.c1); }
RK_HD void fp2_sub(Fp2& r, const Fp2& a, const Fp2& b) { fq_sub(r.c0, a.c0, b.c0); fq_sub(r.c1, a.c1, b.c1); }
RK_HD void fp2_neg(Fp2& r, const Fp2& a) { fq_neg(r.c0, a.c0); fq_neg(r.c1, a.c1); }
RK_HD_NOINLINE void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b) {
    Fp v0, v1, s, t, u;
    fe_mul(v0, a.c0, b.c0);
    fe_mul(v1, a.c1, b.c1);
    fq_add(s, a.c0, a.c1);
    fq_add(t, b.c0, b.c1);
    fe_mul(u, s, t);
    fq_sub(u, u, v0);
    fq_sub(r.c1, u, v1);
    fq_sub(r.c0, v0, v1);
}
RK_HD_NOINLINE void fp2_sqr(Fp2& r, const Fp2& a) {
    Fp s, d, m;
    fq_add(s, a.c0, a.c1);
    fq_sub(d, a.c0, a.c1);
    fe_mul(m, a.c0, a.c1);
    fe_mul(r.c0, s, d);
    fq_add(r.c1, m, m);
}
RK_HD_NOINLINE void fp2_mul_fp(Fp2& r, const Fp2& a, const Fp& k) { fe_mul(r.c0, a.c0, k); fe_mul(r.c1, a.c1, k); }
RK_HD_NOINLINE void fp2_inv(Fp2& r, const Fp2& a) {
    Fp n0, n1, n, ninv;
    fe_sqr(n0, a.c0);
    fe_sqr(n1, a.c1);
    fq_add(n, n0, n1);
    fe_inv(ninv, n);
    fe_mul(r.c0, a.c0, ninv);
    fe_mul(n0, a.c1, ninv);
    fq_neg(r.c1, n0);
}
RK_HD bool fp2_is_zero(const Fp2& a) { return fq_is_zero(a.c0) && fq_is_zero(a.c1); }
RK_HD bool fp2_eq(const Fp2& a, const Fp2& b) { return fq_eq(a.c0, b.c0) && fq_eq(a.c1, b.c1); }

// ---------------------------------------------------------------------------
// G2 (twist, affine).  inf flag explicit.
// ---------------------------------------------------------------------------
struct G2Affine { Fp2 x, y; int inf; };

// on twist: y^2 == x^3 + 4(1+i)
RK_HD bool g2_on_curve(const G2Affine& q) {
    if (q.inf) return true;
    Fp2 l, r3, b;
    fp2_sqr(l, q.y);
    fp2_sqr(r3, q.x);
    fp2_mul(r3, r3, q.x);
    fe_const<FpTag, FP_B_COEFF>(b.c0);
    fe_const<FpTag, FP_B_COEFF>(b.c1);
    fp2_add(r3, r3, b);
    return fp2_eq(l, r3);
}

// ---------------------------------------------------------------------------
// Fp12
// ---------------------------------------------------------------------------
struct Fp12 { Fp c[12]; };

RK_HD void fp12_one(Fp12& r) {
    for (int i = 0; i < 12; i++) fe_zero(r.c[i]);
    fe_const<FpTag, FP_ONE>(r.c[0]);
}
RK_HD bool fp12_is_one(const Fp12& a) {
    Fp one;
    fe_const<FpTag, FP_ONE>(one);
    bool ok = fq_eq(a.c[0], one);
    for (int i = 1; i < 12; i++) ok = ok && fq_is_zero(a.c[i]);
    return ok;
}
// reduce 23 coefficients modulo w^12 = 2 w^6 - 2
RK_HD void fp12_fold(Fp12& r, Fp* t /* [23] */) {
    for (int k = 22; k >= 12; k--) {
        Fp d;
        fq_dbl(d, t[k]);
        fq_add(t[k - 6], t[k - 6], d);
        fq_sub(t[k - 12], t[k - 12], d);
    }
    for (int k = 0; k < 12; k++) r.c[k] = t[k];
}
RK_HD_NOINLINE void fp12_mul(Fp12& r, const Fp12& a, const Fp12& b) {
    Fp t[23];
    for (int k = 0; k < 23; k++) fe_zero(t[k]);
    for (int i = 0; i < 12; i++)
        for (int j = 0; j < 12; j++) {
            Fp m;
            fe_mul(m, a.c[i], b.c[j]);
            fq_add(t[i + j], t[i + j], m);
        }
    fp12_fold(r, t);
}
// square: 78 products instead of 144 (cross terms once, doubled)
RK_HD_NOINLINE void fp12_sqr(Fp12& r, const Fp12& a) {
    Fp t[23];
    for (int k = 0; k < 23; k++) fe_zero(t[k]);
    for (int i = 0; i < 12; i++) {
        Fp m;
        fe_sqr(m, a.c[i]);
        fq_add(t[2 * i], t[2 * i], m);
        for (int j = i + 1; j < 12; j++) {
            fe_mul(m, a.c[i], a.c[j]);
            fq_dbl(m, m);
            fq_add(t[i + j], t[i + j], m);
        }
    }
    fp12_fold(r, t);
}
// a * (l0 + l6 w^6 + l2 w^2 + l8 w^8 + l3 w^3): the shape of a line function
struct LineCoeffs { Fp l0, l6, l2, l8, l3; };
RK_HD_NOINLINE void fp12_mul_line(Fp12& r, const Fp12& a, const LineCoeffs& l) {
    Fp t[23];
    for (int k = 0; k < 23; k++) fe_zero(t[k]);
    for (int i = 0; i < 12; i++) {
        Fp m;
        fe_mul(m, a.c[i], l.l0); fq_add(t[i], t[i], m);
        fe_mul(m, a.c[i], l.l2); fq_add(t[i + 2], t[i + 2], m);
        fe_mul(m, a.c[i], l.l3); fq_add(t[i + 3], t[i + 3], m);
        fe_mul(m, a.c[i], l.l6); fq_add(t[i + 6], t[i + 6], m);
        fe_mul(m, a.c[i], l.l8); fq_add(t[i + 8], t[i + 8], m);
    }
    fp12_fold(r, t);
}
RK_HD void fp12_conj(Fp12& r, const Fp12& a) {      // Frobenius^6: w -> -w
    for (int i = 0; i < 12; i++) {
        if (i & 1) fq_neg(r.c[i], a.c[i]); else r.c[i] = a.c[i];
    }
}
// Frobenius p^k (k = 1, 2) through the constant table FP12_FROBk
template <class TAB>
RK_HD_NOINLINE uint32_t frob_word(int i) { return TAB::at(i); }     // one copy of the 312-word table per TAB
template <class TAB>
RK_HD_NOINLINE void fp12_frob(Fp12& r, const Fp12& a) {
    Fp t[12];
    for (int k = 0; k < 12; k++) fe_zero(t[k]);
    for (int j = 0; j < 12; j++) {
        Fp lo, hi, m;
        for (int l = 0; l < FP_N; l++) { lo.v[l] = frob_word<TAB>(j * 26 + l); hi.v[l] = frob_word<TAB>(j * 26 + 13 + l); }
        fe_mul(m, a.c[j], lo); fq_add(t[j % 6], t[j % 6], m);
        fe_mul(m, a.c[j], hi); fq_add(t[j % 6 + 6], t[j % 6 + 6], m);
    }
    for (int k = 0; k < 12; k++) r.c[k] = t[k];
}
// inverse by Gauss-Jordan on the 12 x 12 multiplication matrix (used once per check)
RK_HD_NOINLINE bool fp12_inv(Fp12& r, const Fp12& a) {
    Fp m[12][13];
    // column j of M = a * w^j ; solve M x = e0.  Build rows directly: M[row][col]
    Fp12 cur = a;
    for (int j = 0; j < 12; j++) {
        for (int row = 0; row < 12; row++) m[row][j] = cur.c[row];
        // cur *= w : shift up, fold w^12 = 2 w^6 - 2
        Fp top = cur.c[11], d;
        for (int k = 11; k >= 1; k--) cur.c[k] = cur.c[k - 1];
        fe_zero(cur.c[0]);
        fq_dbl(d, top);
        fq_add(cur.c[6], cur.c[6], d);
        fq_sub(cur.c[0], cur.c[0], d);
    }
    for (int row = 0; row < 12; row++) fe_zero(m[row][12]);
    fe_const<FpTag, FP_ONE>(m[0][12]);
    for (int col = 0; col < 12; col++) {
        int piv = -1;
        for (int row = col; row < 12; row++) if (!fq_is_zero(m[row][col])) { piv = row; break; }
        if (piv < 0) return false;
        if (piv != col) for (int k = 0; k < 13; k++) { Fp t = m[piv][k]; m[piv][k] = m[col][k]; m[col][k] = t; }
        Fp inv;
        fe_inv(inv, m[col][col]);
        for (int k = col; k < 13; k++) fe_mul(m[col][k], m[col][k], inv);
        for (int row = 0; row < 12; row++) {
            if (row == col || fq_is_zero(m[row][col])) continue;
            Fp f = m[row][col];
            for (int k = col; k < 13; k++) {
                Fp t;
                fe_mul(t, f, m[col][k]);
                fq_sub(m[row][k], m[row][k], t);
            }
        }
    }
    for (int k = 0; k < 12; k++) r.c[k] = m[k][12];
    return true;
}
// a^|x| (plain square and multiply, 64-bit exponent)
RK_HD_NOINLINE void fp12_pow_x(Fp12& r, const Fp12& a) {
    Fp12 acc = a;
    const uint64_t e = ((uint64_t)BLS_X_ABS_W32::at(1) << 32) | BLS_X_ABS_W32::at(0);
    for (int bit = 62; bit >= 0; bit--) {
        fp12_sqr(acc, acc);
        if ((e >> bit) & 1) fp12_mul(acc, acc, a);
    }
    r = acc;
}

// f^((p^12 - 1) / r * 3) == 1 ?   (the factor 3 is coprime to r)
RK_HD_NOINLINE bool final_exp_is_one(const Fp12& f) {
    Fp12 inv, t, f2, a, b, c, d;
    if (!fp12_inv(inv, f)) return false;
    fp12_conj(t, f);
    fp12_mul(t, t, inv);                       // f^(p^6 - 1)
    fp12_frob<FP12_FROB2>(f2, t);
    fp12_mul(f2, f2, t);                       // ^(p^2 + 1): now in the cyclotomic subgroup (inverse = conj)
    // a = f2^(x-1),  x - 1 = -(|x| + 1)
    fp12_pow_x(a, f2); fp12_mul(a, a, f2); fp12_conj(a, a);
    fp12_pow_x(b, a);  fp12_mul(b, b, a);  fp12_conj(b, b);        // b = f2^((x-1)^2)
    // c = b^(x + p)
    fp12_pow_x(c, b); fp12_conj(c, c);
    fp12_frob<FP12_FROB1>(t, b);
    fp12_mul(c, c, t);
    // d = c^(x^2 + p^2 - 1)
    fp12_pow_x(d, c); fp12_pow_x(d, d);                            // x^2 > 0
    fp12_frob<FP12_FROB2>(t, c);
    fp12_mul(d, d, t);
    fp12_conj(t, c);
    fp12_mul(d, d, t);
    // result = d * f2^3
    fp12_sqr(t, f2);
    fp12_mul(t, t, f2);
    fp12_mul(d, d, t);
    return fp12_is_one(d);
}

// ---------------------------------------------------------------------------
// Miller loop for up to NP pairs, product of f_{|x|,Q_k}(P_k)
// ---------------------------------------------------------------------------
RK_HD_NOINLINE void line_coeffs(LineCoeffs& l, const Fp2& lam, const Fp2& xt, const Fp2& yt, const Fp& xp, const Fp& yp) {
    Fp2 c0, c2;
    fp2_mul(c0, lam, xt);
    fp2_sub(c0, c0, yt);                       // lam*xT - yT
    fp2_mul_fp(c2, lam, xp);
    fp2_neg(c2, c2);                           // -lam*xP
    fq_sub(l.l0, c0.c0, c0.c1); l.l6 = c0.c1;  // a + b i  ->  (a - b) + b w^6
    fq_sub(l.l2, c2.c0, c2.c1); l.l8 = c2.c1;
    l.l3 = yp;
}

// Invert NP Fp2 denominators with one Fp2 inversion (Montgomery's trick); entries of dead
// pairs are skipped.  A zero denominator cannot occur for points of prime order r.
template <int NP>
RK_HD void fp2_batch_inv(Fp2* den, const bool* live) {
    Fp2 pre[NP], run;
    bool any = false;
    for (int k = 0; k < NP; k++) {
        if (!live[k]) continue;
        if (!any) { run = den[k]; any = true; pre[k] = den[k]; }     // pre[k] unused for the first live entry
        else { pre[k] = run; fp2_mul(run, run, den[k]); }
    }
    if (!any) return;
    Fp2 inv;
    fp2_inv(inv, run);
    int first = -1;
    for (int k = 0; k < NP; k++) if (live[k]) { first = k; break; }
    for (int k = NP - 1; k > first; k--) {
        if (!live[k]) continue;
        Fp2 t;
        fp2_mul(t, inv, pre[k]);          // 1 / den[k]
        fp2_mul(inv, inv, den[k]);
        den[k] = t;
    }
    den[first] = inv;
}

template <int NP>
RK_HD_NOINLINE void miller_loop(Fp12& f, const G1Affine* ps, const int* p_inf, const G2Affine* qs) {
    fp12_one(f);
    G2Affine t[NP];
    bool live[NP];
    for (int k = 0; k < NP; k++) { t[k] = qs[k]; live[k] = !p_inf[k] && !qs[k].inf; }
    const uint64_t e = ((uint64_t)BLS_X_ABS_W32::at(1) << 32) | BLS_X_ABS_W32::at(0);
    for (int bit = 62; bit >= 0; bit--) {
        fp12_sqr(f, f);
        Fp2 num[NP], den[NP];
        for (int k = 0; k < NP; k++) {            // tangent slopes 3x^2 / 2y, denominators inverted together
            if (!live[k]) continue;
            fp2_sqr(num[k], t[k].x);
            fp2_add(den[k], num[k], num[k]); fp2_add(num[k], den[k], num[k]);
            fp2_add(den[k], t[k].y, t[k].y);
        }
        fp2_batch_inv<NP>(den, live);
        for (int k = 0; k < NP; k++) {
            if (!live[k]) continue;
            Fp2 lam, x3, y3;
            fp2_mul(lam, num[k], den[k]);
            LineCoeffs l;
            line_coeffs(l, lam, t[k].x, t[k].y, ps[k].x, ps[k].y);
            fp12_mul_line(f, f, l);
            fp2_sqr(x3, lam); fp2_sub(x3, x3, t[k].x); fp2_sub(x3, x3, t[k].x);
            fp2_sub(y3, t[k].x, x3); fp2_mul(y3, lam, y3); fp2_sub(y3, y3, t[k].y);
            t[k].x = x3; t[k].y = y3;
        }
        if ((e >> bit) & 1) {
            for (int k = 0; k < NP; k++) {        // chord slopes (yQ - yT) / (xQ - xT)
                if (!live[k]) continue;
                fp2_sub(num[k], qs[k].y, t[k].y);
                fp2_sub(den[k], qs[k].x, t[k].x);
            }
            fp2_batch_inv<NP>(den, live);
            for (int k = 0; k < NP; k++) {
                if (!live[k]) continue;
                Fp2 lam, x3, y3;
                fp2_mul(lam, num[k], den[k]);
                LineCoeffs l;
                line_coeffs(l, lam, t[k].x, t[k].y, ps[k].x, ps[k].y);
                fp12_mul_line(f, f, l);
                fp2_sqr(x3, lam); fp2_sub(x3, x3, t[k].x); fp2_sub(x3, x3, qs[k].x);
                fp2_sub(y3, t[k].x, x3); fp2_mul(y3, lam, y3); fp2_sub(y3, y3, t[k].y);
                t[k].x = x3; t[k].y = y3;
            }
        }
    }
}

// prod_k e(P_k, Q_k) == 1 ?   P_k affine G1 (Montgomery, < 2p), Q_k affine G2 on the twist.
template <int NP>
RK_HD bool pairing_product_is_one(const G1Affine* ps, const int* p_inf, const G2Affine* qs) {
    Fp12 f;
    miller_loop<NP>(f, ps, p_inf, qs);
    return final_exp_is_one(f);
}

// 48-byte big-endian canonical -> strict Montgomery
RK_HD void fp_from_be48_mont(Fp& r, const uint8_t* in) {
    uint32_t w[12];
    for (int k = 0; k < 12; k++) w[k] = ((uint32_t)in[4 * (11 - k)] << 24) | ((uint32_t)in[4 * (11 - k) + 1] << 16) | ((uint32_t)in[4 * (11 - k) + 2] << 8) | in[4 * (11 - k) + 3];
    Fp c;
    fe_unpack<FpTag>(c, w);
    fe_to_mont(r, c);
}
// 192 bytes (x.c0 | x.c1 | y.c0 | y.c1, big-endian canonical) -> G2Affine
RK_HD void g2_from_be192(G2Affine& q, const uint8_t* in) {
    fp_from_be48_mont(q.x.c0, in);
    fp_from_be48_mont(q.x.c1, in + 48);
    fp_from_be48_mont(q.y.c0, in + 96);
    fp_from_be48_mont(q.y.c1, in + 144);
    q.inf = 0;
}
RK_HD void g2_neg(G2Affine& r, const G2Affine& q) { r = q; fp2_neg(r.y, q.y); }

}  // namespace rk

namespace rk {
// ===========================================================================
// Lane-parallel formulation of the same pairing check (BASELINE.json configs[4] is a
// throughput path: the single-thread check above cost 65-127 ms of a 150 ms batch verify).
//
// One warp; lane k < 12 owns coefficient k of every Fp12 value.  An Fp12 product is 144 Fp
// products: lane k forms the two raw coefficients t[k] and t[k + 12] (12 products, perfectly
// balanced: k + 1 and 11 - k terms), the fold modulo w^12 = 2 w^6 - 2 is a closed form per lane.
// Squares take 7 products per lane, a sparse line multiplication 5, a Frobenius map 2.
// The two G2 arguments of a KZG check are constants of the trusted setup ([s]G2 and the
// generator), so the Miller loop's line slopes are precomputed once per context
// (pairing_precompute_lines: 68 steps per point) and the loop itself has no G2 arithmetic left.
// Fp12 inversion uses the norm maps Fp12 -> Fp6 -> Fp2 -> Fp (one Fp inversion) instead of the
// 12 x 12 elimination.  The per-lane functions below are plain RK_HD code: the device kernel runs
// them on the lanes of a warp with shared-memory exchange, the host tests run them in a loop over
// k and compare every operation with the scalar implementation above.
// ===========================================================================
constexpr int PAIRING_STEPS = 68;                  // 63 doublings + 5 additions for |x| = 0xd201000000010000

struct LineStep { Fp l0, l6, a, b; };              // line = l0 + l6 w^6 + (a xP) w^2 + (b xP) w^8 + yP w^3

// The three functions below return LAZY sums: plain limb additions of products (each < 1.2p, cross
// terms of a square doubled), at most 12 of them, so <= 27p -- lp_collect reduces once, after the
// partial sums of all sub-lanes have been added.
// Sub-lanes.  The products of one lane are independent, so a lane may be played by S threads
// ("sub-lanes"): sub-lane s takes the products whose running number is congruent to s modulo S and
// the partial sums are added up afterwards (lp_collect).  S = 1: the whole lane, as before.
// S = 12 turns an Fp12 product into ONE Fp product per thread on 144 threads.

// lane k: raw coefficients t[k] (lo) and t[k + 12] (hi) of a * b
RK_HD_NOINLINE void fp12_mul_lane(int k, const Fp* a, const Fp* b, Fp& lo, Fp& hi, int s = 0, int S = 1) {
    fe_zero(lo); fe_zero(hi);
#pragma unroll 1
    for (int i = s; i < 12; i += S) {
        Fp m;
        if (i <= k) { fp_mul_sel<true>(m, a[i], b[k - i]); fe_add(lo, lo, m); }
        else { fp_mul_sel<true>(m, a[i], b[k + 12 - i]); fe_add(hi, hi, m); }
    }
}
// lane k: the same for a * a, cross terms once and doubled (7 products)
RK_HD_NOINLINE void fp12_sqr_lane(int k, const Fp* a, Fp& lo, Fp& hi, int s = 0, int S = 1) {
    fe_zero(lo); fe_zero(hi);
    int q = 0;
    for (int pass = 0; pass < 2; pass++) {
        const int n = pass == 0 ? k : k + 12;
        Fp& t = pass == 0 ? lo : hi;
        const int i0 = n > 11 ? n - 11 : 0;
#pragma unroll 1
        for (int i = i0; 2 * i <= n; i++, q++) {
            if (q % S != s) continue;
            const int j = n - i;
            Fp m;
            if (i == j) fp_sqr_sel<true>(m, a[i]);
            else { fp_mul_sel<true>(m, a[i], a[j]); fe_add(m, m, m); }
            fe_add(t, t, m);
        }
    }
}
// lane k: the same for a * (l0 + l2 w^2 + l3 w^3 + l6 w^6 + l8 w^8) (5 products)
RK_HD_NOINLINE void fp12_line_lane(int k, const Fp* a, const LineCoeffs& l, Fp& lo, Fp& hi, int s = 0, int S = 1) {
    fe_zero(lo); fe_zero(hi);
    const int es[5] = {0, 2, 3, 6, 8};
#pragma unroll 1
    for (int q = s; q < 5; q += S) {
        const int e = es[q];
        const Fp& c = e == 0 ? l.l0 : e == 2 ? l.l2 : e == 3 ? l.l3 : e == 6 ? l.l6 : l.l8;
        Fp m;
        if (e <= k) { fp_mul_sel<true>(m, a[k - e], c); fe_add(lo, lo, m); }
        else { fp_mul_sel<true>(m, a[k + 12 - e], c); fe_add(hi, hi, m); }
    }
}
// lane k: coefficient k after folding t[0..22] (t[23] ignored) modulo w^12 = 2 w^6 - 2.
//   w^(12+j) = 2 w^(6+j) - 2 w^j            (j <= 5)
//   w^(12+j) = 2 w^j - 4 w^(j-6)            (6 <= j <= 10)
RK_HD_NOINLINE void fp12_fold_lane(int k, const Fp* t, Fp& out) {
    // Lazy: plain limb additions / subtractions against a multiple of p that covers the subtrahend
    // (inputs <= 2p), ONE loose reduction and a conditional subtraction at the end -- half the limb
    // passes of doing every step in the strict (<= 2p) domain.
    Fp r, d;
    if (k <= 5) {
        fe_add(d, t[12 + k], t[12 + k]);                               // 2 T_k <= 4p
        fe_sub<FpTag, 4>(r, t[k], d);                                  // t_k - 2 T_k + 4p <= 6p
        if (k <= 4) {
            fe_add(d, t[18 + k], t[18 + k]); fe_add(d, d, d);          // 4 T_(k+6) <= 8p
            fe_sub<FpTag, 8>(r, r, d);                                 // <= 14p
        }
    } else {
        if (k <= 10) fe_add(d, t[6 + k], t[12 + k]);                   // T_(k-6) + T_k <= 4p
        else d = t[6 + k];
        fe_add(d, d, d);                                               // <= 8p
        fe_add(r, t[k], d);                                            // <= 10p
    }
    fe_reduce_loose<FpTag>(r);
    fe_cond_sub_k<2>(r);
    out = r;
}
// lane k: coefficient k of a^(p^e) (TAB = FP12_FROB1 / FP12_FROB2): two products
template <class TAB>
RK_HD_NOINLINE void fp12_frob_lane(int k, const Fp* a, Fp& out, int s = 0, int S = 1) {
    const int half = k >= 6 ? 13 : 0;
    Fp acc;
    fe_zero(acc);
#pragma unroll 1
    for (int q = s; q < 2; q += S) {
        const int j = (k % 6) + 6 * q;
        Fp c, m;
        for (int l = 0; l < FP_N; l++) c.v[l] = frob_word<TAB>(j * 26 + half + l);
        fp_mul_sel<true>(m, a[j], c);
        fq_add(acc, acc, m);
    }
    out = acc;
}
// the lines of the Miller loop for a FIXED G2 argument q (affine, on the twist, prime order)
RK_HD_NOINLINE void pairing_precompute_lines(LineStep* out /* [PAIRING_STEPS] */, const G2Affine& q) {
    G2Affine t = q;
    const uint64_t e = ((uint64_t)BLS_X_ABS_W32::at(1) << 32) | BLS_X_ABS_W32::at(0);
    int idx = 0;
    auto emit = [&](const Fp2& lam, const Fp2& xt, const Fp2& yt) {
        Fp2 c0;
        fp2_mul(c0, lam, xt);
        fp2_sub(c0, c0, yt);                       // lam xT - yT = a + b i -> (a - b) + b w^6
        LineStep& s = out[idx++];
        fq_sub(s.l0, c0.c0, c0.c1); s.l6 = c0.c1;
        Fp d;
        fq_sub(d, lam.c0, lam.c1);                 // -lam xP = (-(lam0 - lam1) xP) w^2 + (-lam1 xP) w^8
        fq_neg(s.a, d); fe_cond_sub_k<2>(s.a);
        fq_neg(s.b, lam.c1); fe_cond_sub_k<2>(s.b);
    };
    for (int bit = 62; bit >= 0; bit--) {
        Fp2 num, den, lam, x3, y3;
        fp2_sqr(num, t.x);
        fp2_add(den, num, num); fp2_add(num, den, num);            // 3 x^2
        fp2_add(den, t.y, t.y);
        fp2_inv(den, den);
        fp2_mul(lam, num, den);
        emit(lam, t.x, t.y);
        fp2_sqr(x3, lam); fp2_sub(x3, x3, t.x); fp2_sub(x3, x3, t.x);
        fp2_sub(y3, t.x, x3); fp2_mul(y3, lam, y3); fp2_sub(y3, y3, t.y);
        t.x = x3; t.y = y3;
        if ((e >> bit) & 1) {
            fp2_sub(num, q.y, t.y);
            fp2_sub(den, q.x, t.x);
            fp2_inv(den, den);
            fp2_mul(lam, num, den);
            emit(lam, t.x, t.y);
            fp2_sqr(x3, lam); fp2_sub(x3, x3, t.x); fp2_sub(x3, x3, q.x);
            fp2_sub(y3, t.x, x3); fp2_mul(y3, lam, y3); fp2_sub(y3, y3, t.y);
            t.x = x3; t.y = y3;
        }
    }
}

// ---- the check itself, written once against an "exchange" policy X: -------------------------
//   X::LANES              how many (lane, sub-lane) threads this invocation plays (device: 1 = this thread)
//   X::SUB                sub-lanes per lane (threads sharing the products of one coefficient)
//   X::lane(i), X::sub(i) lane and sub-lane index of the i-th of them
//   X::publish(slot, k, v) / X::sync() / X::read(slot) -> const Fp*   the 12-coefficient exchange
//   X::publish_part(s, n, v) / X::read_part(s) -> const Fp* [24]      partial raw coefficients of sub-lane s
// Device: lanes are threads, slots live in shared memory, sync = a warp or CTA barrier.  Host: one
// thread plays all of them phase by phase.  A distributed Fp12 value is `Fp v[X::LANES]`; the sub-lanes
// of a lane all hold the lane's coefficient.
template <class X>
struct LaneFp12 {
    Fp c[X::LANES];
};
// raw coefficients (lo, hi) of every played thread -> t[0..23] in slots 2 | 3
template <class X> RK_HD void lp_collect(X& x, const Fp* lo, const Fp* hi) {
    if (X::SUB == 1) {
        for (int i = 0; i < X::LANES; i++) {
            Fp l = lo[i], h = hi[i];                    // lazy sums (<= 27p) -> strict
            fe_reduce_loose<FpTag>(l); fe_cond_sub_k<2>(l);
            fe_reduce_loose<FpTag>(h); fe_cond_sub_k<2>(h);
            x.publish(2, x.lane(i), l); x.publish(3, x.lane(i), h);
        }
        x.sync();
        return;
    }
    for (int i = 0; i < X::LANES; i++) { x.publish_part(x.sub(i), x.lane(i), lo[i]); x.publish_part(x.sub(i), 12 + x.lane(i), hi[i]); }
    x.sync();
    for (int i = 0; i < X::LANES; i++) {
        const int j = x.sub(i);                         // sub-lane 0 adds up t[k], sub-lane 1 t[12 + k]
        if (j > 1) continue;
        const int n = 12 * j + x.lane(i);
        // the sub-lanes' lazy partial sums, <= 27p in total: plain limb additions (one carry pass each, no
        // comparison against the modulus), then ONE loose reduction and a conditional subtraction: below 2p
        Fp acc = x.read_part(0)[n];
        for (int s = 1; s < X::SUB; s++) fe_add(acc, acc, x.read_part(s)[n]);
        fe_reduce_loose<FpTag>(acc);
        fe_cond_sub_k<2>(acc);
        x.publish(2 + j, x.lane(i), acc);
    }
    x.sync();
}
template <class X> RK_HD_NOINLINE void lp_mul(X& x, LaneFp12<X>& r, const LaneFp12<X>& a, const LaneFp12<X>& b) {
    for (int i = 0; i < X::LANES; i++) if (x.sub(i) == 0) { x.publish(0, x.lane(i), a.c[i]); x.publish(1, x.lane(i), b.c[i]); }
    x.sync();
    Fp lo[X::LANES], hi[X::LANES];
    for (int i = 0; i < X::LANES; i++) fp12_mul_lane(x.lane(i), x.read(0), x.read(1), lo[i], hi[i], x.sub(i), X::SUB);
    lp_collect(x, lo, hi);
    for (int i = 0; i < X::LANES; i++) fp12_fold_lane(x.lane(i), x.read(2), r.c[i]);
    x.sync();
}
template <class X> RK_HD_NOINLINE void lp_sqr(X& x, LaneFp12<X>& r, const LaneFp12<X>& a) {
    for (int i = 0; i < X::LANES; i++) if (x.sub(i) == 0) x.publish(0, x.lane(i), a.c[i]);
    x.sync();
    Fp lo[X::LANES], hi[X::LANES];
    for (int i = 0; i < X::LANES; i++) fp12_sqr_lane(x.lane(i), x.read(0), lo[i], hi[i], x.sub(i), X::SUB);
    lp_collect(x, lo, hi);
    for (int i = 0; i < X::LANES; i++) fp12_fold_lane(x.lane(i), x.read(2), r.c[i]);
    x.sync();
}
template <class X> RK_HD_NOINLINE void lp_line(X& x, LaneFp12<X>& r, const LaneFp12<X>& a, const LineCoeffs& l) {
    for (int i = 0; i < X::LANES; i++) if (x.sub(i) == 0) x.publish(0, x.lane(i), a.c[i]);
    x.sync();
    Fp lo[X::LANES], hi[X::LANES];
    for (int i = 0; i < X::LANES; i++) fp12_line_lane(x.lane(i), x.read(0), l, lo[i], hi[i], x.sub(i), X::SUB);
    lp_collect(x, lo, hi);
    for (int i = 0; i < X::LANES; i++) fp12_fold_lane(x.lane(i), x.read(2), r.c[i]);
    x.sync();
}
template <class TAB, class X> RK_HD_NOINLINE void lp_frob(X& x, LaneFp12<X>& r, const LaneFp12<X>& a) {
    for (int i = 0; i < X::LANES; i++) if (x.sub(i) == 0) x.publish(0, x.lane(i), a.c[i]);
    x.sync();
    if (X::SUB == 1) {
        for (int i = 0; i < X::LANES; i++) fp12_frob_lane<TAB>(x.lane(i), x.read(0), r.c[i]);
        x.sync();
        return;
    }
    for (int i = 0; i < X::LANES; i++) {               // the two products on sub-lanes 0 and 1, every sub-lane adds them
        if (x.sub(i) > 1) continue;
        Fp part;
        fp12_frob_lane<TAB>(x.lane(i), x.read(0), part, x.sub(i), 2);
        x.publish_part(x.sub(i), x.lane(i), part);
    }
    x.sync();
    for (int i = 0; i < X::LANES; i++) fq_add(r.c[i], x.read_part(0)[x.lane(i)], x.read_part(1)[x.lane(i)]);
    x.sync();
}
template <class X> RK_HD void lp_conj(X& x, LaneFp12<X>& r, const LaneFp12<X>& a) {
    for (int i = 0; i < X::LANES; i++) { if (x.lane(i) & 1) fq_neg(r.c[i], a.c[i]); else r.c[i] = a.c[i]; }
}
template <class X> RK_HD_NOINLINE void lp_pow_x(X& x, LaneFp12<X>& r, const LaneFp12<X>& a) {
    LaneFp12<X> acc = a;
    const uint64_t e = ((uint64_t)BLS_X_ABS_W32::at(1) << 32) | BLS_X_ABS_W32::at(0);
    for (int bit = 62; bit >= 0; bit--) {
        lp_sqr(x, acc, acc);
        if ((e >> bit) & 1) lp_mul(x, acc, acc, a);
    }
    r = acc;
}
// 1 / a through the norm maps: with q = p^6, N6 = a * a^q lies in Fp6; N2 = N6 * N6^(p^2) * N6^(p^4)
// lies in Fp2 = {x + y w^6}; its inverse is ((x + 2y) - y w^6) / (x^2 + 2xy + 2y^2) because the
// conjugate of w^6 = 1 + i is 2 - w^6.  1/a = a^q * N6^(p^2) * N6^(p^4) / N2.  One Fp inversion.
// Returns false when a = 0.
template <class X> RK_HD_NOINLINE bool lp_inv(X& x, LaneFp12<X>& r, const LaneFp12<X>& a) {
    LaneFp12<X> aq, n6, f2, f4, m, n2;
    lp_conj(x, aq, a);
    lp_mul(x, n6, a, aq);
    lp_frob<FP12_FROB2>(x, f2, n6);
    lp_frob<FP12_FROB2>(x, f4, f2);
    lp_mul(x, m, f2, f4);
    lp_mul(x, n2, n6, m);                                   // = x + y w^6, every other coefficient 0
    for (int i = 0; i < X::LANES; i++) if (x.sub(i) == 0) x.publish(0, x.lane(i), n2.c[i]);
    x.sync();
    const Fp* n = x.read(0);
    Fp xx = n[0], yy = n[6], t, u, den, dinv;
    fe_sqr(t, xx);                                          // x^2
    fq_add(u, xx, yy); fq_dbl(u, u);                        // 2x + 2y
    fe_mul(den, yy, u);
    fq_add(den, den, t);                                    // x^2 + 2xy + 2y^2
    fe_cond_sub_k<2>(t);
    x.sync();
    if (fq_is_zero(den)) return false;
    fe_inv(dinv, den);
    Fp i0, i6;                                              // inverse of n2: i0 + i6 w^6
    fq_add(u, xx, yy); fq_add(u, u, yy);
    fe_mul(i0, u, dinv); fe_cond_sub_k<2>(i0);
    fe_mul(i6, yy, dinv); fq_neg(i6, i6); fe_cond_sub_k<2>(i6);
    LaneFp12<X> am;
    lp_mul(x, am, aq, m);
    // (am) * (i0 + i6 w^6): a two-term sparse product, done as a general product with a sparse operand
    LaneFp12<X> sp;
    for (int i = 0; i < X::LANES; i++) { fe_zero(sp.c[i]); if (x.lane(i) == 0) sp.c[i] = i0; if (x.lane(i) == 6) sp.c[i] = i6; }
    lp_mul(x, r, am, sp);
    return true;
}
template <class X> RK_HD bool lp_is_one(X& x, const LaneFp12<X>& a) {
    for (int i = 0; i < X::LANES; i++) if (x.sub(i) == 0) x.publish(0, x.lane(i), a.c[i]);
    x.sync();
    const Fp* c = x.read(0);
    Fp one;
    fe_const<FpTag, FP_ONE>(one);
    bool ok = fq_eq(c[0], one);
    for (int k = 1; k < 12; k++) ok = ok && fq_is_zero(c[k]);
    x.sync();
    return ok;
}
// e(P_0, Q_0) * e(P_1, Q_1) == 1 with the lines of Q_0, Q_1 precomputed (lines[pair][step]).
template <class X>
RK_HD bool lane_pairing_product_is_one(X& x, const G1Affine* ps, const int* p_inf, const LineStep* lines0, const LineStep* lines1) {
    LaneFp12<X> f;
    for (int i = 0; i < X::LANES; i++) { fe_zero(f.c[i]); if (x.lane(i) == 0) fe_const<FpTag, FP_ONE>(f.c[i]); }
    const uint64_t e = ((uint64_t)BLS_X_ABS_W32::at(1) << 32) | BLS_X_ABS_W32::at(0);
    int idx = 0;
    auto apply = [&](int step) {
        for (int k = 0; k < 2; k++) {
            if (p_inf[k]) continue;
            const LineStep& s = (k == 0 ? lines0 : lines1)[step];
            LineCoeffs l;
            l.l0 = s.l0; l.l6 = s.l6; l.l3 = ps[k].y;
            fe_mul(l.l2, s.a, ps[k].x); fe_cond_sub_k<2>(l.l2);
            fe_mul(l.l8, s.b, ps[k].x); fe_cond_sub_k<2>(l.l8);
            lp_line(x, f, f, l);
        }
    };
    for (int bit = 62; bit >= 0; bit--) {
        lp_sqr(x, f, f);
        apply(idx++);
        if ((e >> bit) & 1) apply(idx++);
    }
    // final exponentiation, as final_exp_is_one above
    LaneFp12<X> inv, t, f2, a, b, c, d;
    if (!lp_inv(x, inv, f)) return false;
    lp_conj(x, t, f);
    lp_mul(x, t, t, inv);
    lp_frob<FP12_FROB2>(x, f2, t);
    lp_mul(x, f2, f2, t);
    lp_pow_x(x, a, f2); lp_mul(x, a, a, f2); lp_conj(x, a, a);
    lp_pow_x(x, b, a);  lp_mul(x, b, b, a);  lp_conj(x, b, b);
    lp_pow_x(x, c, b); lp_conj(x, c, c);
    lp_frob<FP12_FROB1>(x, t, b);
    lp_mul(x, c, c, t);
    lp_pow_x(x, d, c); lp_pow_x(x, d, d);
    lp_frob<FP12_FROB2>(x, t, c);
    lp_mul(x, d, d, t);
    lp_conj(x, t, c);
    lp_mul(x, d, d, t);
    lp_sqr(x, t, f2);
    lp_mul(x, t, t, f2);
    lp_mul(x, d, d, t);
    return lp_is_one(x, d);
}

// host / single-thread exchange: one thread plays all 12 lanes (x S sub-lanes)
template <int S>
struct HostLanesT {
    static constexpr int SUB = S;
    static constexpr int LANES = 12 * S;
    Fp slot[4][12];                                // slots 2 and 3 are read as one 24-entry array (t[0..23])
    Fp part[S][24];
    RK_HD int lane(int i) const { return i % 12; }
    RK_HD int sub(int i) const { return i / 12; }
    RK_HD void publish(int s, int k, const Fp& v) { slot[s][k] = v; }
    RK_HD void publish_part(int s, int n, const Fp& v) { part[s][n] = v; }
    RK_HD void sync() {}
    RK_HD const Fp* read(int s) const { return slot[s]; }
    RK_HD const Fp* read_part(int s) const { return part[s]; }
};
using HostLanes = HostLanesT<1>;

}  // namespace rk
