// BLS12-381 G1 in extended Jacobian ("XYZZ") coordinates on top of field30.cuh.
//
// x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2.  Infinity is ZZ == 0 (all limbs zero,
// set explicitly -- a product of non-zero residues is never 0 mod p).
// Formulas: madd-2008-s / add-2008-s / dbl-2008-s-1 (a = 0).  Every exceptional
// case (infinity operands, P + P, P + (-P)) is handled, so the MSM built on
// these is exact for arbitrary scalars, which the bit-exact parity bar needs.
//
// Bounds (p = modulus): accumulator X, Y < 6p; ZZ, ZZZ < 1.1p; affine operand
// coordinates < 2p.  See field30.cuh for why no reduction is ever needed.
//
// Replaces the G1Projective arithmetic rust-kzg's g1_lincomb runs for
// blob_to_kzg_commitment_rust / compute_kzg_proof_rust
// (reference call sites lib/src/primitives/eip4844.rs:75,85).
#pragma once
#include "field30.cuh"

namespace rk {

struct G1Affine {
    Fp x, y;
};
struct G1Xyzz {
    Fp x, y, zz, zzz;
};

RK_HD void g1_set_inf(G1Xyzz& a) {
    fe_zero(a.x); fe_zero(a.y); fe_zero(a.zz); fe_zero(a.zzz);
}
RK_HD bool g1_is_inf(const G1Xyzz& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < FP_N; i++) o |= a.zz.v[i];
    return o == 0;
}
RK_HD void g1_from_affine(G1Xyzz& r, const G1Affine& p) {
    fe_set(r.x, p.x); fe_set(r.y, p.y);
    fe_const<FpTag, FP_ONE>(r.zz); fe_const<FpTag, FP_ONE>(r.zzz);
}

// Out-of-line Fp product / square.  Inlined, one mixed addition is ~85 KB of straight-line code
// (10 products), far beyond the instruction cache; called, a loop body is ~15 KB.  The MSM kernels
// inline (their warps run in lockstep and share instruction fetches); the latency-bound
// verification kernels (one or two warps per SM walking 255-bit scalars) call: template
// parameter CALLS of the point operations below.  The CUDA ABI keeps both 13-word operands and the result in registers (checked:
// no local-memory traffic, tools/ubench/noinline_test.cu), and the call costs ~3 %.
#ifdef __CUDACC__
static __device__ __noinline__ Fp fp_mul_call(Fp a, Fp b) { Fp r; fe_mul(r, a, b); return r; }
static __device__ __noinline__ Fp fp_sqr_call(Fp a) { Fp r; fe_sqr(r, a); return r; }
#endif
template <bool CALLS>
RK_HD void fp_mul_sel(Fp& r, const Fp& a, const Fp& b) {
#ifdef __CUDA_ARCH__
    if (CALLS) { r = fp_mul_call(a, b); return; }
#endif
    fe_mul(r, a, b);
}
template <bool CALLS>
RK_HD void fp_sqr_sel(Fp& r, const Fp& a) {
#ifdef __CUDA_ARCH__
    if (CALLS) { r = fp_sqr_call(a); return; }
#endif
    fe_sqr(r, a);
}

// r = 2 * (x, y, zz, zzz); works for affine input with zz = zzz = 1 too.
template <bool CALLS = false>
RK_HD void g1_dbl(G1Xyzz& r, const G1Xyzz& a) {
    if (g1_is_inf(a)) { g1_set_inf(r); return; }
    Fp U, V, Wv, S, M, t, X3;
    fe_dbl(U, a.y);                 // < 12p
    fp_sqr_sel<CALLS>(V, U);                   // < 1.3p
    fp_mul_sel<CALLS>(Wv, U, V);
    fp_mul_sel<CALLS>(S, a.x, V);
    fp_sqr_sel<CALLS>(t, a.x);
    fe_add(M, t, t); fe_add(M, M, t);      // 3 X^2 < 3.4p
    fp_sqr_sel<CALLS>(X3, M);
    fe_add(t, S, S);                        // 2S < 2.1p
    fe_sub<FpTag, 4>(X3, X3, t);            // < 5.1p
    fe_sub<FpTag, 6>(t, S, X3);             // < 7.1p
    fp_mul_sel<CALLS>(t, M, t);
    fp_mul_sel<CALLS>(U, Wv, a.y);                     // W*Y1 < 1.1p
    fp_mul_sel<CALLS>(r.zz, V, a.zz);
    fp_mul_sel<CALLS>(r.zzz, Wv, a.zzz);
    fe_sub<FpTag, 2>(r.y, t, U);
    fe_set(r.x, X3);
}

// acc += (x2, y2) affine (never infinity).
template <bool CALLS = false>
RK_HD void g1_madd(G1Xyzz& acc, const Fp& x2, const Fp& y2) {
    if (g1_is_inf(acc)) {
        fe_set(acc.x, x2); fe_set(acc.y, y2);
        fe_const<FpTag, FP_ONE>(acc.zz); fe_const<FpTag, FP_ONE>(acc.zzz);
        return;
    }
    Fp P, Rr, PP, PPP, Q, t;
    fp_mul_sel<CALLS>(P, x2, acc.zz);
    fp_mul_sel<CALLS>(Rr, y2, acc.zzz);
    fe_sub<FpTag, 6>(P, P, acc.x);          // < 7.1p
    fe_sub<FpTag, 6>(Rr, Rr, acc.y);
    if (fe_is_zero_mod(P)) {
        if (fe_is_zero_mod(Rr)) {
            G1Xyzz d;
            fe_set(d.x, x2); fe_set(d.y, y2);
            fe_const<FpTag, FP_ONE>(d.zz); fe_const<FpTag, FP_ONE>(d.zzz);
            g1_dbl<CALLS>(acc, d);
        } else {
            g1_set_inf(acc);
        }
        return;
    }
    fp_sqr_sel<CALLS>(PP, P);
    fp_mul_sel<CALLS>(PPP, P, PP);
    fp_mul_sel<CALLS>(Q, acc.x, PP);
    fp_sqr_sel<CALLS>(acc.x, Rr);           // X3 = R^2 - PPP - 2Q
    fe_add(t, Q, Q);
    fe_add(t, t, PPP);                      // < 3.1p
    fe_sub<FpTag, 4>(acc.x, acc.x, t);      // < 5.1p
    fe_sub<FpTag, 6>(t, Q, acc.x);          // < 7.1p
    fp_mul_sel<CALLS>(t, Rr, t);
    fp_mul_sel<CALLS>(Q, acc.y, PPP);
    fe_sub<FpTag, 2>(acc.y, t, Q);          // < 3.1p
    fp_mul_sel<CALLS>(acc.zz, acc.zz, PP);
    fp_mul_sel<CALLS>(acc.zzz, acc.zzz, PPP);
}

// a += b (both XYZZ)
template <bool CALLS = false>
RK_HD void g1_add(G1Xyzz& a, const G1Xyzz& b) {
    if (g1_is_inf(b)) return;
    if (g1_is_inf(a)) { a = b; return; }
    Fp U1, S1, P, Rr, PP, PPP, Q, t;
    fp_mul_sel<CALLS>(U1, a.x, b.zz);
    fp_mul_sel<CALLS>(P, b.x, a.zz);
    fp_mul_sel<CALLS>(S1, a.y, b.zzz);
    fp_mul_sel<CALLS>(Rr, b.y, a.zzz);
    fe_sub<FpTag, 2>(P, P, U1);             // < 3.1p
    fe_sub<FpTag, 2>(Rr, Rr, S1);
    if (fe_is_zero_mod(P)) {
        if (fe_is_zero_mod(Rr)) {
            G1Xyzz d = a;
            g1_dbl<CALLS>(a, d);
        } else {
            g1_set_inf(a);
        }
        return;
    }
    fp_sqr_sel<CALLS>(PP, P);
    fp_mul_sel<CALLS>(PPP, P, PP);
    fp_mul_sel<CALLS>(Q, U1, PP);
    fp_sqr_sel<CALLS>(a.x, Rr);
    fe_add(t, Q, Q);
    fe_add(t, t, PPP);
    fe_sub<FpTag, 4>(a.x, a.x, t);
    fe_sub<FpTag, 6>(t, Q, a.x);
    fp_mul_sel<CALLS>(t, Rr, t);
    fp_mul_sel<CALLS>(Q, S1, PPP);
    fe_sub<FpTag, 2>(a.y, t, Q);
    fp_mul_sel<CALLS>(a.zz, a.zz, b.zz);
    fp_mul_sel<CALLS>(a.zz, a.zz, PP);
    fp_mul_sel<CALLS>(a.zzz, a.zzz, b.zzz);
    fp_mul_sel<CALLS>(a.zzz, a.zzz, PPP);
}

// XYZZ -> affine Montgomery coordinates (< 1.1p); returns false for infinity.
// `zzz_inv` must be 1/ZZZ (Montgomery); callers batch that inversion.
RK_HD void g1_to_affine_with_inv(G1Affine& r, const G1Xyzz& a, const Fp& zzz_inv) {
    Fp t;
    fe_mul(t, a.zz, zzz_inv);      // ZZ/ZZZ
    fe_sqr(t, t);                  // = 1/ZZ   (ZZ^3 = ZZZ^2)
    fe_mul(r.x, a.x, t);
    fe_mul(r.y, a.y, zzz_inv);
}

// (x1, y1) += (x2, y2), both affine and finite, by way of the complete XYZZ addition and one
// inversion: the fallback of the batched-affine MSM for the additions its shared-inversion
// formula cannot do (equal x: doubling or cancellation).  Returns false when the sum is the
// point at infinity.  Output coordinates < 1.2p.
RK_HD_NOINLINE bool g1_affine_add_slow(Fp& x1, Fp& y1, const Fp& x2, const Fp& y2) {
    G1Xyzz a;
    fe_set(a.x, x1); fe_set(a.y, y1);
    fe_const<FpTag, FP_ONE>(a.zz); fe_const<FpTag, FP_ONE>(a.zzz);
    g1_madd(a, x2, y2);
    if (g1_is_inf(a)) return false;
    Fp inv;
    fe_inv_safegcd(inv, a.zzz);
    G1Affine r;
    g1_to_affine_with_inv(r, a, inv);
    fe_set(x1, r.x); fe_set(y1, r.y);
    return true;
}

// r = k * p, k given as 8 little-endian 32-bit words (plain double-and-add; variable-base
// work is off the throughput path: verification and tests only).
RK_HD_NOINLINE void g1_scalar_mul(G1Xyzz& r, const G1Affine& p, const uint32_t* k) {
    G1Xyzz acc;
    g1_set_inf(acc);
    for (int bit = 255; bit >= 0; bit--) {
        G1Xyzz t;
        g1_dbl<true>(t, acc);
        acc = t;
        if ((k[bit >> 5] >> (bit & 31)) & 1) g1_madd<true>(acc, p.x, p.y);
    }
    r = acc;
}
// The same with a fixed 4-bit window: 14 additions to build P .. 15P, then per nibble four
// doublings and at most one addition -- 256 doublings + <= 78 additions instead of 255 + ~128, and,
// what matters more on a GPU, lanes of a warp with different scalars no longer diverge on every bit
// (a warp executes a bit's addition if ANY lane has it set: 255 additions per warp in practice).
// Exceptional cases are g1_add's (complete).  k < 2^256.
RK_HD_NOINLINE void g1_scalar_mul_w4(G1Xyzz& r, const G1Affine& p, const uint32_t* k, int nibbles = 64) {
    G1Xyzz tab[15];                                   // tab[d - 1] = d * P (XYZZ)
    g1_from_affine(tab[0], p);
    g1_dbl<true>(tab[1], tab[0]);
    for (int d = 2; d < 15; d++) { tab[d] = tab[d - 1]; g1_madd<true>(tab[d], p.x, p.y); }
    G1Xyzz acc;
    g1_set_inf(acc);
    for (int nib = nibbles - 1; nib >= 0; nib--) {
        for (int s = 0; s < 4; s++) { G1Xyzz t; g1_dbl<true>(t, acc); acc = t; }
        const uint32_t d = (k[nib >> 3] >> (4 * (nib & 7))) & 15u;
        if (d) g1_add<true>(acc, tab[d - 1]);
    }
    r = acc;
}
// Membership of an affine curve point in the prime-order subgroup G1, by the endomorphism test
// (M. Scott, "A note on group membership tests for G1, G2 and GT on BLS pairing-friendly curves",
// ePrint 2021/1130): with u the BLS parameter (|u| = 0xd201000000010000) and r = u^4 - u^2 + 1,
// -u^2 is a cube root of unity mod r and acts on G1 as phi(x, y) = (BETA x, y); the paper proves
// that for BLS12-381   P in G1  <=>  [u^2] P = -phi(P) = (BETA x, -y).   Two 64-bit
// double-and-add passes (126 doublings + 10 additions) instead of a 255-bit multiplication by r:
// a third of the work.  The host tests check it against [r]P == infinity on points of G1, on random
// curve points and on points of every small prime order dividing the cofactor.
RK_HD_NOINLINE bool g1_in_subgroup(const G1Affine& p) {
    const uint64_t e = ((uint64_t)BLS_X_ABS_W32::at(1) << 32) | BLS_X_ABS_W32::at(0);
    G1Xyzz t, acc;
    g1_from_affine(acc, p);                           // [|u|] P, affine base
    for (int bit = 62; bit >= 0; bit--) {
        g1_dbl<true>(t, acc); acc = t;
        if ((e >> bit) & 1) g1_madd<true>(acc, p.x, p.y);
    }
    const G1Xyzz base = acc;                          // [|u|] ([|u|] P), projective base
    for (int bit = 62; bit >= 0; bit--) {
        g1_dbl<true>(t, acc); acc = t;
        if ((e >> bit) & 1) g1_add<true>(acc, base);
    }
    if (g1_is_inf(acc)) return false;                 // P is never infinity here
    // acc == (BETA x, -y) ?   X = BETA x ZZ,  Y = -y ZZZ
    Fp beta, bx, lhs, d;
    fe_const<FpTag, FP_BETA>(beta);
    fe_mul(bx, beta, p.x);
    fe_mul(lhs, bx, acc.zz);
    fe_sub<FpTag, 6>(d, lhs, acc.x);
    if (!fe_is_zero_mod(d)) return false;
    fe_mul(lhs, p.y, acc.zzz);
    fe_add(d, lhs, acc.y);                            // y ZZZ + Y == 0 ?
    return fe_is_zero_mod(d);
}

// XYZZ -> affine (one Fermat inversion).  Returns false for infinity.
RK_HD bool g1_xyzz_to_affine(G1Affine& r, const G1Xyzz& a) {
    if (g1_is_inf(a)) return false;
    Fp inv;
    fe_inv(inv, a.zzz);
    g1_to_affine_with_inv(r, a, inv);
    return true;
}

// ---------------------------------------------------------------------------
// Serialisation (reference ZG1::to_bytes via
// kzg_proof_to_bytes, lib/src/primitives/eip4844.rs:97-99)
// ---------------------------------------------------------------------------
// canonical integer limbs -> 48 big-endian bytes
RK_HD void fp_to_be48(uint8_t* out, const Fp& canon) {
    uint32_t w[12];
    fe_pack<FpTag>(w, canon);
#pragma unroll
    for (int k = 0; k < 12; k++) {
        uint32_t x = w[11 - k];
        out[4 * k + 0] = (uint8_t)(x >> 24);
        out[4 * k + 1] = (uint8_t)(x >> 16);
        out[4 * k + 2] = (uint8_t)(x >> 8);
        out[4 * k + 3] = (uint8_t)x;
    }
}

// x, y Montgomery (< 8p).  Writes the 48-byte zcash compressed form.
RK_HD void g1_compress_affine(uint8_t* out, const G1Affine& p) {
    Fp x, y;
    fe_from_mont(x, p.x);
    fe_from_mont(y, p.y);
    fp_to_be48(out, x);
    // y > (p-1)/2 ?
    bool gt = false, decided = false;
    for (int i = FP_N - 1; i >= 0; i--) {
        uint32_t h = FP_HALF::at(i);
        if (!decided && y.v[i] != h) { gt = y.v[i] > h; decided = true; }
    }
    out[0] |= gt ? 0xA0 : 0x80;
}
RK_HD void g1_compress_inf(uint8_t* out) {
    out[0] = 0xC0;
    for (int i = 1; i < 48; i++) out[i] = 0;
}

// Full conversion of one XYZZ point (one Fermat inversion; used per output).
RK_HD void g1_compress(uint8_t* out, const G1Xyzz& a) {
    if (g1_is_inf(a)) { g1_compress_inf(out); return; }
    Fp inv;
    fe_inv(inv, a.zzz);
    G1Affine aff;
    g1_to_affine_with_inv(aff, a, inv);
    g1_compress_affine(out, aff);
}

// 48 compressed bytes -> affine Montgomery.  Returns 0 ok, 1 infinity, <0 error.
// No subgroup check (the trusted setup is validated against its known digest).
RK_HD int g1_decompress(G1Affine& r, const uint8_t* in) {
    if (!(in[0] & 0x80)) return -1;
    if (in[0] & 0x40) {
        uint32_t o = in[0] & 0x3F;
        for (int i = 1; i < 48; i++) o |= in[i];
        return o ? -2 : 1;
    }
    uint32_t w[12];
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const uint8_t* b = in + 4 * (11 - k);
        uint32_t first = (k == 11) ? (uint32_t)(b[0] & 0x1F) : (uint32_t)b[0];
        w[k] = (first << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3];
    }
    Fp xc;
    fe_unpack<FpTag>(xc, w);
    // x < p ?
    {
        bool lt = false, decided = false;
        for (int i = FP_N - 1; i >= 0; i--) {
            uint32_t m = FP_MOD::at(i);
            if (!decided && xc.v[i] != m) { lt = xc.v[i] < m; decided = true; }
        }
        if (!lt) return -3;
    }
    Fp x, rhs, y, t, b4;
    fe_to_mont(x, xc);
    fe_sqr(rhs, x);
    fe_mul(rhs, rhs, x);
    fe_const<FpTag, FP_B_COEFF>(b4);
    fe_add(rhs, rhs, b4);                          // x^3 + 4  (< 2.1p)
    fe_pow_const<FpTag, FP_EXP_SQRT, 12>(y, rhs);
    fe_sqr(t, y);
    fe_sub<FpTag, 4>(t, t, rhs);
    if (!fe_is_zero_mod(t)) return -4;             // not on the curve
    Fp yc;
    fe_from_mont(yc, y);
    bool gt = false, decided = false;
    for (int i = FP_N - 1; i >= 0; i--) {
        uint32_t h = FP_HALF::at(i);
        if (!decided && yc.v[i] != h) { gt = yc.v[i] > h; decided = true; }
    }
    bool want = (in[0] & 0x20) != 0;
    if (gt != want) fe_neg<FpTag, 2>(y, y);
    fe_set(r.x, x);
    fe_set(r.y, y);
    return 0;
}

}  // namespace rk
