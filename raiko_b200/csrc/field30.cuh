// BLS12-381 Fp / Fr arithmetic on 30-bit limbs (sm_100a first, host-compilable).
//
// Why 30-bit limbs: on B200 a carry-less IMAD.WIDE.U32 (32x32+64 -> 64) issues at
// the full integer-multiply rate (measured 18.5 T/s, profiles/imad_ubench_r01.txt)
// while the carry-chained form that mad.lo.cc / madc.hi.cc compile to
// (IMAD.WIDE.U32.X with a predicate carry) runs at half of that.  With 30-bit
// limbs a column of 13 partial products fits a 64-bit accumulator, so the whole
// Montgomery product is 2*13^2+13 = 351 carry-less IMAD.WIDE/IMADs instead of
// 300 carry-chained ones (= 600 issue slots); carries are resolved with a few
// shifts on the otherwise idle ALU pipe.
//
// Representation: value = sum v[i] * 2^(30 i), every limb < 2^30 ("normalised").
// Values are kept only loosely reduced: 0 <= value < 8*mod unless stated.  The
// Montgomery radix is R = 2^(30 N) (2^390 for Fp, 2^270 for Fr) which leaves
// 9 / 15 spare bits, so products never need a final conditional subtraction:
//   mul(a, b) < a*b/R + mod   (a < 8 mod, b < 8 mod  =>  result < 1.11 mod).
//
// This file is the arithmetic the reference gets from the `bls12_381` crate
// through rust-kzg (call sites lib/src/primitives/eip4844.rs:54-87); it is a
// from-scratch formulation, not a translation of that crate's 64-bit limbs.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RK_HD __host__ __device__ __forceinline__
#define RK_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define RK_HD inline
#define RK_HD_NOINLINE static
#endif

#define RK_CONST_ARRAY(T, NAME, N, ...)                                  \
    struct NAME {                                                        \
        static constexpr int n = N;                                      \
        RK_HD static constexpr T at(int i) {                             \
            constexpr T t[N] = {__VA_ARGS__};                            \
            return t[i];                                                 \
        }                                                                \
    };
#include "kzg_constants.h"

namespace rk {

constexpr uint32_t LIMB_BITS = 30;
constexpr uint32_t LIMB_MASK = (1u << LIMB_BITS) - 1u;

// ---------------------------------------------------------------------------
// Field configuration tags
// ---------------------------------------------------------------------------
struct FpTag {
    static constexpr int N = FP_N;           // 13 limbs
    static constexpr int W32 = 12;           // packed 32-bit words
    static constexpr uint32_t PINV = FP_PINV;
    static constexpr uint32_t MINV = FP_MINV;
    using MOD = FP_MOD; using ONE = FP_ONE; using R2 = FP_R2;
    using FROM_REF = FP_FROM_REF; using TO_REF = FP_TO_REF;
    using EXP_INV = FP_EXP_INV;
    RK_HD static constexpr uint32_t modx(int k, int i) {
        return k == 1 ? FP_MOD::at(i) : k == 2 ? FP_MOD_X2::at(i) : k == 3 ? FP_MOD_X3::at(i)
             : k == 4 ? FP_MOD_X4::at(i) : k == 5 ? FP_MOD_X5::at(i) : k == 6 ? FP_MOD_X6::at(i)
             : k == 7 ? FP_MOD_X7::at(i) : FP_MOD_X8::at(i);
    }
};
struct FrTag {
    static constexpr int N = FR_N;           // 9 limbs
    static constexpr int W32 = 8;
    static constexpr uint32_t PINV = FR_PINV;
    static constexpr uint32_t MINV = FR_MINV;
    using MOD = FR_MOD; using ONE = FR_ONE; using R2 = FR_R2;
    using FROM_REF = FR_FROM_REF; using TO_REF = FR_TO_REF;
    using EXP_INV = FR_EXP_INV;
    RK_HD static constexpr uint32_t modx(int k, int i) {
        return k == 1 ? FR_MOD::at(i) : k == 2 ? FR_MOD_X2::at(i) : k == 3 ? FR_MOD_X3::at(i)
             : k == 4 ? FR_MOD_X4::at(i) : k == 5 ? FR_MOD_X5::at(i) : k == 6 ? FR_MOD_X6::at(i)
             : k == 7 ? FR_MOD_X7::at(i) : FR_MOD_X8::at(i);
    }
};

template <class F>
struct Fe {
    uint32_t v[F::N];
};

// c += a * b  (32 x 32 + 64 -> 64, no carry out).
// RK_MAC_FORM selects how the device code states it (measured, profiles/r01/):
//   0: C expression  -> ptxas fuses to IMAD.WIDE.U32 Rd, Ra, Rb, Rc.  With two register
//      multiplicands AND a 64-bit register addend this form issues at HALF rate on B200
//      (4 source words; 29/clk/SM vs 60/clk/SM for the immediate or RZ-addend forms).
//   1: PTX mad.wide.u32 -> ptxas splits it into IMAD.WIDE.U32 Rd, Ra, Rb, RZ (full rate)
//      plus 3-input 64-bit IADD3 / IADD3.X adds on the ALU pipe.
#ifndef RK_MAC_FORM
#define RK_MAC_FORM 0
#endif
RK_HD void mac(uint64_t& c, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && RK_MAC_FORM == 1
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b));
#else
    c += (uint64_t)a * b;
#endif
}
// same with a compile-time constant multiplicand (Montgomery reduction rows): the
// immediate form of IMAD.WIDE keeps its 64-bit addend at full rate.
RK_HD void mac_const(uint64_t& c, uint32_t a, uint32_t k) { c += (uint64_t)a * k; }

// Hide a 32-bit value's provenance from the optimiser (emits no instruction).
// Without this, LLVM sees `(uint64_t)(x & 0x3fffffff) * y`, rewrites the operand
// as a 64-bit AND, can no longer select mul.wide.u32, and ptxas leaves a dead
// "+ 0" IADD3 (the known-zero high cross term) behind every IMAD.WIDE.
RK_HD uint32_t launder(uint32_t x) {
#ifdef __CUDA_ARCH__
    asm("" : "+r"(x));
#endif
    return x;
}

using Fp = Fe<FpTag>;
using Fr = Fe<FrTag>;

// ---------------------------------------------------------------------------
// Montgomery product.  RK_MUL_FORM selects the formulation (all give the same integer result,
// (a b + M mod) / R with M fixed by a b mod R; host-checked in tests/test_field_host.py):
//   0  operand scanning into 2N 64-bit columns, then N reduction rows; the columns that would
//      exceed 64 bits are split at 30 bits first (round 1).
//   1  the same, but the split is at bit 32: the high WORD of a hot column moves up one column
//      (x4), which costs two instructions per column instead of five.
//   2  row-interleaved: product row i is followed at once by reduction row i and the window of
//      N+1 live columns slides down one limb, so the product needs 28 accumulator registers
//      instead of 52; one split pass (at bit 32) in the middle keeps every column below 2^64.
//   4  = 3 + 1: the library default.  5 = 3 with the square done the same way (77 instead of 91
//      multiply-accumulates; measured slower).
//   3  form 0 with one level of subtractive Karatsuba on the product rows (fe_mul only):
//      a = A0 + A1 X, X = 2^(30 H), H = ceil(N / 2):  a b = A0 B0 + X (A0 B0 + A1 B1 + (A0 - A1)(B1 - B0))
//      + X^2 A1 B1 -- 134 multiply-accumulates instead of 169 for N = 13, the differences are signed
//      31-bit limbs (no carries), their products go straight onto the columns (signed IMAD.WIDE,
//      two's-complement columns), the columns end up holding exactly what form 0 computes.
// ---------------------------------------------------------------------------
#ifndef RK_MUL_FORM
#define RK_MUL_FORM 4
#endif

// lo + 4 * up as one multiply-add (IMAD.WIDE with an immediate multiplicand is full rate; stated as a
// shift, the compiler spends five ALU instructions on the two words).
RK_HD uint64_t mad4(uint64_t lo, uint32_t up) {
#ifdef __CUDA_ARCH__
    uint64_t r;
    asm("mad.wide.u32 %0, %1, 4, %2;" : "=l"(r) : "r"(up), "l"(lo));
    return r;
#else
    return lo + ((uint64_t)up << 2);
#endif
}

// Carry-save split of columns c[lo..hi] at bit 32: column k keeps its low word and receives four
// times the high word of column k-1 (2^32 = 4 * 2^30); c[hi+1] receives the last one.  No ripple.
template <int LO, int HI, int LEN>
RK_HD void split_columns32(uint64_t (&c)[LEN]) {
    uint32_t up = 0;
#pragma unroll
    for (int k = LO; k <= HI; k++) {
        const uint32_t hi = (uint32_t)(c[k] >> 32);
        c[k] = mad4((uint32_t)c[k], up);
        up = hi;
    }
    c[HI + 1] = mad4(c[HI + 1], up);
}

// Montgomery reduction of 2N-1 product columns (shared by mul and sqr, forms 0 and 1)
//   c[k] < 2^63.71 on entry (at most N products of two 30-bit limbs).
template <class F>
RK_HD void mont_reduce_columns(Fe<F>& r, uint64_t (&c)[2 * F::N]) {
    constexpr int N = F::N;
    // A column k holds min(k+1, 2N-1-k) products now and receives as many m*mod
    // products below.  More than 7+7 could overflow 64 bits, so split those
    // columns first (carry-save: no ripple, every step independent).
    constexpr int HOT_LO = 7, HOT_HI = 2 * N - 2 - 7;
    if (HOT_LO <= HOT_HI) {
#if RK_MUL_FORM == 1 || RK_MUL_FORM == 4
        split_columns32<HOT_LO, HOT_HI, 2 * N>(c);
#else
        uint64_t carry_in = 0;
#pragma unroll
        for (int k = HOT_LO; k <= HOT_HI; k++) {
            uint64_t hi = c[k] >> LIMB_BITS;
            c[k] = (c[k] & LIMB_MASK) + carry_in;
            carry_in = hi;
        }
        c[HOT_HI + 1] += carry_in;
#endif
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t m = launder(((uint32_t)c[i] * F::PINV) & LIMB_MASK);
#pragma unroll
        for (int j = 0; j < N; j++) mac_const(c[i + j], m, F::MOD::at(j));
        c[i + 1] += c[i] >> LIMB_BITS;   // low 30 bits of c[i] are zero now
    }
    // result = columns N .. 2N-1, normalised
#pragma unroll
    for (int k = N; k < 2 * N - 1; k++) {
        c[k + 1] += c[k] >> LIMB_BITS;
        r.v[k - N] = launder((uint32_t)c[k] & LIMB_MASK);
    }
    r.v[N - 1] = launder((uint32_t)c[2 * N - 1]);
}

// Form 2 building blocks.  The window c[0..N] holds weights i .. i+N during row i.
template <class F>
RK_HD void mont_window_step(uint64_t (&c)[F::N + 1]) {
    constexpr int N = F::N;
    uint32_t m = launder(((uint32_t)c[0] * F::PINV) & LIMB_MASK);
#pragma unroll
    for (int j = 0; j < N; j++) mac_const(c[j], m, F::MOD::at(j));
    c[1] += c[0] >> LIMB_BITS;           // low 30 bits of c[0] are zero now
#pragma unroll
    for (int j = 0; j < N; j++) c[j] = c[j + 1];
    c[N] = 0;
}
template <class F>
RK_HD void mont_window_finish(Fe<F>& r, uint64_t (&c)[F::N + 1]) {
    constexpr int N = F::N;
#pragma unroll
    for (int k = 0; k < N - 1; k++) {
        c[k + 1] += c[k] >> LIMB_BITS;
        r.v[k] = launder((uint32_t)c[k] & LIMB_MASK);
    }
    r.v[N - 1] = launder((uint32_t)c[N - 1]);      // c[N] is zero: the result is below 2^(30 N)
}

// c += a * b for signed 32-bit operands (signed IMAD.WIDE; the column is read as two's complement)
RK_HD void mac_signed(uint64_t& c, int32_t a, int32_t b) { c += (uint64_t)((int64_t)a * (int64_t)b); }
RK_HD int32_t launder_s(int32_t x) {
#ifdef __CUDA_ARCH__
    asm("" : "+r"(x));
#endif
    return x;
}

template <class F>
RK_HD void fe_mul(Fe<F>& r, const Fe<F>& a, const Fe<F>& b) {
    constexpr int N = F::N;
#if RK_MUL_FORM == 3 || RK_MUL_FORM == 4 || RK_MUL_FORM == 5   // 4 = 3 with the hot-column split of form 1, 5 = 3 with a Karatsuba square too
    constexpr int H = (N + 1) / 2, L = N - H;         // low half H limbs, high half L <= H limbs
    uint64_t c[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) c[k] = 0;
    // A0 B0 -> columns 0 .. 2H-2;  A1 B1 -> columns 2H .. 2H+2L-2
#pragma unroll
    for (int i = 0; i < H; i++)
#pragma unroll
        for (int j = 0; j < H; j++) mac(c[i + j], a.v[i], b.v[j]);
#pragma unroll
    for (int i = 0; i < L; i++)
#pragma unroll
        for (int j = 0; j < L; j++) mac(c[2 * H + i + j], a.v[H + i], b.v[H + j]);
    // middle term, columns H .. 3H-2:  A0 B0 + A1 B1 (read before any of them is written) ...
    {
        uint64_t u[2 * H - 1];
#pragma unroll
        for (int j = 0; j < 2 * H - 1; j++) u[j] = c[j] + (j < 2 * L - 1 ? c[2 * H + j] : 0);
#pragma unroll
        for (int j = 0; j < 2 * H - 1; j++) c[H + j] += u[j];
    }
    // ... + (A0 - A1)(B1 - B0), limb-wise signed differences (|d| < 2^30), 7 products of < 2^60 per column
    {
        int32_t da[H], db[H];
#pragma unroll
        for (int i = 0; i < H; i++) {
            da[i] = launder_s((int32_t)a.v[i] - (i < L ? (int32_t)a.v[H + i] : 0));
            db[i] = launder_s((i < L ? (int32_t)b.v[H + i] : 0) - (int32_t)b.v[i]);
        }
#pragma unroll
        for (int i = 0; i < H; i++)
#pragma unroll
            for (int j = 0; j < H; j++) mac_signed(c[H + i + j], da[i], db[j]);
    }
    mont_reduce_columns<F>(r, c);
#elif RK_MUL_FORM == 2
    // Every row adds one product and one reduction product (< 2^60 each) to a column, and a
    // column lives for at most N rows: split once after row 6 (14 terms), the remaining
    // N - 7 <= 6 rows add at most 12 more.
    static_assert(N <= 13, "one split pass covers at most 13 rows");
    uint64_t c[N + 1];
#pragma unroll
    for (int k = 0; k <= N; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j < N; j++) mac(c[j], a.v[i], b.v[j]);
        mont_window_step<F>(c);
        if (i == 6 && N > 7) split_columns32<0, N - 1, N + 1>(c);
    }
    mont_window_finish<F>(r, c);
#else
    uint64_t c[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) mac(c[i + j], a.v[i], b.v[j]);
    mont_reduce_columns<F>(r, c);
#endif
}

template <class F>
RK_HD void fe_sqr(Fe<F>& r, const Fe<F>& a) {
    constexpr int N = F::N;
    uint32_t a2[N];
#pragma unroll
    for (int i = 0; i < N; i++) a2[i] = launder(a.v[i] << 1);
#if RK_MUL_FORM == 5
    // a^2 = A0^2 + X (A0^2 + A1^2 - (A0 - A1)^2) + X^2 A1^2: 77 multiply-accumulates instead of 91 for N = 13
    constexpr int H = (N + 1) / 2, L = N - H;
    uint64_t c[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
        mac(c[2 * i], a.v[i], a.v[i]);
#pragma unroll
        for (int j = i + 1; j < H; j++) mac(c[i + j], a2[i], a.v[j]);
    }
#pragma unroll
    for (int i = 0; i < L; i++) {
        mac(c[2 * H + 2 * i], a.v[H + i], a.v[H + i]);
#pragma unroll
        for (int j = i + 1; j < L; j++) mac(c[2 * H + i + j], a2[H + i], a.v[H + j]);
    }
    {
        uint64_t u[2 * H - 1];
#pragma unroll
        for (int j = 0; j < 2 * H - 1; j++) u[j] = c[j] + (j < 2 * L - 1 ? c[2 * H + j] : 0);
#pragma unroll
        for (int j = 0; j < 2 * H - 1; j++) c[H + j] += u[j];
    }
    {
        int32_t d[H], nd[H], nd2[H];
#pragma unroll
        for (int i = 0; i < H; i++) {
            d[i] = launder_s((int32_t)a.v[i] - (i < L ? (int32_t)a.v[H + i] : 0));     // |d| < 2^30
            nd[i] = launder_s(-d[i]);
            nd2[i] = launder_s(-2 * d[i]);                                               // |2d| < 2^31
        }
#pragma unroll
        for (int i = 0; i < H; i++) {
            mac_signed(c[H + 2 * i], nd[i], d[i]);
#pragma unroll
            for (int j = i + 1; j < H; j++) mac_signed(c[H + i + j], nd2[i], d[j]);
        }
    }
    mont_reduce_columns<F>(r, c);
#elif RK_MUL_FORM == 2
    // Row i adds a_i^2 (weight 2i) and the doubled products 2 a_i a_j, j > i: a column gathers its
    // product weight earlier than in the general product (up to 2 units of 2^60 per row from
    // products plus one from the reduction), so the window is split after rows 3 and 7.
    static_assert(N <= 13, "split schedule below is for at most 13 rows");
    uint64_t c[N + 1];
#pragma unroll
    for (int k = 0; k <= N; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        mac(c[i], a.v[i], a.v[i]);
#pragma unroll
        for (int j = i + 1; j < N; j++) mac(c[j], a2[i], a.v[j]);
        mont_window_step<F>(c);
        if ((i == 3 || i == 7) && i < N - 1) split_columns32<0, N - 1, N + 1>(c);
    }
    mont_window_finish<F>(r, c);
#else
    uint64_t c[2 * N];
#pragma unroll
    for (int k = 0; k < 2 * N; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        mac(c[2 * i], a.v[i], a.v[i]);
#pragma unroll
        for (int j = i + 1; j < N; j++) mac(c[i + j], a2[i], a.v[j]);
    }
    mont_reduce_columns<F>(r, c);
#endif
}

// r = a + b, normalised.  No modular reduction: the caller tracks the bound.
template <class F>
RK_HD void fe_add(Fe<F>& r, const Fe<F>& a, const Fe<F>& b) {
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        uint32_t t = a.v[i] + b.v[i] + carry;
        carry = t >> LIMB_BITS;
        r.v[i] = launder(t & LIMB_MASK);
    }
}

// r = a - b + K*mod  (K in 1..8, caller guarantees b <= K*mod so r >= 0)
template <class F, int K>
RK_HD void fe_sub(Fe<F>& r, const Fe<F>& a, const Fe<F>& b) {
    int32_t carry = 0;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        int32_t t = (int32_t)(a.v[i] + F::modx(K, i)) - (int32_t)b.v[i] + carry;
        carry = t >> LIMB_BITS;                 // arithmetic shift: -1, 0 or 1
        r.v[i] = launder((uint32_t)t & LIMB_MASK);
    }
}

// r = K*mod - a
template <class F, int K>
RK_HD void fe_neg(Fe<F>& r, const Fe<F>& a) {
    int32_t carry = 0;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        int32_t t = (int32_t)F::modx(K, i) - (int32_t)a.v[i] + carry;
        carry = t >> LIMB_BITS;
        r.v[i] = launder((uint32_t)t & LIMB_MASK);
    }
}

// r = a + b limb by limb, NOT normalised (limbs < 2^31).  Only for a value whose sole use is as the
// subtrahend of fe_sub / fe_sub_reduce below, whose carry pass absorbs the wide limbs.
template <class F>
RK_HD void fe_add_raw(Fe<F>& r, const Fe<F>& a, const Fe<F>& b) {
#pragma unroll
    for (int i = 0; i < F::N; i++) r.v[i] = a.v[i] + b.v[i];
}

// r = (neg ? -a : a) - b + K*mod in one carry pass (neg is per thread: no divergence).  The caller
// guarantees a + b <= K*mod.  b may be a raw sum (limbs < 2^31).
template <class F, int K>
RK_HD void fe_sub_cneg(Fe<F>& r, const Fe<F>& a, const Fe<F>& b, bool neg) {
    const uint32_t sb = neg ? 1u : 0u, sm = 0u - sb;
    int32_t carry = (int32_t)sb;                 // -a = ~a + 1 limb by limb: the +1 enters once, at limb 0 ...
    const uint32_t fix = sm & LIMB_MASK;         // ... and ~a_i = (a_i ^ mask30) - 2^30 + 2^30 per limb, see below
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        // neg: (a_i ^ 0x3fffffff) = 2^30 - 1 - a_i, so sum_i (2^30-1-a_i) 2^(30 i) + 1 = 2^(30 N) - a: the
        // 2^(30 N) leaves through the discarded carry out of the top limb (computed modulo 2^(30 N)).
        int32_t t = (int32_t)((a.v[i] ^ fix) + F::modx(K, i)) - (int32_t)b.v[i] + carry;
        carry = t >> LIMB_BITS;
        r.v[i] = launder((uint32_t)t & LIMB_MASK);
    }
}

// Multiples 0..8 of the modulus as normalised limbs, row-major [9][N]: the table fe_sub_reduce reads
// (kernels keep it in shared memory; fe_fill_multiples writes it).
template <class F>
RK_HD void fe_fill_multiples(uint32_t* tab) {
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        tab[i] = 0;
        tab[1 * F::N + i] = F::modx(1, i); tab[2 * F::N + i] = F::modx(2, i); tab[3 * F::N + i] = F::modx(3, i);
        tab[4 * F::N + i] = F::modx(4, i); tab[5 * F::N + i] = F::modx(5, i); tab[6 * F::N + i] = F::modx(6, i);
        tab[7 * F::N + i] = F::modx(7, i); tab[8 * F::N + i] = F::modx(8, i);
    }
}

// r = a - b + (K - q)*mod in ONE carry pass, q <= K estimated from the top limbs before the pass:
// the subtraction and the loose reduction that used to follow it (fe_sub<K> + fe_reduce_loose, two
// passes and a 64-bit multiply-subtract per limb).  a normalised, b normalised or a raw sum (limbs
// < 2^31), b <= K*mod, K <= 8.  Result in [0, max(mod + 12 * 2^(30 (N-1)), a)): q never exceeds
// value / mod, and it stops at K only when a - b itself is already that small.
// tab = fe_fill_multiples' table.
template <class F, int K>
RK_HD void fe_sub_reduce(Fe<F>& r, const Fe<F>& a, const Fe<F>& b, const uint32_t* tab) {
    constexpr int N = F::N;
    // top limb of the unreduced result, not counting the carry into it (which is in [-2, 1])
    const int32_t T = (int32_t)(a.v[N - 1] + F::modx(K, N - 1)) - (int32_t)b.v[N - 1];
    uint32_t q = T < 2 ? 0u : (uint32_t)(T - 2) / (F::MOD::at(N - 1) + 1u);      // q*mod <= value
    q = q > (uint32_t)K ? (uint32_t)K : q;
    const uint32_t* row = tab + ((uint32_t)K - q) * N;
    int32_t carry = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        int32_t t = (int32_t)(a.v[i] + row[i]) - (int32_t)b.v[i] + carry;
        carry = t >> LIMB_BITS;
        if (i < N - 1) r.v[i] = launder((uint32_t)t & LIMB_MASK);
        else r.v[i] = (uint32_t)t;
    }
}

template <class F>
RK_HD void fe_dbl(Fe<F>& r, const Fe<F>& a) { fe_add<F>(r, a, a); }

template <class F>
RK_HD void fe_set(Fe<F>& r, const Fe<F>& a) {
#pragma unroll
    for (int i = 0; i < F::N; i++) r.v[i] = a.v[i];
}
template <class F, class C>
RK_HD void fe_const(Fe<F>& r) {
#pragma unroll
    for (int i = 0; i < F::N; i++) r.v[i] = C::at(i);
}
template <class F>
RK_HD void fe_zero(Fe<F>& r) {
#pragma unroll
    for (int i = 0; i < F::N; i++) r.v[i] = 0;
}

// a == k*mod for some 0 <= k <= 8 ?  (a < 9*mod).  Cheap reject on limb 0.
template <class F>
RK_HD bool fe_is_zero_mod(const Fe<F>& a) {
    uint32_t k = (a.v[0] * F::MINV) & LIMB_MASK;   // a = k*mod  =>  a0 = k*mod0 (mod 2^30)
    if (k > 8) return false;
    bool eq = true;
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        uint64_t e = (uint64_t)k * F::MOD::at(i) + carry;     // < 9 * 2^30
        carry = e >> LIMB_BITS;
        eq = eq && (a.v[i] == ((uint32_t)e & LIMB_MASK));
    }
    return eq;
}

// a >= b as integers (both normalised)
template <class F>
RK_HD bool fe_geq_limbs(const Fe<F>& a, const uint32_t* b) {
    for (int i = F::N - 1; i >= 0; i--) {
        if (a.v[i] > b[i]) return true;
        if (a.v[i] < b[i]) return false;
    }
    return true;
}

// Fully reduce a value < 2*mod into [0, mod).
template <class F>
RK_HD void fe_cond_sub_mod(Fe<F>& a) {
    uint32_t t[F::N];
    int32_t carry = 0;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        int32_t d = (int32_t)a.v[i] - (int32_t)F::MOD::at(i) + carry;
        carry = d >> LIMB_BITS;
        t[i] = (uint32_t)d & LIMB_MASK;
    }
    if (carry == 0) {
#pragma unroll
        for (int i = 0; i < F::N; i++) a.v[i] = t[i];
    }
}

// Loose reduction of any normalised value (< 2^(30 N)): subtracts q * mod with the quotient
// estimated from the top limb, q = top / (mod_top + 1) <= a / mod.  Result in
// [0, mod * (1 + (q + 2) / mod_top)), i.e. below 1.0001 mod for the fields here.
template <class F>
RK_HD void fe_reduce_loose(Fe<F>& a) {
    constexpr int N = F::N;
    const uint32_t q = a.v[N - 1] / (F::MOD::at(N - 1) + 1u);
    int64_t carry = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        int64_t t = (int64_t)a.v[i] - (int64_t)((uint64_t)q * F::MOD::at(i)) + carry;
        if (i < N - 1) { a.v[i] = launder((uint32_t)t & LIMB_MASK); carry = t >> LIMB_BITS; }
        else a.v[i] = (uint32_t)t;
    }
}

// Montgomery -> canonical integer in [0, mod)
template <class F>
RK_HD void fe_from_mont(Fe<F>& r, const Fe<F>& a) {
    Fe<F> one;
    fe_zero<F>(one);
    one.v[0] = 1;
    fe_mul<F>(r, a, one);          // < a/R + mod < 2 mod
    fe_cond_sub_mod<F>(r);
}
// canonical (or any value < 8 mod) -> Montgomery
template <class F>
RK_HD void fe_to_mont(Fe<F>& r, const Fe<F>& a) {
    Fe<F> r2;
    fe_const<F, typename F::R2>(r2);
    fe_mul<F>(r, a, r2);
}

// 32-bit packed words <-> 30-bit limbs (value must be < 2^(32*W32))
template <class F>
RK_HD void fe_unpack(Fe<F>& r, const uint32_t* w) {
#pragma unroll
    for (int i = 0; i < F::N; i++) {
        const int bit = 30 * i, wi = bit >> 5, sh = bit & 31;
        uint32_t lo = wi < F::W32 ? w[wi] : 0u;
        uint32_t hi = (wi + 1) < F::W32 ? w[wi + 1] : 0u;
        uint32_t x = sh == 0 ? lo : ((lo >> sh) | (sh > 2 ? (hi << (32 - sh)) : 0u));
        r.v[i] = launder(x & LIMB_MASK);
    }
}
template <class F>
RK_HD void fe_pack(uint32_t* w, const Fe<F>& a) {
#pragma unroll
    for (int k = 0; k < F::W32; k++) {
        const int bit = 32 * k, li = bit / 30, sh = bit % 30;   // word k starts inside limb li
        uint32_t x = a.v[li] >> sh;
        if (li + 1 < F::N) x |= a.v[li + 1] << (30 - sh);
        w[k] = x;
    }
}

// r = a^e, e given as little-endian 32-bit words from a constant table.
// 4-bit fixed window: nbits squarings + nbits/4 multiplications.
template <class F, class EXP, int NW>
RK_HD_NOINLINE void fe_pow_const(Fe<F>& r, const Fe<F>& a) {
    Fe<F> tab[16];
    fe_const<F, typename F::ONE>(tab[0]);
    fe_set<F>(tab[1], a);
    for (int i = 2; i < 16; i++) fe_mul<F>(tab[i], tab[i - 1], a);
    Fe<F> acc;
    fe_const<F, typename F::ONE>(acc);
    for (int w = NW - 1; w >= 0; w--) {
        uint32_t word = EXP::at(w);
        for (int nib = 7; nib >= 0; nib--) {
            for (int s = 0; s < 4; s++) fe_sqr<F>(acc, acc);
            uint32_t d = (word >> (4 * nib)) & 15u;
            Fe<F> t;
            fe_set<F>(t, tab[d]);
            fe_mul<F>(acc, acc, t);
        }
    }
    fe_set<F>(r, acc);
}

// Fermat inversion (0 -> 0).  Input < 8 mod, output < 1.11 mod.  Kept as the independent
// cross-check of fe_inv_safegcd in the host tests; the product paths use fe_inv below.
template <class F>
RK_HD void fe_inv_fermat(Fe<F>& r, const Fe<F>& a) {
    fe_pow_const<F, typename F::EXP_INV, F::W32>(r, a);
}

// ---------------------------------------------------------------------------
// Inversion by Bernstein-Yang division steps ("safegcd", half-delta variant), 30 steps
// per outer iteration on the low limbs, the accumulated 2x2 transition matrix then applied
// to the full-width (f, g) and to (d, e) modulo the field -- the same signed-30-bit-limb
// shape as this file's products, so every product is again a carry-less IMAD.WIDE.
// 26-27 outer iterations for a 381-bit modulus (20 for 255 bits): ~35 Fp-mul equivalents,
// against ~440 for the Fermat ladder above.  That order of magnitude is what makes batched
// AFFINE point addition (g1.cuh: one shared inversion per few dozen additions) cheaper than
// inversion-free XYZZ addition in the MSM hot loop.
//
// Input: Montgomery form, < 8 mod, normalised limbs, != 0 mod p.  Output: Montgomery form of
// the inverse, fully reduced to [0, mod).  (d, e) start as (0, R^2), so the Montgomery
// factors cancel: (aR)^-1 * R^2 = a^-1 R.  Any multiple of the modulus (zero included) returns 0.
// ---------------------------------------------------------------------------
template <class F>
RK_HD void fe_inv_safegcd(Fe<F>& out, const Fe<F>& a) {
    constexpr int N = F::N;
    constexpr int32_t M30 = (int32_t)LIMB_MASK;
    int32_t d[N], e[N], f[N], g[N];
#pragma unroll
    for (int i = 0; i < N; i++) { d[i] = 0; e[i] = (int32_t)F::R2::at(i); f[i] = (int32_t)F::MOD::at(i); g[i] = (int32_t)a.v[i]; }
    int32_t zeta = -1;                              // -(delta + 1/2)
    for (int it = 0; it < 4 * N; it++) {            // proven bound for these sizes is below 37; the loop ends on g == 0
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < N; i++) nz |= (uint32_t)g[i];
        if (nz == 0) break;
        // ---- 30 division steps on the low words; (u v; q r) accumulates the transition matrix.
        // Stated with multiplies by -1/0/+1 instead of mask-and-add: one IMAD replaces the
        // xor / sub / and / add of a conditional signed accumulation (17 instructions per step
        // instead of 27).
        uint32_t u = 1, v = 0, q = 0, r = 1, fl = (uint32_t)f[0] | ((uint32_t)f[1] << 30), gl = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
#pragma unroll 6
        for (int s = 0; s < 30; s++) {
            const uint32_t odd = gl & 1u;                          // g odd?
            const uint32_t neg = (uint32_t)zeta >> 31;             // zeta < 0  (delta > 0)?
            const uint32_t sw = odd & neg;                         // swap step
            const uint32_t sg = odd - 2u * sw;                     // +1: g += f, -1 (as 2^32-1): g -= f, 0: g even
            gl += fl * sg; q += u * sg; r += v * sg;               // (g, q, r) += sg * (f, u, v)
            zeta = (int32_t)(((uint32_t)zeta ^ (0u - sw)) - 1u);   // swap: -zeta - 2, else zeta - 1
            fl += gl * sw; u += q * sw; v += r * sw;               // swap: (f, u, v) += new (g, q, r)
            gl >>= 1; u <<= 1; v <<= 1;
        }
        const int32_t U = (int32_t)u, V = (int32_t)v, Q = (int32_t)q, Rr = (int32_t)r;
        // ---- (d, e) <- (U d + V e, Q d + R e) / 2^30  modulo the field
        {
            const int32_t sd = d[N - 1] >> 31, se = e[N - 1] >> 31;
            int32_t md = (U & sd) + (V & se), me = (Q & sd) + (Rr & se);
            int64_t cd = (int64_t)U * d[0];
            cd += (int64_t)V * e[0];
            int64_t ce = (int64_t)Q * d[0];
            ce += (int64_t)Rr * e[0];
            md -= (int32_t)((F::MINV * (uint32_t)cd + (uint32_t)md) & LIMB_MASK);
            me -= (int32_t)((F::MINV * (uint32_t)ce + (uint32_t)me) & LIMB_MASK);
            cd += (int64_t)(int32_t)F::MOD::at(0) * md;
            ce += (int64_t)(int32_t)F::MOD::at(0) * me;
            cd >>= 30; ce >>= 30;
#pragma unroll
            for (int i = 1; i < N; i++) {
                // one multiply-accumulate per statement: each becomes a single IMAD.WIDE
                cd += (int64_t)U * d[i];
                cd += (int64_t)V * e[i];
                cd += (int64_t)(int32_t)F::MOD::at(i) * md;
                ce += (int64_t)Q * d[i];
                ce += (int64_t)Rr * e[i];
                ce += (int64_t)(int32_t)F::MOD::at(i) * me;
                d[i - 1] = (int32_t)cd & M30; cd >>= 30;
                e[i - 1] = (int32_t)ce & M30; ce >>= 30;
            }
            d[N - 1] = (int32_t)cd; e[N - 1] = (int32_t)ce;
        }
        // ---- (f, g) <- (U f + V g, Q f + R g) / 2^30  (exact)
        {
            int64_t cf = (int64_t)U * f[0];
            cf += (int64_t)V * g[0];
            int64_t cg = (int64_t)Q * f[0];
            cg += (int64_t)Rr * g[0];
            cf >>= 30; cg >>= 30;
#pragma unroll
            for (int i = 1; i < N; i++) {
                cf += (int64_t)U * f[i];
                cf += (int64_t)V * g[i];
                cg += (int64_t)Q * f[i];
                cg += (int64_t)Rr * g[i];
                f[i - 1] = (int32_t)cf & M30; cf >>= 30;
                g[i - 1] = (int32_t)cg & M30; cg >>= 30;
            }
            f[N - 1] = (int32_t)cf; g[N - 1] = (int32_t)cg;
        }
    }
    // f = +-gcd(a, mod) now: +-1 unless a was a multiple of the modulus, which maps to 0 like the
    // Fermat ladder's a^(mod-2).  Result = sign(f) * d, brought from (-2 mod, mod) into [0, mod).
    const int32_t neg = f[N - 1] >> 31;
    {
        // +1 = (1, 0, ..., 0);  -1 = (2^30-1, ..., 2^30-1, -1) with the signed top limb
        uint32_t dev = (uint32_t)f[0] ^ (neg ? LIMB_MASK : 1u);
#pragma unroll
        for (int i = 1; i < N - 1; i++) dev |= (uint32_t)f[i] ^ ((uint32_t)neg & LIMB_MASK);
        dev |= (uint32_t)f[N - 1] ^ (uint32_t)neg;
        if (dev != 0) {
#pragma unroll
            for (int i = 0; i < N; i++) d[i] = 0;
        }
    }
    int32_t add = d[N - 1] >> 31, carry = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        int32_t t = d[i] + ((int32_t)F::MOD::at(i) & add);
        t = (t ^ neg) - neg;
        t += carry;
        if (i < N - 1) { carry = t >> 30; t &= M30; }
        d[i] = t;
    }
    add = d[N - 1] >> 31; carry = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        int32_t t = d[i] + ((int32_t)F::MOD::at(i) & add) + carry;
        if (i < N - 1) { carry = t >> 30; t &= M30; }
        out.v[i] = (uint32_t)t;
    }
}

// Field inversion as the rest of the code calls it (0 -> 0).  Input < 8 mod, output in [0, mod).
template <class F>
RK_HD void fe_inv(Fe<F>& r, const Fe<F>& a) { fe_inv_safegcd<F>(r, a); }

}  // namespace rk
