// SHA-256 (FIPS 180-4) for device and host.  The reference hashes through the
// `sha2` crate: sha256(blob), sha256(blob_hash || versioned_hash)
// (lib/src/primitives/eip4844.rs:44-48) and sha256(commitment)
// (eip4844.rs:91-95).  One thread owns one message here; blobs are independent
// so a batch gives the device one chain per thread, and the hash kernel's ALU /
// shift work overlaps the MSM kernel's integer-multiply work on another stream.
#pragma once
#include <stdint.h>
#include "field30.cuh"

namespace rk {

RK_HD uint32_t sha_rotr(uint32_t x, int n) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(x, x, n);
#else
    return (x >> n) | (x << (32 - n));
#endif
}

RK_CONST_ARRAY(uint32_t, SHA256_K, 64,
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u)

struct Sha256State {
    uint32_t h[8];
};

RK_HD void sha256_init(Sha256State& s) {
    s.h[0] = 0x6a09e667u; s.h[1] = 0xbb67ae85u; s.h[2] = 0x3c6ef372u; s.h[3] = 0xa54ff53au;
    s.h[4] = 0x510e527fu; s.h[5] = 0x9b05688cu; s.h[6] = 0x1f83d9abu; s.h[7] = 0x5be0cd19u;
}

// a + b on the multiply pipe (IMAD.IADD) instead of the ALU pipe.  One warp alone on a scheduler --
// the rounds half of the two-warp hash -- is bound by the ALU pipe's issue rate (a warp instruction
// every two cycles), not by its dependent chain: of the ~14 ALU instructions of a round only the six
// rotates and four three-input logic ops have to be there, so the additions go to the other pipe.
// `one` must be the value 1 in a register whose content neither the compiler nor the assembler can
// know (the kernels pass `nblobs > 0`): with a literal 1 ptxas folds pairs of these back into
// three-input IADD3s on the ALU pipe.
RK_HD uint32_t sha_add_fma(uint32_t a, uint32_t b, uint32_t one) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(b));
    return r;
#else
    return a * one + b;
#endif
}

// One round, written so that the dependent chain through e (and a) is short -- rotate, 3-input xor,
// two additions -- instead of the six steps of the textbook statement: h and d are values of e and a
// from three rounds earlier, so h + K + W and d + h + K + W are formed off the critical path.
// A single SHA-256 chain is pure latency (the blob hash is the critical path of a 6-blob request).
// FMA_ADDS: additions as IMAD.IADD (see sha_add_fma); used by the rounds-only half of the two-warp hash.
template <bool FMA_ADDS = false>
RK_HD void sha256_round(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d, uint32_t& e, uint32_t& f, uint32_t& g,
                        uint32_t& h, uint32_t kw, uint32_t one = 1u) {
    const uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
    const uint32_t ch = g ^ (e & (f ^ g));
    const uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
    const uint32_t mj = (a & b) | (c & (a | b));
    uint32_t e_new, a_new;
    if (FMA_ADDS) {
        const uint32_t hkw = sha_add_fma(h, kw, one);
        const uint32_t dhkw = sha_add_fma(d, hkw, one);
        const uint32_t sc = sha_add_fma(S1, ch, one);
        const uint32_t sm = sha_add_fma(S0, mj, one);
        e_new = sha_add_fma(dhkw, sc, one);
        a_new = sha_add_fma(sha_add_fma(hkw, sc, one), sm, one);
    } else {
        const uint32_t hkw = h + kw;
        const uint32_t dhkw = d + hkw;
        e_new = dhkw + S1 + ch;
        const uint32_t t1 = hkw + S1 + ch;
        a_new = t1 + S0 + mj;
    }
    h = g; g = f; f = e; e = e_new; d = c; c = b; b = a; a = a_new;
}

// One compression; w[16] = the block as big-endian words (destroyed).
RK_HD void sha256_compress(Sha256State& s, uint32_t (&w)[16]) {
    uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        // textbook statement here: with the message schedule interleaved, the one-warp batch
        // kernel is faster this way (3.3 ms per 4736 blobs against 4.0 with sha256_round)
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + SHA256_K::at(i) + w[i & 15];
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

// The two halves of a compression, for the latency-oriented two-warp hash (k_sha_blob_duo):
// message schedule (no dependence on the chaining state) and the 64 rounds.
RK_HD void sha256_schedule(uint32_t* wout /* [64] */, const uint32_t (&blk)[16]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { w[i] = blk[i]; wout[i] = blk[i]; }
#pragma unroll
    for (int i = 16; i < 64; i++) {
        uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
        uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
        uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
        w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        wout[i] = w[i & 15];
    }
}
// the same with the round constants added: kw[i] = K[i] + w[i], what sha256_rounds consumes
RK_HD void sha256_schedule_k(uint32_t* kwout /* [64] */, const uint32_t (&blk)[16]) {
    sha256_schedule(kwout, blk);
#pragma unroll
    for (int i = 0; i < 64; i++) kwout[i] += SHA256_K::at(i);
}
// kw[i] = K[i] + w[i] (the schedule half adds the round constants)
template <bool FMA_ADDS = false>
RK_HD void sha256_rounds(Sha256State& s, const uint32_t* kw /* [64], expanded, round constants added */, uint32_t one = 1u) {
    uint32_t a = s.h[0], b = s.h[1], c = s.h[2], d = s.h[3], e = s.h[4], f = s.h[5], g = s.h[6], h = s.h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        sha256_round<FMA_ADDS>(a, b, c, d, e, f, g, h, kw[i], one);
    }
    s.h[0] += a; s.h[1] += b; s.h[2] += c; s.h[3] += d; s.h[4] += e; s.h[5] += f; s.h[6] += g; s.h[7] += h;
}

RK_HD uint32_t load_be32(const uint8_t* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}
RK_HD void store_be32(uint8_t* p, uint32_t x) {
    p[0] = (uint8_t)(x >> 24); p[1] = (uint8_t)(x >> 16); p[2] = (uint8_t)(x >> 8); p[3] = (uint8_t)x;
}

// Hash a short message (len <= 119 bytes: at most two blocks).  Byte-wise: only used
// for the 48-byte commitment and the 64-byte (blob_hash || versioned_hash) input.
RK_HD void sha256_short(const uint8_t* msg, int len, uint8_t out[32]) {
    Sha256State s;
    sha256_init(s);
    uint8_t buf[128];
    for (int i = 0; i < 128; i++) buf[i] = 0;
    for (int i = 0; i < len; i++) buf[i] = msg[i];
    buf[len] = 0x80;
    int nblk = (len + 9 <= 64) ? 1 : 2;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) buf[nblk * 64 - 1 - i] = (uint8_t)(bits >> (8 * i));
    for (int b = 0; b < nblk; b++) {
        uint32_t w[16];
        for (int i = 0; i < 16; i++) w[i] = load_be32(buf + 64 * b + 4 * i);
        sha256_compress(s, w);
    }
    for (int i = 0; i < 8; i++) store_be32(out + 4 * i, s.h[i]);
}

}  // namespace rk
