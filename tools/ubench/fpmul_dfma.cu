// Gate for an FP64-pipe formulation of the 381-bit Montgomery product ("next" #4).
//
// Question: the integer formulation (field30.cuh: 13 x 30-bit limbs, IMAD.WIDE) sustains 25 G Fp
// products/s register-resident on a B200 (profiles/r01/fpmul_ubench_r01.txt) while the FP64 pipe
// (63 DFMA/clk/SM, profiles/r01/dfma_rates_r01.txt) idles.  Would 8 x 52-bit limbs on DFMA win?
//
// Method (Emmart, Zheng, Weems, "Faster modular exponentiation using double precision floating point
// arithmetic on the GPU", ARITH 2018): for integers a, b < 2^52 held exactly in doubles,
//     hi  = fma_rz(a, b, 2^104)            mantissa = (a b) >> 52
//     lo  = fma_rz(a, b, (2^104 + 2^52) - hi)   mantissa = (a b) mod 2^52
// and the raw 64-bit patterns are accumulated per column with INTEGER adds (the exponent fields are
// constants that are subtracted once per column).  Per partial product: 2 DFMA + 1 DADD on the FP64
// pipe and two 64-bit integer additions (4 IADD3-class instructions) on the ALU pipe.
//
// This file measures the PRODUCT ROWS ONLY (64 partial products -> 16 columns, then carry
// normalisation and the conversion of the low half back to doubles so that products can be chained).
// A full Montgomery product needs a second set of 64 partial products against the modulus plus eight
// 52-bit quotient digits, i.e. more than twice this work.  Gate: adopt only if the full product can
// beat 25 G/s by >= 1.4x = 35 G/s, i.e. only if THIS kernel sustains >= ~74 G product-rows/s.
// The kernel is self-checking: products of random operands are compared with unsigned __int128
// arithmetic on the host before anything is timed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fpmul_dfma fpmul_dfma.cu && ./fpmul_dfma
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int L = 8;                       // 8 x 52 = 416 bits >= 381
constexpr uint64_t M52 = (1ull << 52) - 1;

__device__ __forceinline__ double u52_to_double(uint64_t v) {          // v < 2^52, exact
    return __longlong_as_double((long long)(v | 0x4330000000000000ull)) - 4503599627370496.0;   // (2^52 + v) - 2^52
}

// cols[0..15] = sum of the partial products of a and b by column, as plain integers
// (column k holds lo parts of i + j = k and hi parts of i + j = k - 1; each part < 2^52, at most 16 of them)
__device__ __forceinline__ void product_rows(uint64_t (&cols)[2 * L], const double (&a)[L], const double (&b)[L]) {
    const double C1 = 20282409603651670423947251286016.0;               // 2^104
    const double C2 = 20282409603651674927546878656512.0;               // 2^104 + 2^52
    const uint64_t B1 = 0x4670000000000000ull;                          // bits of 2^104
    const uint64_t B2 = 0x4330000000000000ull;                          // bits of 2^52
    uint64_t acc[2 * L];
#pragma unroll
    for (int k = 0; k < 2 * L; k++) acc[k] = 0;
#pragma unroll
    for (int i = 0; i < L; i++) {
#pragma unroll
        for (int j = 0; j < L; j++) {
            const double hi = __fma_rz(a[i], b[j], C1);
            const double lo = __fma_rz(a[i], b[j], C2 - hi);
            acc[i + j + 1] += (uint64_t)__double_as_longlong(hi);
            acc[i + j] += (uint64_t)__double_as_longlong(lo);
        }
    }
    // remove the exponent patterns: column k received (number of lo parts) x B2 + (number of hi parts) x B1
#pragma unroll
    for (int k = 0; k < 2 * L; k++) {
        const int nlo = k < L ? k + 1 : 2 * L - 1 - k;                  // pairs with i + j = k
        const int nhi = (k >= 1) ? ((k - 1) < L ? k : 2 * L - k) : 0;   // pairs with i + j = k - 1
        cols[k] = acc[k] - (uint64_t)nlo * B2 - (uint64_t)nhi * B1;
    }
}

// one chained step: x <- low 8 limbs of (x * b), normalised to 52 bits and converted back to doubles
__device__ __forceinline__ void chain_step(double (&x)[L], const double (&b)[L]) {
    uint64_t cols[2 * L];
    product_rows(cols, x, b);
    uint64_t carry = 0, top = 0;
#pragma unroll
    for (int k = 0; k < 2 * L; k++) {
        const uint64_t t = cols[k] + carry;
        carry = t >> 52;
        if (k < L) x[k] = u52_to_double(t & M52);
        else top ^= t;                                                  // keep the high half alive
    }
    x[0] = u52_to_double((__double_as_longlong(x[0]) ^ top) & M52 & ~1ull | 1ull);   // fold it in (never zero)
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_chain(const uint64_t* in, uint64_t* out, int iters) {
    double x[L], b[L];
    const int t = blockIdx.x * THREADS + threadIdx.x;
#pragma unroll
    for (int i = 0; i < L; i++) { x[i] = u52_to_double((in[i] + 977ull * t) & M52); b[i] = u52_to_double((in[L + i] ^ (131ull * t)) & M52); }
    for (int k = 0; k < iters; k++) chain_step(x, b);
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < L; i++) r ^= (uint64_t)__double_as_longlong(x[i]);
    out[t] = r;
}

// correctness: full 16-column product of given operands
__global__ void k_check(const uint64_t* a, const uint64_t* b, uint64_t* cols_out, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    double x[L], y[L];
#pragma unroll
    for (int i = 0; i < L; i++) { x[i] = u52_to_double(a[L * t + i]); y[i] = u52_to_double(b[L * t + i]); }
    uint64_t cols[2 * L];
    product_rows(cols, x, y);
#pragma unroll
    for (int k = 0; k < 2 * L; k++) cols_out[2 * L * t + k] = cols[k];
}

template <typename F> static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    // ---- self-check against unsigned __int128 ----
    const int n = 4096;
    std::vector<uint64_t> ha(L * n), hb(L * n), hc(2 * L * n);
    uint64_t s = 0x243F6A8885A308D3ull;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (s >> 11) & M52; };
    for (auto& v : ha) v = rnd();
    for (auto& v : hb) v = rnd();
    for (int i = 0; i < L; i++) { ha[i] = M52; hb[i] = M52; }                       // worst case first
    uint64_t *da, *db, *dc;
    cudaMalloc(&da, 8 * ha.size()); cudaMalloc(&db, 8 * hb.size()); cudaMalloc(&dc, 8 * hc.size());
    cudaMemcpy(da, ha.data(), 8 * ha.size(), cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), 8 * hb.size(), cudaMemcpyHostToDevice);
    k_check<<<(n + 127) / 128, 128>>>(da, db, dc, n);
    cudaMemcpy(hc.data(), dc, 8 * hc.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < n && bad < 5; t++) {
        unsigned __int128 want[2 * L] = {0};
        for (int i = 0; i < L; i++) for (int j = 0; j < L; j++) {
            unsigned __int128 pr = (unsigned __int128)ha[L * t + i] * hb[L * t + j];
            want[i + j] += (uint64_t)(pr & M52);
            want[i + j + 1] += (uint64_t)(pr >> 52);
        }
        for (int k = 0; k < 2 * L; k++) if ((uint64_t)want[k] != hc[2 * L * t + k]) { bad++; printf("MISMATCH t=%d col=%d\n", t, k); break; }
    }
    printf("self-check: %d products of 8 x 52-bit limbs vs unsigned __int128: %s\n", n, bad ? "FAILED" : "ok");
    if (bad) return 1;
    // ---- rate ----
    uint64_t* out; cudaMalloc(&out, 8ull * sms * 2048);
    const int it = 2000;
    const double clk = 1.965e9;
    printf("device %s, %d SMs; product rows only (64 partial products + normalisation + re-conversion per op)\n", p.name, sms);
#define RUN(T, MB) { const int grid = sms * MB; float ms = time_ms([&] { k_chain<T, MB><<<grid, T>>>(da, out, it); });                       \
      const double ops = (double)grid * T * it; const double rate = ops / (ms * 1e-3);                                                      \
      printf("threads/SM %4d : %8.3f ms  %7.2f G product-rows/s  = %5.2f SM-clocks per op; FP64 pipe (192 ops) %4.1f%% of 63/clk/SM\n",      \
             T * MB, ms, rate / 1e9, sms * clk / rate, 100.0 * rate * 192 / (63.0 * sms * clk)); }
    RUN(128, 1) RUN(256, 1) RUN(256, 2) RUN(256, 3) RUN(256, 4)
    printf("gate: a full Montgomery product needs > 2x this work; adopting FP64 limbs needs >= 35 G full products/s, i.e. >= ~74 G product-rows/s here "
           "(integer formulation today: 25 G full products/s)\n");
    return 0;
}
