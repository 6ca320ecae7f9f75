// IMAD.WIDE.U32 / IMAD rate by operand form, every instruction stated as volatile inline PTX so that
// nothing is hoisted, merged or strength-reduced (wide_operands2.cu's C statements let the compiler fold
// the repeated immediate products).  Question: is IMAD.WIDE Rd, Ra, imm, Rc64 -- the Montgomery REDUCTION
// row of field30.cuh -- full rate like the RZ-addend form, or half rate like Ra, Rb, Rc64?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wide_operands3 wide_operands3.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define NCH 12
__device__ __forceinline__ uint32_t launder(uint32_t x) { asm("" : "+r"(x)); return x; }
__device__ __constant__ uint32_t dummy;
template <int MODE>
__global__ void __launch_bounds__(256) k(uint64_t* out, uint32_t seed, int iters) {
    // C statements (ptxas splits an explicit mad.wide.u32 into IMAD.WIDE ..., RZ + 64-bit adds); every product of
    // one pass over (u, i) is a distinct (multiplicand, multiplier) pair and the multiplicands change between
    // passes, so nothing can be hoisted or merged.  The SASS forms are checked with cuobjdump (see the .txt).
    constexpr uint32_t IMM[8] = {0x2affffacu, 0x27fbffffu, 0x3fffaaabu, 0x153ffffbu, 0x0f6241ebu, 0x1cc1a0f6u, 0x04bf6730u, 0x2d3a1270u};
    uint64_t w[NCH], v[NCH]; uint32_t a[NCH], b[NCH], c[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) { w[i] = threadIdx.x + i; v[i] = 0; a[i] = seed * (i + 3) + threadIdx.x; b[i] = seed * (i + 7) ^ threadIdx.x; c[i] = seed + i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int i = 0; i < NCH; i++) {
                const int j = (i + u) % NCH;
                if (MODE == 0) w[i] += (uint64_t)a[j] * b[(i + 5 * u + 1) % NCH];        // Ra, Rb, Rc64: all distinct
                if (MODE == 1) w[i] += (uint64_t)a[j] * IMM[u];                          // Ra, imm, Rc64: the Montgomery reduction row
                if (MODE == 2) v[i] ^= (uint64_t)a[j] * b[(i + 5 * u + 1) % NCH];        // Ra, Rb, RZ (+ 2 LOP3)
                if (MODE == 3) c[i] += a[j] * b[(i + 5 * u + 1) % NCH];                  // IMAD Ra, Rb, Rc (32-bit)
                if (MODE == 4) c[i] += a[j] * IMM[u];                                    // IMAD Ra, imm, Rc (32-bit)
                if (MODE == 5) w[i] += (uint64_t)a[u] * b[i];                            // operand scanning: Ra fixed over a row of 12
            }
        }
#pragma unroll
        for (int i = 0; i < NCH; i++) { a[i] = launder(a[i] ^ (uint32_t)w[i] ^ c[i]); b[i] = launder(b[i] + (uint32_t)(w[(i + 1) % NCH] >> 32) + (uint32_t)v[i]); }
    }
    uint64_t r = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) r ^= w[i] ^ v[i] ^ a[i] ^ c[i];
    out[blockIdx.x * 256 + threadIdx.x] = r;
}
template <typename F> static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    uint64_t* out; cudaMalloc(&out, 8 * 256 * sms * 8);
    const int iters = 2048;
    const char* names[6] = {"WIDE Ra,Rb,Rc64 distinct", "WIDE Ra,imm,Rc64", "WIDE Ra,Rb,RZ (+2 LOP3)", "IMAD Ra,Rb,Rc", "IMAD Ra,imm,Rc", "WIDE Ra(row),Rb,Rc64"};
#define RUN(M) { float ms = time_ms([&] { k<M><<<sms * 4, 256>>>(out, 12345, iters); }); double ops = (double)sms * 4 * 256 * iters * 8 * NCH; \
    printf("%-26s %8.3f ms  %6.2f T/s  (%.1f per clk per SM at 1.965 GHz)\n", names[M], ms, ops / ms / 1e9, ops / (ms * 1e-3) / sms / 1.965e9); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5)
    return 0;
}
