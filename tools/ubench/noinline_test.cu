// Does an out-of-line (real CALL) field multiplication keep its operands in registers, and
// what does the call cost?  The inlined mixed addition is ~85 KB of code (10 products); with
// calls the loop body would fit the instruction cache.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../raiko_b200/csrc/g1.cuh"
using namespace rk;
__device__ __noinline__ Fp mul_call(Fp a, Fp b) { Fp r; fe_mul(r, a, b); return r; }
__device__ __noinline__ Fp sqr_call(Fp a) { Fp r; fe_sqr(r, a); return r; }
__device__ __forceinline__ void madd_calls(G1Xyzz& acc, const Fp& x2, const Fp& y2) {
    Fp P, Rr, PP, PPP, Q, t;
    P = mul_call(x2, acc.zz);
    Rr = mul_call(y2, acc.zzz);
    fe_sub<FpTag, 6>(P, P, acc.x);
    fe_sub<FpTag, 6>(Rr, Rr, acc.y);
    PP = sqr_call(P);
    PPP = mul_call(P, PP);
    Q = mul_call(acc.x, PP);
    acc.x = sqr_call(Rr);
    fe_add(t, Q, Q); fe_add(t, t, PPP);
    fe_sub<FpTag, 4>(acc.x, acc.x, t);
    fe_sub<FpTag, 6>(t, Q, acc.x);
    t = mul_call(Rr, t);
    Q = mul_call(acc.y, PPP);
    fe_sub<FpTag, 2>(acc.y, t, Q);
    acc.zz = mul_call(acc.zz, PP);
    acc.zzz = mul_call(acc.zzz, PPP);
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(uint32_t* io, int iters) {
    Fp a, b;
    for (int i = 0; i < 13; i++) { a.v[i] = (io[i] + threadIdx.x * 977u) & LIMB_MASK; b.v[i] = (io[13 + i] + threadIdx.x * 131u) & LIMB_MASK; }
    a.v[12] &= 0xfffff; b.v[12] &= 0xfffff;
    G1Xyzz acc; acc.x = a; acc.y = b; fe_const<FpTag, FP_ONE>(acc.zz); fe_const<FpTag, FP_ONE>(acc.zzz);
    for (int k = 0; k < iters; k++) { if (MODE == 0) g1_madd(acc, a, b); else madd_calls(acc, a, b); }
    fe_add(a, acc.x, acc.zz);
    uint32_t r = 0;
    for (int i = 0; i < 13; i++) r ^= a.v[i];
    io[64 + blockIdx.x * 256 + threadIdx.x] = r;
}
template <typename F> static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    uint32_t* io; cudaMalloc(&io, 4 * (64 + 256 * sms)); cudaMemset(io, 0x5a, 256);
    int it = 2000;
    float m0 = time_ms([&] { k<0><<<sms, 256>>>(io, it); });
    float m1 = time_ms([&] { k<1><<<sms, 256>>>(io, it); });
    printf("inline madd : %.3f ms  %.2f G madd/s\ncalls madd  : %.3f ms  %.2f G madd/s\n", m0, (double)sms * 256 * it / m0 / 1e6, m1, (double)sms * 256 * it / m1 / 1e6);
    return 0;
}
