// Sustained throughput of the product's own Fp multiply / square / mixed add, register
// resident, at several occupancies.  Tells how much of the integer-multiply peak the
// instruction mix can reach when memory is out of the picture.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../raiko_b200/csrc/g1.cuh"
using namespace rk;

template <int MODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_fp(uint32_t* io, int iters) {
    Fp a, b;
    for (int i = 0; i < 13; i++) { a.v[i] = (io[i] + threadIdx.x * 977u) & LIMB_MASK; b.v[i] = (io[13 + i] + threadIdx.x * 131u) & LIMB_MASK; }
    a.v[12] &= 0xfffff; b.v[12] &= 0xfffff;
    if (MODE == 0) for (int k = 0; k < iters; k++) { fe_mul(a, a, b); }
    if (MODE == 1) for (int k = 0; k < iters; k++) { fe_sqr(a, a); }
    if (MODE == 2) {   // two independent chains
        Fp c = b;
        for (int k = 0; k < iters; k++) { fe_mul(a, a, b); fe_mul(c, c, b); }
        fe_add(a, a, c);
    }
    if (MODE == 3) {   // madd chain
        G1Xyzz acc; acc.x = a; acc.y = b; fe_const<FpTag, FP_ONE>(acc.zz); fe_const<FpTag, FP_ONE>(acc.zzz);
        for (int k = 0; k < iters; k++) { g1_madd(acc, a, b); }
        fe_add(a, acc.x, acc.zz);
    }
    uint32_t r = 0;
    for (int i = 0; i < 13; i++) r ^= a.v[i];
    io[64 + blockIdx.x * THREADS + threadIdx.x] = r;
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
    int sms = p.multiProcessorCount;
    uint32_t* io; cudaMalloc(&io, 4 * (64 + 1024 * sms * 4)); cudaMemset(io, 0x5a, 4 * 64);
    const double peak = 64.0 * sms * 1.965e9;
    const int it = 2000;
#define RUN(MODE, T, MB, name, unitops, pipe)                                                           \
    { int grid = sms * MB; float ms = time_ms([&] { k_fp<MODE, T, MB><<<grid, T>>>(io, it); });         \
      double ops = (double)grid * T * it * unitops; double rate = ops / (ms * 1e-3);                     \
      printf("%-22s threads/SM %4d : %8.3f ms  %7.2f G ops/s   multiply-pipe %5.1f%% of nominal\n", name, T * MB, ms, rate / 1e9, 100 * rate * pipe / peak); }
    RUN(0, 256, 1, "mul chain", 1, 353) RUN(0, 256, 2, "mul chain", 1, 353) RUN(0, 256, 4, "mul chain", 1, 353) RUN(0, 256, 6, "mul chain", 1, 353)
    RUN(1, 256, 1, "sqr chain", 1, 275) RUN(1, 256, 2, "sqr chain", 1, 275) RUN(1, 256, 4, "sqr chain", 1, 275) RUN(1, 256, 6, "sqr chain", 1, 275)
    RUN(2, 256, 1, "2 mul chains", 2, 353) RUN(2, 256, 2, "2 mul chains", 2, 353) RUN(2, 256, 4, "2 mul chains", 2, 353)
    RUN(3, 256, 1, "madd chain", 1, 3374) RUN(3, 256, 2, "madd chain", 1, 3374) RUN(3, 384, 1, "madd chain", 1, 3374)
    return 0;
}
