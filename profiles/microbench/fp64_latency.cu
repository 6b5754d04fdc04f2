// FP64 dependent-chain latencies and single-warp issue rates on B200 (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I include -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ilqr_model_rt.h"

#define REP 2048
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b) {
    double x = a + threadIdx.x * 1e-9, y = b, z = a * 0.5, w = b * 0.25;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < REP; ++i) {
        if (OP == 0) x = __fma_rn(x, y, y);            // dependent DFMA
        if (OP == 1) x = __dadd_rn(x, y);              // dependent DADD
        if (OP == 2) x = __dmul_rn(x, y);              // dependent DMUL
        if (OP == 3) x = y / x;                        // dependent division
        if (OP == 4) x = sqrt(x) + y;                  // dependent sqrt (+add)
        if (OP == 5) { double s, c; ilqr_sincos(x, &s, &c); x = s + c; }  // sincos chain
        if (OP == 6) { x = __fma_rn(x, y, y); z = __fma_rn(z, y, y); }    // 2 independent chains
        if (OP == 7) { x = __fma_rn(x, y, y); z = __fma_rn(z, y, y); w = __fma_rn(w, y, y); a = __fma_rn(a, y, y); } // 4 chains
        if (OP == 8) x = 1.0 / x + y;                  // rcp (+add)
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + z + w + a;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name, int ops_per_iter) {
    double* out; long long* cyc; cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 8);
    chain<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    chain<OP><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %8.2f cycles / iteration (%d op(s) per iteration)\n", name, (double)h / REP, ops_per_iter);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("DFMA dependent", 1); run<1>("DADD dependent", 1); run<2>("DMUL dependent", 1);
    run<3>("div dependent", 1); run<4>("sqrt+add dependent", 2); run<5>("ilqr_sincos + add", 1);
    run<6>("DFMA 2 independent chains", 2); run<7>("DFMA 4 independent chains", 4); run<8>("rcp+add dependent", 2);
    return 0;
}
