// FP64 throughput on B200 (sm_100a): the roofline denominators for the wide-model Riccati kernel (SURVEY 8d asks the
// builder to measure them).  (1) DFMA: every thread runs 8 independent fma chains; (2) DMMA: mma.sync.m8n8k4.f64 with
// 4 independent accumulator fragments per warp; (3) DMMA m16n8k8 (if the assembler accepts it for this target).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu ; run: ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters, double a, double b) {
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void k_dmma884(double* out, int iters, double a, double b) {
    double c[8];
    for (int i = 0; i < 8; ++i) c[i] = threadIdx.x + i;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        dmma884(c[0], c[1], a, b); dmma884(c[2], c[3], a, b); dmma884(c[4], c[5], a, b); dmma884(c[6], c[7], a, b);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef WITH_M16N8K8
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__global__ void k_dmma1688(double* out, int iters, double av, double bv) {
    double c0[4], c1[4], c2[4], c3[4], a[4], b[2];
    for (int i = 0; i < 4; ++i) { c0[i] = threadIdx.x + i; c1[i] = c0[i] + 1; c2[i] = c0[i] + 2; c3[i] = c0[i] + 3; a[i] = av + i; }
    b[0] = bv; b[1] = bv + 1;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) { dmma1688(c0, a, b); dmma1688(c1, a, b); dmma1688(c2, a, b); dmma1688(c3, a, b); }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += c0[i] + c1[i] + c2[i] + c3[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

template <typename F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, max SM clock %d MHz\n", p.name, p.multiProcessorCount, clk / 1000);
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 64 * 1024);
    const int iters = 20000;
    const int sms = p.multiProcessorCount;
    printf("%-28s %10s %10s %12s\n", "kernel", "warps/SM", "ms", "TFLOP/s");
    for (int wps : {1, 2, 4, 8, 16, 32}) {
        const int threads = wps >= 8 ? 256 : 32 * wps, ctas = sms * (wps >= 8 ? wps / 8 : 1);
        double ms = time_ms([&] { k_dfma<<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("%-28s %10d %10.3f %12.2f\n", "DFMA (8 chains/thread)", wps, ms, 2.0 * 8 * iters * (double)ctas * threads / ms / 1e9);
    }
    for (int wps : {1, 2, 4, 8, 16, 32}) {
        const int threads = wps >= 8 ? 256 : 32 * wps, ctas = sms * (wps >= 8 ? wps / 8 : 1);
        double ms = time_ms([&] { k_dmma884<<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        // one m8n8k4 = 8*8*4 FMAs = 512 flops per warp instruction
        printf("%-28s %10d %10.3f %12.2f\n", "DMMA m8n8k4 (4 frags/warp)", wps, ms, 512.0 * 4 * iters * (double)ctas * (threads / 32) / ms / 1e9);
    }
#ifdef WITH_M16N8K8
    for (int wps : {1, 2, 4, 8, 16, 32}) {
        const int threads = wps >= 8 ? 256 : 32 * wps, ctas = sms * (wps >= 8 ? wps / 8 : 1);
        double ms = time_ms([&] { k_dmma1688<<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("%-28s %10d %10.3f %12.2f\n", "DMMA m16n8k8 (4 frags/warp)", wps, ms, 2048.0 * 4 * iters * (double)ctas * (threads / 32) / ms / 1e9);
    }
#endif
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
