// Do global stores cost DRAM *reads* on B200?  ncu of k_linback / k_forward shows ~0.75 sector read from DRAM per sector
// written (L2 write misses that fill the rest of the line).  This microbenchmark writes a large buffer with the store
// patterns the engine could use; run it under  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
// and compare bytes read per kernel.  Layout as in the engine: rows of 256 bytes (32 problems x 8 bytes) `pitch` doubles apart.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ROWS = 64; // rows written per warp

__global__ void k_st64_strided(double* out, size_t pitch) { // the engine's pattern: one 8-byte store per lane, rows far apart
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < ROWS; ++r) out[(size_t)r * pitch + b] = (double)r + b;
}
__global__ void k_st64_cs(double* out, size_t pitch) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < ROWS; ++r) __stcs(&out[(size_t)r * pitch + b], (double)r + b);
}
__global__ void k_st64_wt(double* out, size_t pitch) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < ROWS; ++r) __stwt(&out[(size_t)r * pitch + b], (double)r + b);
}
__global__ void k_st64_cg(double* out, size_t pitch) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < ROWS; ++r) __stcg(&out[(size_t)r * pitch + b], (double)r + b);
}
__global__ void k_st128_strided(double* out, size_t pitch) { // 16 bytes per lane: a warp covers 512 contiguous bytes (two rows' worth)
    const size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    for (int r = 0; r < ROWS / 2; ++r) {
        double2 v = make_double2((double)r + b, (double)r - b);
        *reinterpret_cast<double2*>(&out[(size_t)r * 2 * pitch + b]) = v;
    }
}
__global__ void k_st64_dense(double* out, size_t pitch) { // same bytes, fully contiguous (row r of warp w right after row r-1)
    const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x & 31;
    for (int r = 0; r < ROWS; ++r) out[(w * ROWS + r) * 32 + lane] = (double)r + lane;
}
__global__ void k_bulk_store(double* out, size_t pitch) { // rows staged in shared memory, written by the TMA engine (cp.async.bulk.global.shared)
    __shared__ __align__(128) double stage[2][ROWS][32]; // 2 warps per block (launched with 64 threads)
    const int wid = threadIdx.x / 32, lane = threadIdx.x & 31;
    const size_t b0 = (size_t)blockIdx.x * blockDim.x + wid * 32;
    for (int r = 0; r < ROWS; ++r) stage[wid][r][lane] = (double)r + b0 + lane;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        for (int r = 0; r < ROWS; ++r) {
            const unsigned sa = (unsigned)__cvta_generic_to_shared(&stage[wid][r][0]);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;" ::"l"(out + (size_t)r * pitch + b0), "r"(sa) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
__global__ void k_st64_rmw(double* out, size_t pitch) { // read the row first (accumulator pattern), then write it
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int r = 0; r < ROWS; ++r) { double v = out[(size_t)r * pitch + b]; out[(size_t)r * pitch + b] = v + 1.0; }
}

int main() {
    const size_t pitch = (size_t)148 * 8 * 32 * 64; // 2.4 M columns -> rows 19 MB apart; one launch writes ROWS * pitch * 8 = 1.24 GB (>> the 126 MB L2)
    const size_t n = (size_t)ROWS * pitch;
    double* buf; cudaMalloc(&buf, n * 8);
    cudaMemset(buf, 0, n * 8);
    const int threads = 128; const int blocks = (int)(pitch / threads);
    // a big L2-flushing memset between kernels so that every launch starts with its lines not resident
    double* flush; cudaMalloc(&flush, (size_t)512 << 20);
    auto fl = [&] { cudaMemset(flush, 1, (size_t)512 << 20); cudaDeviceSynchronize(); };
    fl(); k_st64_strided<<<blocks, threads>>>(buf, pitch); cudaDeviceSynchronize();
    fl(); k_st64_cs<<<blocks, threads>>>(buf, pitch); cudaDeviceSynchronize();
    fl(); k_st64_wt<<<blocks, threads>>>(buf, pitch); cudaDeviceSynchronize();
    fl(); k_st64_cg<<<blocks, threads>>>(buf, pitch); cudaDeviceSynchronize();
    fl(); k_st128_strided<<<blocks / 2, threads>>>(buf, pitch); cudaDeviceSynchronize();
    fl(); k_st64_dense<<<blocks, threads>>>(buf, pitch); cudaDeviceSynchronize();
    fl(); k_bulk_store<<<blocks * 2, 64>>>(buf, pitch); cudaDeviceSynchronize();
    fl(); k_st64_rmw<<<blocks, threads>>>(buf, pitch); cudaDeviceSynchronize();
    printf("bytes written per kernel: %.1f MB; status %s\n", n * 8 / 1e6, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
