// Streaming-read probe: which access pattern / launch shape reaches HBM speed on B200?
// Reads a fp64 array (8 B/entry) and a u16 array (2 B/entry) the way KR's SpMV streams its operand.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_probe stream_probe.cu && ./stream_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Regs { double a[8]; unsigned c[4]; };

__device__ __forceinline__ void load256(const double *v, const uint16_t *c, int64_t e, Regs &R) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(R.a[0]), "=d"(R.a[1]), "=d"(R.a[2]), "=d"(R.a[3]) : "l"(v + e));
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(R.a[4]), "=d"(R.a[5]), "=d"(R.a[6]), "=d"(R.a[7]) : "l"(v + e + 4));
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(R.c[0]), "=r"(R.c[1]), "=r"(R.c[2]), "=r"(R.c[3]) : "l"(c + e));
}
__device__ __forceinline__ double consume(const Regs &R) {
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += R.a[i];
    return s + (double)(R.c[0] ^ R.c[1] ^ R.c[2] ^ R.c[3]);
}

// mode 0: every warp owns a contiguous run of 256-entry chunks; mode 1: chunks are dealt round-robin over all warps
template <int D>
__global__ void __launch_bounds__(1024, 1) k_warp_stream(const double *v, const uint16_t *c, int64_t n_chunks, int mode, double *out) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    int64_t c0, c1, step;
    if (mode == 0) { c0 = n_chunks * gw / nw; c1 = n_chunks * (gw + 1) / nw; step = 1; }
    else { c0 = gw; c1 = n_chunks; step = nw; }
    Regs R[D];
#pragma unroll
    for (int d = 0; d < D; ++d) if (c0 + d * step < c1) load256(v, c, (c0 + d * step) * 256 + 8 * lane, R[d]);
    double s = 0;
    for (int64_t k = c0; k < c1; k += D * step) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            if (k + d * step < c1) {
                s += consume(R[d]);
                if (k + (d + D) * step < c1) load256(v, c, (k + (d + D) * step) * 256 + 8 * lane, R[d]);
            }
        }
    }
    if (s == 1.2345) out[0] = s;
}

// plain grid-stride read, 16 B per thread per step
__global__ void k_plain(const double2 *v, int64_t n2, double *out) {
    double s = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        double2 x = __ldg(v + i);
        s += x.x + x.y;
    }
    if (s == 1.2345) out[0] = s;
}

template <typename F>
static float time_it(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main() {
    const int64_t n = (int64_t)48 << 20;                 // 48 Mi entries: 384 MB + 96 MB, well beyond L2
    const int64_t n_chunks = n / 256;
    double *v, *out; uint16_t *c;
    CK(cudaMalloc(&v, n * 8)); CK(cudaMalloc(&c, n * 2)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(v, 0, n * 8)); CK(cudaMemset(c, 0, n * 2));
    const double gb = n * 10 / 1e9;
    for (int mode = 0; mode < 2; ++mode) {
        for (int threads : {512, 1024}) {
            for (int grid : {148, 296}) {
                if (grid * threads > 148 * 2048) continue;
                float ms;
                ms = time_it([&] { k_warp_stream<2><<<grid, threads>>>(v, c, n_chunks, mode, out); }, 5);
                printf("warp_stream mode=%d D=2 grid=%d threads=%d  %.1f us  %.0f GB/s\n", mode, grid, threads, ms * 1e3, gb / ms * 1e3);
                ms = time_it([&] { k_warp_stream<3><<<grid, threads>>>(v, c, n_chunks, mode, out); }, 5);
                printf("warp_stream mode=%d D=3 grid=%d threads=%d  %.1f us  %.0f GB/s\n", mode, grid, threads, ms * 1e3, gb / ms * 1e3);
                ms = time_it([&] { k_warp_stream<4><<<grid, threads>>>(v, c, n_chunks, mode, out); }, 5);
                printf("warp_stream mode=%d D=4 grid=%d threads=%d  %.1f us  %.0f GB/s\n", mode, grid, threads, ms * 1e3, gb / ms * 1e3);
            }
        }
    }
    for (int grid : {148, 296, 592, 1184, 4736}) {
        for (int threads : {256, 512, 1024}) {
            float ms = time_it([&] { k_plain<<<grid, threads>>>((const double2 *)v, n / 2, out); }, 5);
            printf("plain grid=%d threads=%d  %.1f us  %.0f GB/s\n", grid, threads, ms * 1e3, n * 8 / 1e9 / ms * 1e3);
        }
    }
    CK(cudaGetLastError());
    return 0;
}
