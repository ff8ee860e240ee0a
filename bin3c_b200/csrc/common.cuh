// Shared helpers for the bin3c_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <functional>
#include <mutex>
#include <unordered_map>

#include "../../include/bin3c_b200.h"

namespace b3c {

constexpr int kNumSMs = 148;            // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
// cross-GPU flag waits give up after this many SM cycles (b3c_set_option(B3C_OPT_PEER_TIMEOUT_MS); default 60 s)
extern std::atomic<long long> g_peer_timeout_cycles;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define B3C_CUDA(call)                                                                    \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            b3c::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return B3C_ERR_CUDA;                                                          \
        }                                                                                 \
    } while (0)

#define B3C_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        b3c::count_launch();                                                              \
        B3C_CUDA(cudaGetLastError());                                                     \
    } while (0)

#define B3C_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            b3c::set_error(__VA_ARGS__);                                                  \
            return B3C_ERR_ARG;                                                           \
        }                                                                                 \
    } while (0)

inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- CUDA-graph cache ------------------------------------------------------------------------------
// The sort-reduce of an accumulator is ~60 short launches whose grids and arguments are fixed per workspace (all
// sizes are read from device counters), so the sequence is captured once per (workspace, sequence id, aux pointer)
// and replayed with ONE cudaGraphLaunch: the inter-launch gaps (a few microseconds each) are what dominates these
// sequences at multi-GPU shard sizes.  `enqueue` must only enqueue work on `s` (kernels, device-to-device copies,
// memsets): no host-pointer copies, no synchronisation, no attribute calls.  b3c_set_option(B3C_OPT_USE_GRAPHS, 0)
// turns replay off (every call enqueues directly).
int graph_run(cudaStream_t s, const void *ws, int id, const void *aux, const std::function<int()> &enqueue);
void graph_forget(const void *ws);            // the workspace was re-planned: drop its graphs
extern std::atomic<int> g_use_graphs;

// carve a workspace: returns the offset of a block of `bytes` and advances the cursor
struct Carver {
    int64_t cur = 0;
    int64_t take(int64_t bytes) {
        int64_t o = cur;
        cur = align_up(cur + bytes, 256);
        return o;
    }
};

// ---- device helpers -----------------------------------------------------------------

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming 128-bit load: read once, do not keep in L1
__device__ __forceinline__ uint4 ld_stream_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream_s32(const int *p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
__device__ __forceinline__ unsigned warp_max_u32(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

// count / (s_i * s_j) with the reference's operation order (contact_map.py:110-113: t = s_i*s_j; r = 1.0/t;
// d = d*r) and zero site counts taken as one (contact_map.py:1103-1108, Q6): the value k_site_norm writes.
__device__ __forceinline__ double site_scaled(uint32_t count, int32_t s_row, int32_t s_col) {
    const double si = s_row == 0 ? 1.0 : (double)s_row;
    const double sj = s_col == 0 ? 1.0 : (double)s_col;
    return __dmul_rn((double)count, __ddiv_rn(1.0, __dmul_rn(si, sj)));
}

// inclusive warp scan
template <typename T>
__device__ __forceinline__ T warp_scan_incl(T v) {
    const unsigned l = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(kFullMask, v, o);
        if (l >= (unsigned)o) v += t;
    }
    return v;
}

// Exclusive block scan of one value per thread (blockDim.x <= 1024, multiple of 32).
// `s_warp` must hold 33 elements.  Returns the exclusive prefix; *total gets the block sum.
template <typename T>
__device__ __forceinline__ T block_scan_excl(T v, T *s_warp, T *total) {
    const unsigned l = lane_id(), w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    T incl = warp_scan_incl(v);
    if (l == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
        T x = (l < nw) ? s_warp[l] : T(0);
        T xi = warp_scan_incl(x);
        s_warp[l] = xi - x;
        if (l == 31) s_warp[32] = xi;
    }
    __syncthreads();
    T r = s_warp[w] + incl - v;
    *total = s_warp[32];
    __syncthreads();
    return r;
}

// ---- generic device-wide exclusive scan over int64 (three small kernels) --------------
// out[i] = sum_{k<i} in[k] for i in [0, n]; out has n+1 elements.  `d_tmp` needs
// scan_tmp_elems(n) int64 elements.
int64_t scan_tmp_elems(int64_t n);
int scan_exclusive_i64(const int64_t *d_in, int64_t *d_out, int64_t n, int64_t *d_tmp, cudaStream_t s);

}  // namespace b3c
