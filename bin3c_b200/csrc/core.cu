// Library plumbing: error reporting, launch accounting, generic exclusive scan.
#include <stdarg.h>

#include <map>

#include "common.cuh"

namespace b3c {

static thread_local char t_err[1024] = "";
std::atomic<int64_t> g_launches{0};
std::atomic<long long> g_peer_timeout_cycles{60000LL * 2000000LL};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

// ---- CUDA-graph cache (see common.cuh) ----------------------------------------------------
std::atomic<int> g_use_graphs{1};
namespace {
struct GraphKey {
    const void *ws, *aux;
    int id;
    cudaStream_t s;
    bool operator<(const GraphKey &o) const {
        if (ws != o.ws) return ws < o.ws;
        if (aux != o.aux) return aux < o.aux;
        if (id != o.id) return id < o.id;
        return s < o.s;
    }
};
struct GraphEntry {
    cudaGraphExec_t exec;
    int64_t launches;
};
std::mutex g_graph_mu;
std::map<GraphKey, GraphEntry> g_graphs;
}  // namespace

int graph_run(cudaStream_t s, const void *ws, int id, const void *aux, const std::function<int()> &enqueue) {
    if (!g_use_graphs.load()) return enqueue();
    const GraphKey key{ws, aux, id, s};
    {
        std::lock_guard<std::mutex> g(g_graph_mu);
        auto it = g_graphs.find(key);
        if (it != g_graphs.end()) {
            B3C_CUDA(cudaGraphLaunch(it->second.exec, s));
            count_launch((int)it->second.launches);
            return B3C_OK;
        }
    }
    const int64_t before = g_launches.load();
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        (void)cudaGetLastError();
        return enqueue();                              // the stream cannot be captured (already capturing?): run directly
    }
    const int rc = enqueue();
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s, &graph);
    if (rc != B3C_OK || e != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        if (rc != B3C_OK) return rc;
        set_error("CUDA graph capture failed: %s", cudaGetErrorString(e));
        return B3C_ERR_CUDA;
    }
    GraphEntry ent;
    ent.launches = g_launches.load() - before;         // counted while capturing; stands for this first replay
    const cudaError_t ei = cudaGraphInstantiate(&ent.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
        set_error("cudaGraphInstantiate: %s", cudaGetErrorString(ei));
        return B3C_ERR_CUDA;
    }
    B3C_CUDA(cudaGraphLaunch(ent.exec, s));
    std::lock_guard<std::mutex> g(g_graph_mu);
    g_graphs[key] = ent;
    return B3C_OK;
}

void graph_forget(const void *ws) {
    std::lock_guard<std::mutex> g(g_graph_mu);
    for (auto it = g_graphs.begin(); it != g_graphs.end();) {
        if (it->first.ws == ws) {
            cudaGraphExecDestroy(it->second.exec);
            it = g_graphs.erase(it);
        } else {
            ++it;
        }
    }
}

// ---- exclusive scan ---------------------------------------------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

int64_t scan_tmp_elems(int64_t n) { return ceil_div(n > 0 ? n : 1, SCAN_TILE) + 1; }

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(const int64_t *__restrict__ in, int64_t n,
                                                                  int64_t *__restrict__ sums) {
    __shared__ int64_t s_w[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) v += in[base + k];
    int64_t tot;
    block_scan_excl<int64_t>(v, s_w, &tot);
    if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

// single block: exclusive scan of sums[0..nb) in place; sums[nb] = grand total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(int64_t *__restrict__ sums, int64_t nb) {
    __shared__ int64_t s_w[33];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        int64_t v = (i < nb) ? sums[i] : 0;
        int64_t tot;
        int64_t ex = block_scan_excl<int64_t>(v, s_w, &tot);
        const int64_t carry = s_carry;
        if (i < nb) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = s_carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const int64_t *__restrict__ in, int64_t n,
                                                             const int64_t *__restrict__ sums,
                                                             int64_t *__restrict__ out) {
    __shared__ int64_t s_w[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t x[SCAN_ITEMS];
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        x[k] = (base + k < n) ? in[base + k] : 0;
        v += x[k];
    }
    int64_t tot;
    int64_t ex = block_scan_excl<int64_t>(v, s_w, &tot) + sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += x[k];
    }
    // the element one past the end carries the grand total
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = sums[gridDim.x];
}

int scan_exclusive_i64(const int64_t *d_in, int64_t *d_out, int64_t n, int64_t *d_tmp, cudaStream_t s) {
    const int64_t nb = ceil_div(n > 0 ? n : 1, SCAN_TILE);
    k_scan_block_sums<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(d_in, n, d_tmp);
    B3C_LAUNCH_CHECK();
    k_scan_sums<<<1, SCAN_THREADS, 0, s>>>(d_tmp, nb);
    B3C_LAUNCH_CHECK();
    k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(d_in, n, d_tmp, d_out);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

}  // namespace b3c

extern "C" {

int b3c_version(void) { return 100; }
const char *b3c_last_error(void) { return b3c::t_err; }
int64_t b3c_launch_count(void) { return b3c::g_launches.load(); }

}  // extern "C"
