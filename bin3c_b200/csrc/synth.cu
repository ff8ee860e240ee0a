// Synthetic pair-record stream generated on the device (bench / test tool, not part of the hot path).
//
// the large BASELINE configs (C3: 500 M pairs = 4 GB, C4: 2 G pairs = 16 GB of packed
// records) should not be drawn on the host and pushed over PCIe for every run, so their pair stream is
// COUNTER-BASED: record t of a community is a pure function of (seed, t) and of the community's tables, which
// makes every sub-range of the stream reproducible on any rank, on any number of GPUs, and on the host
// (bin3c_b200/synth.py: StreamV2.host_records is the NumPy mirror the oracle is fed from; the two are compared
// bit for bit in tests/test_gpu_parity.py).
//
// Draw k of record t:  h = mix64(key + (8 t + k + 1) * 0x9E3779B97F4A7C15),  r = (h >> 11) * 2^-53  in [0, 1)
// (splitmix64's finaliser).  Every floating-point step is a single correctly rounded operation, so the host
// mirror reproduces it exactly.  Recipe of the record (the same as synth.CommunityTables.sample_pairs):
//   end 1: contig ~ length * abundance (upper-bound search in cum_w1)
//   end 2: same contig w.p. p_same; a contig of the same genome ~ length w.p. p_genome; else any contig ~ length
//   ~excl_end_frac of either end lands on an excluded reference; mate order swapped w.p. 1/2; pass bit ~ p_pass
#include "common.cuh"

namespace b3c {

struct SynthArgs {
    const double *cum_w1;          // [N]   normalised cumulative end-1 weight (genome-sorted order s)
    const double *cum_len;         // [N+1] cumulative contig length, cum_len[0] = 0
    const int32_t *genome_sorted;  // [N]   genome of contig s
    const int32_t *g_start;        // [G]   first contig of a genome
    const int32_t *g_end;          // [G]   one past its last contig
    const int32_t *tid_of_s;       // [N]   BAM reference id of contig s
    const int32_t *excl_tids;      // [n_excl] excluded (short) references
    int32_t N, G, n_excl;
    double p_same, p_genome, p_pass, excl_end_frac;
    uint64_t key;
};

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z ^= z >> 30;
    z *= 0xbf58476d1ce4e5b9ull;
    z ^= z >> 27;
    z *= 0x94d049bb133111ebull;
    z ^= z >> 31;
    return z;
}

__device__ __forceinline__ double draw(uint64_t key, uint64_t t, int k) {
    const uint64_t h = mix64(key + (8ull * t + (uint64_t)k + 1ull) * 0x9E3779B97F4A7C15ull);
    return __dmul_rn((double)(h >> 11), 1.1102230246251565e-16);          // 2^-53: exact
}

// first index in [0, n) with a[i] > v, n if none (numpy.searchsorted(a, v, side='right'))
__device__ __forceinline__ int32_t upper_bound(const double *__restrict__ a, int32_t n, double v) {
    int32_t lo = 0, hi = n;
    while (lo < hi) {
        const int32_t mid = (int32_t)(((uint32_t)lo + (uint32_t)hi) >> 1);
        if (__ldg(a + mid) <= v) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_synth_pairs(SynthArgs A, uint64_t first, int64_t count,
                                                     uint64_t *__restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint64_t t = first + (uint64_t)i;
        int32_t s1 = upper_bound(A.cum_w1, A.N, draw(A.key, t, 0));
        if (s1 > A.N - 1) s1 = A.N - 1;
        const double kind = draw(A.key, t, 1);
        int32_t s2 = s1;
        if (kind >= A.p_same) {
            const double r2 = draw(A.key, t, 2);
            int32_t lo = 0, hi = A.N - 1;
            double u;
            if (kind < __dadd_rn(A.p_same, A.p_genome)) {
                const int32_t g = __ldg(A.genome_sorted + s1);
                lo = __ldg(A.g_start + g);
                hi = __ldg(A.g_end + g) - 1;
                const double a = __ldg(A.cum_len + lo), b = __ldg(A.cum_len + hi + 1);
                u = __dadd_rn(a, __dmul_rn(r2, __dsub_rn(b, a)));
            } else {
                u = __dmul_rn(r2, __ldg(A.cum_len + A.N));
            }
            s2 = upper_bound(A.cum_len, A.N + 1, u) - 1;
            if (s2 < lo) s2 = lo;
            if (s2 > hi) s2 = hi;
        }
        int32_t t1 = __ldg(A.tid_of_s + s1), t2 = __ldg(A.tid_of_s + s2);
        const double e = draw(A.key, t, 3);
        if (e < __dmul_rn(2.0, A.excl_end_frac)) {
            int32_t x = (int32_t)__dmul_rn(draw(A.key, t, 4), (double)A.n_excl);
            if (x > A.n_excl - 1) x = A.n_excl - 1;
            const int32_t xt = __ldg(A.excl_tids + x);
            if (e < A.excl_end_frac) t1 = xt;
            else t2 = xt;
        }
        const bool swap = draw(A.key, t, 5) < 0.5;
        const bool pass = draw(A.key, t, 6) < A.p_pass;
        const uint32_t a = (uint32_t)(swap ? t2 : t1), b = (uint32_t)(swap ? t1 : t2);
        out[i] = (uint64_t)a | ((uint64_t)(pass ? 1u : 0u) << 31) | ((uint64_t)b << 32);
    }
}

}  // namespace b3c

using namespace b3c;

extern "C" {

int b3c_synth_pairs(const double *d_cum_w1, const double *d_cum_len, const int32_t *d_genome_sorted,
                    const int32_t *d_g_start, const int32_t *d_g_end, const int32_t *d_tid_of_s,
                    const int32_t *d_excl_tids, int32_t n_contigs, int32_t n_genomes, int32_t n_excl, double p_same,
                    double p_genome, double p_pass, double excl_end_frac, uint64_t seed, uint64_t first, int64_t count,
                    uint64_t *d_records, void *stream) {
    B3C_REQUIRE(d_cum_w1 && d_cum_len && d_genome_sorted && d_g_start && d_g_end && d_tid_of_s && d_excl_tids,
                "null table");
    B3C_REQUIRE(n_contigs > 0 && n_genomes > 0 && n_excl > 0 && count >= 0, "bad sizes");
    if (count == 0) return B3C_OK;
    B3C_REQUIRE(d_records != nullptr, "null output");
    SynthArgs A;
    A.cum_w1 = d_cum_w1;
    A.cum_len = d_cum_len;
    A.genome_sorted = d_genome_sorted;
    A.g_start = d_g_start;
    A.g_end = d_g_end;
    A.tid_of_s = d_tid_of_s;
    A.excl_tids = d_excl_tids;
    A.N = n_contigs;
    A.G = n_genomes;
    A.n_excl = n_excl;
    A.p_same = p_same;
    A.p_genome = p_genome;
    A.p_pass = p_pass;
    A.excl_end_frac = excl_end_frac;
    A.key = mix64(seed);
    int64_t blocks = ceil_div(count, 256 * 8);
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    k_synth_pairs<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(A, first, count, d_records);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

}  // extern "C"
