// Row-segmented operations on the contact matrix: off-diagonal maxima and the acceptance mask,
// site normalisation, diag(x).A.diag(x) scaling, and compress + edge weighting.
// All are one-pass, HBM-bound, warp-per-row segmented loops over CSR.
#include "common.cuh"

namespace b3c {

constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;

static inline unsigned row_grid(int32_t n) {
    int64_t blocks = ceil_div(n > 0 ? n : 1, ROW_WARPS);
    const int64_t cap = (int64_t)kNumSMs * 16;
    return (unsigned)(blocks < cap ? blocks : cap);
}

// ---- max_offdiag (sparse_utils.py:269-281) ----------------------------------------------
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS) k_max_offdiag(int32_t n, int32_t row_lo,
                                                             const int64_t *__restrict__ indptr,
                                                             const int32_t *__restrict__ indices,
                                                             const T *__restrict__ val, T *__restrict__ out) {
    const unsigned lane = lane_id();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); r < n; r += nw) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        T m = T(0);     // the zeroed diagonal is always a member of the column
        for (int64_t e = lo + lane; e < hi; e += 32) {
            const T v = val[e];
            if (indices[e] != (int32_t)r + row_lo && v > m) m = v;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const T t = __shfl_xor_sync(kFullMask, m, o);
            if (t > m) m = t;
        }
        if (lane == 0) out[r] = m;
    }
}

// ---- acceptance mask (contact_map.py:888-905) ------------------------------------------------
__global__ void k_accept_mask(int32_t n, const int32_t *__restrict__ len, const uint32_t *__restrict__ sig,
                              int64_t min_len, int64_t min_sig, uint8_t *__restrict__ mask) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        mask[i] = ((int64_t)len[i] >= min_len && (int64_t)sig[i] >= min_sig) ? 1 : 0;
}

// ---- site normalisation (contact_map.py:1103-1108, 110-113) ---------------------------------
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS) k_site_norm(int32_t n, int32_t row_lo,
                                                           const int64_t *__restrict__ indptr,
                                                           const int32_t *__restrict__ indices,
                                                           const T *in, const int32_t *__restrict__ sites,
                                                           double *out) {   // in may alias out (in-place f64 form)
    const unsigned lane = lane_id();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); r < n; r += nw) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        const int32_t sr = sites[r + row_lo];
        const double si = sr == 0 ? 1.0 : (double)sr;                 // zero sites count as one (Q6)
        for (int64_t e = lo + lane; e < hi; e += 32) {
            const int32_t sc = __ldg(sites + indices[e]);
            const double sj = sc == 0 ? 1.0 : (double)sc;
            const double t = __dmul_rn(si, sj);                       // s_i * s_j
            const double rcp = __ddiv_rn(1.0, t);                     // 1.0 / t
            out[e] = __dmul_rn((double)in[e], rcp);                   // d * r
        }
    }
}

// ---- extent-map length normalisation (contact_map.py:1147-1165, mean_selector :25-46) -----------------
// out[e] = in[e] / (1e-3 * mean(L_r, L_c)), L the length of the sequence a bin belongs to (one entry per bin);
// mean_type 0 geometric (x*y)**0.5, 1 harmonic 2*x*y/(x+y), 2 arithmetic 0.5*(x+y), -1: no normalisation (a plain
// uint32 -> float64 copy).  Operations in the reference's order, each correctly rounded.
template <typename T>
__global__ void __launch_bounds__(ROW_THREADS) k_extent_norm(int32_t n, int32_t row_lo,
                                                             const int64_t *__restrict__ indptr,
                                                             const int32_t *__restrict__ indices, const T *in,
                                                             const double *__restrict__ bin_len, int mean_type,
                                                             double *out) {
    const unsigned lane = lane_id();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); r < n; r += nw) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        const double li = mean_type >= 0 ? bin_len[r + row_lo] : 1.0;
        for (int64_t e = lo + lane; e < hi; e += 32) {
            double v = (double)in[e];
            if (mean_type >= 0) {
                const double lj = __ldg(bin_len + indices[e]);
                double m;
                if (mean_type == 0) m = __dsqrt_rn(__dmul_rn(li, lj));
                else if (mean_type == 1) m = __ddiv_rn(__dmul_rn(__dmul_rn(2.0, li), lj), __dadd_rn(li, lj));
                else m = __dmul_rn(0.5, __dadd_rn(li, lj));
                v = __ddiv_rn(v, __dmul_rn(1e-3, m));
            }
            out[e] = v;
        }
    }
}

// ---- diag(x).A.diag(x) (sparse_utils.py:223-224) ----------------------------------------------
__global__ void __launch_bounds__(ROW_THREADS) k_kr_scale(int32_t n, int32_t row_lo,
                                                          const int64_t *__restrict__ indptr,
                                                          const int32_t *__restrict__ indices,
                                                          const double *__restrict__ a, const double *__restrict__ x,
                                                          double *__restrict__ out) {
    const unsigned lane = lane_id();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    for (int64_t r = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); r < n; r += nw) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        const double xi = x[r + row_lo];
        for (int64_t e = lo + lane; e < hi; e += 32)
            out[e] = __dmul_rn(xi, __dmul_rn(a[e], __ldg(x + indices[e])));     // x_i * (a_ij * x_j), Q9
    }
}

// ---- is_hermitian without the dense temporary (sparse_utils.py:10-18, Q11) -----------------------
__global__ void __launch_bounds__(ROW_THREADS) k_asym_count(int32_t n, const int64_t *__restrict__ indptr,
                                                            const int32_t *__restrict__ indices,
                                                            const double *__restrict__ a, double tol,
                                                            unsigned long long *__restrict__ count) {
    const unsigned lane = lane_id();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    unsigned bad = 0;
    for (int64_t r = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); r < n; r += nw) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        for (int64_t e = lo + lane; e < hi; e += 32) {
            const int32_t c = indices[e];
            int64_t l = indptr[c], h = indptr[c + 1];       // find column r in row c
            while (l < h) {
                const int64_t m = (l + h) >> 1;
                if (indices[m] < (int32_t)r) l = m + 1;
                else h = m;
            }
            const double mirror = (l < indptr[c + 1] && indices[l] == (int32_t)r) ? a[l] : 0.0;
            if (fabs(a[e] - mirror) >= tol) ++bad;
        }
    }
    bad = warp_sum(bad);
    if (lane == 0 && bad) atomicAdd(count, (unsigned long long)bad);
}

// ---- compress + edge weighting ------------------------------------------------------------------
// workspace layout (int64 elements unless noted)
struct CompressWs {
    int64_t o_flag, o_newidx64, o_kept, o_kept_ex, o_edge, o_edge_ex, o_scan, o_max, o_attr;
    int64_t o_nseg, o_seg_ex, o_seg_row, o_seg_len, total;
};
// The fused edge kernels work on row SEGMENTS (a row cut into pieces of seg_len entries, seg_len = max(EDGE_SEG, mean
// row length): at most 2 n_local + 1 of them), so their per-row arrays are sized for that.
constexpr int EDGE_SEG = 4096;
static CompressWs compress_layout(int32_t n) {
    Carver c;
    CompressWs w;
    const int64_t n1 = (int64_t)n + 1, n2 = 2 * (int64_t)n + 2;
    w.o_flag = c.take(n1 * 8);
    w.o_newidx64 = c.take(n1 * 8);
    w.o_kept = c.take(n2 * 8);
    w.o_kept_ex = c.take(n2 * 8);
    w.o_edge = c.take(n2 * 8);
    w.o_edge_ex = c.take(n2 * 8);
    w.o_scan = c.take(scan_tmp_elems(n2) * 8);
    w.o_max = c.take(64);
    w.o_attr = c.take((int64_t)n * 16);
    w.o_nseg = c.take(n1 * 8);
    w.o_seg_ex = c.take(n1 * 8);
    w.o_seg_row = c.take(n2 * 4);
    w.o_seg_len = c.take(64);
    w.total = c.cur;
    return w;
}

// ---- row segments: a heavy-tailed community has rows of 10^5 entries next to rows of ten; a warp per ROW leaves
// the longest row to one warp (its trips were C4's edge stage), a warp per SEGMENT does not --------------------------
__global__ void k_seg_len(int32_t n_local, const int64_t *__restrict__ indptr, int64_t *__restrict__ seg_len) {
    const int64_t mean = (indptr[n_local] - indptr[0] + n_local - 1) / (n_local > 0 ? n_local : 1);
    *seg_len = mean > EDGE_SEG ? mean : EDGE_SEG;
}
__global__ void k_seg_count(int32_t n_local, const int64_t *__restrict__ indptr, const int64_t *__restrict__ seg_len,
                            int64_t *__restrict__ nseg) {
    const int64_t T = *seg_len;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_local; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t len = indptr[r + 1] - indptr[r];
        nseg[r] = len > T ? (len + T - 1) / T : 1;
    }
}
__global__ void k_seg_rows(int32_t n_local, const int64_t *__restrict__ seg_ex, int32_t *__restrict__ seg_row) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_local; r += (int64_t)gridDim.x * blockDim.x)
        for (int64_t q = seg_ex[r]; q < seg_ex[r + 1]; ++q) seg_row[q] = (int32_t)r;
}

__global__ void k_mask_flags(int32_t n, const uint8_t *__restrict__ mask, int64_t *__restrict__ flag) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        flag[i] = mask[i] ? 1 : 0;
}

__global__ void k_newidx(int32_t n, const uint8_t *__restrict__ mask, const int64_t *__restrict__ ex,
                         int32_t *__restrict__ newidx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        newidx[i] = mask[i] ? (int32_t)ex[i] : -1;
}

// per OLD row: kept entries, kept upper-triangle entries (edges), running max of kept values
__global__ void __launch_bounds__(ROW_THREADS) k_compress_count(int32_t n, int32_t row_lo,
                                                                const int64_t *__restrict__ indptr,
                                                                const int32_t *__restrict__ indices,
                                                                const double *__restrict__ data,
                                                                const uint8_t *__restrict__ mask,
                                                                int64_t *__restrict__ kept, int64_t *__restrict__ edge,
                                                                unsigned long long *__restrict__ vmax) {
    const unsigned lane = lane_id();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    double wmax = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); r < n; r += nw) {
        unsigned k = 0, ed = 0;
        if (mask[r + row_lo]) {
            const int64_t lo = indptr[r], hi = indptr[r + 1];
            for (int64_t e = lo + lane; e < hi; e += 32) {
                const int32_t c = indices[e];
                if (__ldg(mask + c)) {
                    ++k;
                    ed += (c >= (int32_t)r + row_lo) ? 1u : 0u;
                    if (data) wmax = fmax(wmax, data[e]);
                }
            }
        }
        k = warp_sum(k);
        ed = warp_sum(ed);
        if (lane == 0) {
            kept[r] = k;
            edge[r] = ed;
        }
    }
    wmax = warp_max(wmax);
    // non-negative doubles order like their bit patterns
    if (lane == 0 && wmax > 0.0) atomicMax(vmax, (unsigned long long)__double_as_longlong(wmax));
}

__global__ void __launch_bounds__(ROW_THREADS) k_compress_fill(
    int32_t n, int32_t row_lo, const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
    const double *__restrict__ data, const uint8_t *__restrict__ mask, const int32_t *__restrict__ newidx,
    const int64_t *__restrict__ kept_ex, const int64_t *__restrict__ edge_ex,
    const int64_t *__restrict__ n_accepted, const double *__restrict__ vmax, int scale,
    int64_t *__restrict__ sub_indptr,
    int32_t *__restrict__ sub_indices, double *__restrict__ sub_data, int32_t *__restrict__ eu,
    int32_t *__restrict__ ev, double *__restrict__ ew, double *__restrict__ scl_out) {
    const unsigned lane = lane_id(), lt = lanemask_lt();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    const double vm = *vmax;
    const double scl = scale ? __ddiv_rn(1.0, vm) : 1.0;                 // cluster.py:316
    if (blockIdx.x == 0 && threadIdx.x == 0 && scl_out) *scl_out = scl;
    for (int64_t r = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); r <= n; r += nw) {
        if (r == n) {
            // one past the last accepted row closes the compressed indptr
            if (lane == 0 && sub_indptr) sub_indptr[*n_accepted] = kept_ex[n];
            continue;
        }
        if (!mask[r + row_lo]) continue;
        const int32_t nr = newidx[r + row_lo];
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        int64_t kbase = kept_ex[r], ebase = edge_ex[r];
        if (lane == 0 && sub_indptr) sub_indptr[nr] = kbase;
        for (int64_t e0 = lo; e0 < hi; e0 += 32) {
            const int64_t e = e0 + lane;
            int32_t c = -1;
            bool keep = false;
            if (e < hi) {
                c = indices[e];
                keep = __ldg(mask + c) != 0;
            }
            const bool is_edge = keep && c >= (int32_t)r + row_lo;
            const unsigned mk = __ballot_sync(kFullMask, keep);
            const unsigned me = __ballot_sync(kFullMask, is_edge);
            if (keep) {
                const double v = data[e];
                const int32_t nc = __ldg(newidx + c);
                if (sub_indices) {
                    const int64_t d = kbase + __popc(mk & lt);
                    sub_indices[d] = nc;
                    sub_data[d] = v;
                }
                if (is_edge && eu) {
                    const int64_t d = ebase + __popc(me & lt);
                    eu[d] = nr;
                    ev[d] = nc;
                    ew[d] = __dmul_rn(v, scl);                              // cluster.py:321
                }
            }
            kbase += __popc(mk);
            ebase += __popc(me);
        }
    }
}

// ---- fused form: edge list straight from counts, sites and the KR scale vector -----------------------
// The value of an entry is recomputed where it is consumed, x_i * ((count / (s_i s_j)) * x_j) -- what
// k_site_norm followed by k_kr_scale would have stored -- so neither intermediate matrix exists.  What a
// column needs (gapless id or -1, site count, x) is packed into one 16-byte record per contig, so an
// entry costs ONE scattered 128-bit gather instead of four scalar ones (a scattered warp load costs one
// L1 wavefront per distinct line whatever its width).
struct __align__(16) ContigAttr {
    int32_t newidx;          // index among the accepted contigs, -1 if rejected
    int32_t site;
    double x;
};
static_assert(sizeof(ContigAttr) == 16, "one 128-bit gather per entry");

__device__ __forceinline__ ContigAttr ld_attr(const ContigAttr *p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    ContigAttr a;
    a.newidx = (int32_t)v.x;
    a.site = (int32_t)v.y;
    a.x = __hiloint2double((int)v.w, (int)v.z);
    return a;
}

__global__ void k_edge_attr(int32_t n, const uint8_t *__restrict__ mask, const int64_t *__restrict__ ex,
                            const int32_t *__restrict__ sites, const double *__restrict__ x,
                            ContigAttr *__restrict__ attr, int32_t *__restrict__ newidx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        ContigAttr a;
        a.newidx = mask[i] ? (int32_t)ex[i] : -1;
        a.site = sites[i];
        a.x = x[i];
        attr[i] = a;
        newidx[i] = a.newidx;
    }
}

__device__ __forceinline__ double edge_value(uint32_t count, const ContigAttr &ar, const ContigAttr &ac) {
    return __dmul_rn(ar.x, __dmul_rn(site_scaled(count, ar.site, ac.site), ac.x));
}

__global__ void __launch_bounds__(ROW_THREADS) k_edges_count(int32_t row_lo, const int64_t *__restrict__ indptr,
                                                             const int32_t *__restrict__ indices,
                                                             const uint32_t *__restrict__ counts,
                                                             const ContigAttr *__restrict__ attr,
                                                             const int64_t *__restrict__ seg_ex,
                                                             const int32_t *__restrict__ seg_row, int32_t n_local,
                                                             const int64_t *__restrict__ seg_len,
                                                             int64_t *__restrict__ kept, int64_t *__restrict__ edge,
                                                             unsigned long long *__restrict__ vmax) {
    const unsigned lane = lane_id();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    const int64_t n_seg = seg_ex[n_local], T = *seg_len;
    double wmax = 0.0;
    for (int64_t q = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); q < n_seg; q += nw) {
        unsigned k = 0, ed = 0;
        const int32_t r = seg_row[q];
        const int32_t gr = r + row_lo;
        const ContigAttr ar = ld_attr(attr + gr);
        if (ar.newidx >= 0) {
            const int64_t lo = indptr[r] + (q - seg_ex[r]) * T;
            const int64_t hi = min(indptr[r + 1], lo + T);
            // four 32-entry windows per trip: their columns and counts, then their four scattered gathers, are in
            // flight together
#pragma unroll 1
            for (int64_t e0 = lo + lane; e0 < hi; e0 += 128) {
                int32_t c[4];
                uint32_t cnt[4];
                ContigAttr ac[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t e = e0 + 32 * u;
                    c[u] = e < hi ? indices[e] : -1;
                    cnt[u] = e < hi ? counts[e] : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ac[u].newidx = -1;
                    if (c[u] >= 0) ac[u] = ld_attr(attr + c[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (ac[u].newidx >= 0) {
                        ++k;
                        ed += (c[u] >= gr) ? 1u : 0u;
                        wmax = fmax(wmax, edge_value(cnt[u], ar, ac[u]));
                    }
            }
        }
        k = warp_sum(k);
        ed = warp_sum(ed);
        if (lane == 0) {
            kept[q] = k;
            edge[q] = ed;
        }
    }
    wmax = warp_max(wmax);
    if (lane == 0 && wmax > 0.0) atomicMax(vmax, (unsigned long long)__double_as_longlong(wmax));
}

__global__ void __launch_bounds__(ROW_THREADS) k_edges_fill(int32_t row_lo, const int64_t *__restrict__ indptr,
                                                            const int32_t *__restrict__ indices,
                                                            const uint32_t *__restrict__ counts,
                                                            const ContigAttr *__restrict__ attr,
                                                            const int64_t *__restrict__ seg_ex,
                                                            const int32_t *__restrict__ seg_row, int32_t n_local,
                                                            const int64_t *__restrict__ seg_len,
                                                            const int64_t *__restrict__ edge_ex,
                                                            const double *__restrict__ vmax, int scale,
                                                            int32_t *__restrict__ eu, int32_t *__restrict__ ev,
                                                            double *__restrict__ ew, double *__restrict__ scl_out) {
    const unsigned lane = lane_id(), lt = lanemask_lt();
    const int64_t nw = (int64_t)gridDim.x * ROW_WARPS;
    const int64_t n_seg = seg_ex[n_local], T = *seg_len;
    const double scl = scale ? __ddiv_rn(1.0, *vmax) : 1.0;                 // cluster.py:316
    if (blockIdx.x == 0 && threadIdx.x == 0 && scl_out) *scl_out = scl;
    for (int64_t q = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5); q < n_seg; q += nw) {
        const int32_t r = seg_row[q];
        const int32_t gr = r + row_lo;
        const ContigAttr ar = ld_attr(attr + gr);
        if (ar.newidx < 0) continue;
        const int64_t lo = indptr[r] + (q - seg_ex[r]) * T;
        const int64_t hi = min(indptr[r + 1], lo + T);
        if (edge_ex[q + 1] == edge_ex[q]) continue;                         // no upper-triangle entry in this piece
        int64_t ebase = edge_ex[q];
        // columns are sorted: the upper-triangle entries (c >= row) are the tail of the row; the edges keep the row's
        // column order (segments are numbered along the row, four windows per trip as in k_edges_count)
#pragma unroll 1
        for (int64_t e0 = lo + lane; e0 < hi + lane; e0 += 128) {           // (+ lane: every lane makes the same trips)
            int32_t c[4];
            uint32_t cnt[4];
            ContigAttr ac[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t e = e0 + 32 * u;
                c[u] = e < hi ? indices[e] : -1;
                cnt[u] = e < hi ? counts[e] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ac[u].newidx = -1;
                if (c[u] >= gr) ac[u] = ld_attr(attr + c[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool is_edge = ac[u].newidx >= 0;
                const unsigned me = __ballot_sync(kFullMask, is_edge);
                if (is_edge) {
                    const int64_t d = ebase + __popc(me & lt);
                    eu[d] = ar.newidx;
                    ev[d] = ac[u].newidx;
                    ew[d] = __dmul_rn(edge_value(cnt[u], ar, ac[u]), scl);  // cluster.py:321
                }
                ebase += __popc(me);
            }
        }
    }
}

}  // namespace b3c

using namespace b3c;

extern "C" {

int b3c_max_offdiag_u32(int32_t n_local, int32_t row_lo, const int64_t *d_indptr, const int32_t *d_indices,
                        const uint32_t *d_counts, uint32_t *d_signal, void *stream) {
    B3C_REQUIRE(n_local > 0 && row_lo >= 0 && d_indptr && d_signal, "bad arguments");
    k_max_offdiag<uint32_t><<<row_grid(n_local), ROW_THREADS, 0, (cudaStream_t)stream>>>(
        n_local, row_lo, d_indptr, d_indices, d_counts, d_signal);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_max_offdiag_f64(int32_t n_local, int32_t row_lo, const int64_t *d_indptr, const int32_t *d_indices,
                        const double *d_data, double *d_signal, void *stream) {
    B3C_REQUIRE(n_local > 0 && row_lo >= 0 && d_indptr && d_signal, "bad arguments");
    k_max_offdiag<double><<<row_grid(n_local), ROW_THREADS, 0, (cudaStream_t)stream>>>(
        n_local, row_lo, d_indptr, d_indices, d_data, d_signal);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_acceptance_mask(int32_t n, const int32_t *d_lengths, const uint32_t *d_signal, int64_t min_len,
                        int64_t min_sig, uint8_t *d_mask, void *stream) {
    B3C_REQUIRE(n > 0 && d_lengths && d_signal && d_mask, "bad arguments");
    k_accept_mask<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(n, d_lengths, d_signal, min_len,
                                                                               min_sig, d_mask);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_site_norm(int32_t n_local, int32_t row_lo, const int64_t *d_indptr, const int32_t *d_indices,
                  const uint32_t *d_counts, const int32_t *d_sites, double *d_out, void *stream) {
    B3C_REQUIRE(n_local > 0 && row_lo >= 0 && d_indptr && d_sites && d_out, "bad arguments");
    k_site_norm<uint32_t><<<row_grid(n_local), ROW_THREADS, 0, (cudaStream_t)stream>>>(
        n_local, row_lo, d_indptr, d_indices, d_counts, d_sites, d_out);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_site_norm_f64(int32_t n_local, int32_t row_lo, const int64_t *d_indptr, const int32_t *d_indices,
                      double *d_data, const int32_t *d_sites, void *stream) {
    B3C_REQUIRE(n_local > 0 && row_lo >= 0 && d_indptr && d_sites && d_data, "bad arguments");
    k_site_norm<double><<<row_grid(n_local), ROW_THREADS, 0, (cudaStream_t)stream>>>(
        n_local, row_lo, d_indptr, d_indices, d_data, d_sites, d_data);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_extent_norm(int32_t n_local, int32_t row_lo, const int64_t *d_indptr, const int32_t *d_indices,
                    const uint32_t *d_counts, const double *d_bin_len, int32_t mean_type, double *d_out,
                    void *stream) {
    B3C_REQUIRE(n_local > 0 && row_lo >= 0 && d_indptr && d_out, "bad arguments");
    B3C_REQUIRE(mean_type >= -1 && mean_type <= 2 && (mean_type < 0 || d_bin_len), "mean_type must be -1..2");
    k_extent_norm<uint32_t><<<row_grid(n_local), ROW_THREADS, 0, (cudaStream_t)stream>>>(
        n_local, row_lo, d_indptr, d_indices, d_counts, d_bin_len, mean_type, d_out);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_kr_scale(int32_t n_local, int32_t row_lo, const int64_t *d_indptr, const int32_t *d_indices,
                 const double *d_data, const double *d_x, double *d_out, void *stream) {
    B3C_REQUIRE(n_local > 0 && row_lo >= 0 && d_indptr && d_x && d_out, "bad arguments");
    k_kr_scale<<<row_grid(n_local), ROW_THREADS, 0, (cudaStream_t)stream>>>(n_local, row_lo, d_indptr, d_indices,
                                                                            d_data, d_x, d_out);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_asymmetry_count(int32_t n, const int64_t *d_indptr, const int32_t *d_indices, const double *d_data,
                        double tol, uint64_t *d_scratch, int64_t *h_count, void *stream) {
    B3C_REQUIRE(n > 0 && d_indptr && d_scratch && h_count, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    B3C_CUDA(cudaMemsetAsync(d_scratch, 0, 8, s));
    k_asym_count<<<row_grid(n), ROW_THREADS, 0, s>>>(n, d_indptr, d_indices, d_data, tol,
                                                     (unsigned long long *)d_scratch);
    B3C_LAUNCH_CHECK();
    B3C_CUDA(cudaMemcpyAsync(h_count, d_scratch, 8, cudaMemcpyDeviceToHost, s));
    B3C_CUDA(cudaStreamSynchronize(s));
    return B3C_OK;
}

int64_t b3c_compress_workspace_bytes(int32_t n) {
    if (n <= 0) return B3C_ERR_ARG;
    return compress_layout(n).total;
}

int b3c_compress_count(int32_t n, int32_t row_lo, int32_t n_local, const int64_t *d_indptr,
                       const int32_t *d_indices, const double *d_data, const uint8_t *d_mask, int32_t *d_newidx,
                       void *d_ws, int64_t ws_bytes, double *d_vmax, int64_t *h_out, void *stream) {
    B3C_REQUIRE(n > 0 && n_local > 0 && row_lo >= 0 && row_lo + n_local <= n, "bad row block");
    B3C_REQUIRE(d_indptr && d_mask && d_newidx && d_ws && d_vmax && h_out, "null pointer");
    const CompressWs w = compress_layout(n);
    if (ws_bytes < w.total) {
        set_error("compress workspace too small: %lld < %lld", (long long)ws_bytes, (long long)w.total);
        return B3C_ERR_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    int64_t *flag = (int64_t *)(ws + w.o_flag), *nidx = (int64_t *)(ws + w.o_newidx64);
    int64_t *kept = (int64_t *)(ws + w.o_kept), *kept_ex = (int64_t *)(ws + w.o_kept_ex);
    int64_t *edge = (int64_t *)(ws + w.o_edge), *edge_ex = (int64_t *)(ws + w.o_edge_ex);
    int64_t *scan = (int64_t *)(ws + w.o_scan);
    B3C_CUDA(cudaMemsetAsync(d_vmax, 0, 8, s));
    const unsigned g = (unsigned)ceil_div(n, 256);
    // gapless ids over the WHOLE mask (every rank computes the same table)
    k_mask_flags<<<g, 256, 0, s>>>(n, d_mask, flag);
    B3C_LAUNCH_CHECK();
    int rc = scan_exclusive_i64(flag, nidx, n, scan, s);
    if (rc) return rc;
    k_newidx<<<g, 256, 0, s>>>(n, d_mask, nidx, d_newidx);
    B3C_LAUNCH_CHECK();
    // non-negative doubles order like their bit patterns, so the running maximum is kept as bits
    k_compress_count<<<row_grid(n_local), ROW_THREADS, 0, s>>>(n_local, row_lo, d_indptr, d_indices, d_data, d_mask,
                                                               kept, edge, (unsigned long long *)d_vmax);
    B3C_LAUNCH_CHECK();
    rc = scan_exclusive_i64(kept, kept_ex, n_local, scan, s);
    if (rc) return rc;
    rc = scan_exclusive_i64(edge, edge_ex, n_local, scan, s);
    if (rc) return rc;
    B3C_CUDA(cudaMemcpyAsync(&h_out[0], nidx + n, 8, cudaMemcpyDeviceToHost, s));
    B3C_CUDA(cudaMemcpyAsync(&h_out[1], kept_ex + n_local, 8, cudaMemcpyDeviceToHost, s));
    B3C_CUDA(cudaMemcpyAsync(&h_out[2], edge_ex + n_local, 8, cudaMemcpyDeviceToHost, s));
    B3C_CUDA(cudaStreamSynchronize(s));
    return B3C_OK;
}

int b3c_compress_fill(int32_t n, int32_t row_lo, int32_t n_local, const int64_t *d_indptr,
                      const int32_t *d_indices, const double *d_data, const uint8_t *d_mask,
                      const int32_t *d_newidx, void *d_ws, const double *d_vmax, int scale,
                      int64_t *d_sub_indptr, int32_t *d_sub_indices, double *d_sub_data, int32_t *d_edge_u,
                      int32_t *d_edge_v, double *d_edge_w, double *d_scl, void *stream) {
    B3C_REQUIRE(n > 0 && n_local > 0 && row_lo >= 0 && row_lo + n_local <= n, "bad row block");
    B3C_REQUIRE(d_indptr && d_data && d_mask && d_newidx && d_ws && d_vmax, "null pointer");
    B3C_REQUIRE((d_sub_indices == nullptr) == (d_sub_data == nullptr), "sub_indices/sub_data must come together");
    B3C_REQUIRE(d_sub_indptr == nullptr || (row_lo == 0 && n_local == n),
                "the compressed matrix is only produced for a whole matrix, not a row block");
    B3C_REQUIRE((d_edge_u == nullptr) == (d_edge_v == nullptr) && (d_edge_u == nullptr) == (d_edge_w == nullptr),
                "edge arrays must come together");
    const CompressWs w = compress_layout(n);
    char *ws = (char *)d_ws;
    k_compress_fill<<<row_grid(n_local + 1), ROW_THREADS, 0, (cudaStream_t)stream>>>(
        n_local, row_lo, d_indptr, d_indices, d_data, d_mask, d_newidx, (const int64_t *)(ws + w.o_kept_ex),
        (const int64_t *)(ws + w.o_edge_ex), (const int64_t *)(ws + w.o_newidx64) + n, d_vmax, scale, d_sub_indptr,
        d_sub_indices, d_sub_data, d_edge_u, d_edge_v, d_edge_w, d_scl);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_edges_count(int32_t n, int32_t row_lo, int32_t n_local, const int64_t *d_indptr, const int32_t *d_indices,
                    const uint32_t *d_counts, const int32_t *d_sites, const double *d_x, const uint8_t *d_mask,
                    int32_t *d_newidx, void *d_ws, int64_t ws_bytes, double *d_vmax, int64_t *h_out, void *stream) {
    B3C_REQUIRE(n > 0 && n_local > 0 && row_lo >= 0 && row_lo + n_local <= n, "bad row block");
    B3C_REQUIRE(d_indptr && d_counts && d_sites && d_x && d_mask && d_newidx && d_ws && d_vmax && h_out, "null pointer");
    const CompressWs w = compress_layout(n);
    if (ws_bytes < w.total) {
        set_error("compress workspace too small: %lld < %lld", (long long)ws_bytes, (long long)w.total);
        return B3C_ERR_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    int64_t *flag = (int64_t *)(ws + w.o_flag), *nidx = (int64_t *)(ws + w.o_newidx64);
    int64_t *kept = (int64_t *)(ws + w.o_kept), *kept_ex = (int64_t *)(ws + w.o_kept_ex);
    int64_t *edge = (int64_t *)(ws + w.o_edge), *edge_ex = (int64_t *)(ws + w.o_edge_ex);
    int64_t *scan = (int64_t *)(ws + w.o_scan);
    ContigAttr *attr = (ContigAttr *)(ws + w.o_attr);
    B3C_CUDA(cudaMemsetAsync(d_vmax, 0, 8, s));
    const unsigned g = (unsigned)ceil_div(n, 256);
    k_mask_flags<<<g, 256, 0, s>>>(n, d_mask, flag);
    B3C_LAUNCH_CHECK();
    int rc = scan_exclusive_i64(flag, nidx, n, scan, s);
    if (rc) return rc;
    k_edge_attr<<<g, 256, 0, s>>>(n, d_mask, nidx, d_sites, d_x, attr, d_newidx);
    B3C_LAUNCH_CHECK();
    // row segments (at most 2 n_local + 1), then a warp per segment
    int64_t *nseg = (int64_t *)(ws + w.o_nseg), *seg_ex = (int64_t *)(ws + w.o_seg_ex);
    int64_t *seg_len = (int64_t *)(ws + w.o_seg_len);
    int32_t *seg_row = (int32_t *)(ws + w.o_seg_row);
    const int64_t n2 = 2 * (int64_t)n_local + 1;
    const unsigned gl = (unsigned)ceil_div(n_local, 256);
    k_seg_len<<<1, 1, 0, s>>>(n_local, d_indptr, seg_len);
    B3C_LAUNCH_CHECK();
    k_seg_count<<<gl, 256, 0, s>>>(n_local, d_indptr, seg_len, nseg);
    B3C_LAUNCH_CHECK();
    rc = scan_exclusive_i64(nseg, seg_ex, n_local, scan, s);
    if (rc) return rc;
    k_seg_rows<<<gl, 256, 0, s>>>(n_local, seg_ex, seg_row);
    B3C_LAUNCH_CHECK();
    B3C_CUDA(cudaMemsetAsync(kept, 0, (size_t)n2 * 8, s));
    B3C_CUDA(cudaMemsetAsync(edge, 0, (size_t)n2 * 8, s));
    k_edges_count<<<row_grid(n_local), ROW_THREADS, 0, s>>>(row_lo, d_indptr, d_indices, d_counts, attr, seg_ex, seg_row,
                                                            n_local, seg_len, kept, edge, (unsigned long long *)d_vmax);
    B3C_LAUNCH_CHECK();
    rc = scan_exclusive_i64(kept, kept_ex, n2, scan, s);
    if (rc) return rc;
    rc = scan_exclusive_i64(edge, edge_ex, n2, scan, s);
    if (rc) return rc;
    B3C_CUDA(cudaMemcpyAsync(&h_out[0], nidx + n, 8, cudaMemcpyDeviceToHost, s));
    B3C_CUDA(cudaMemcpyAsync(&h_out[1], kept_ex + n2, 8, cudaMemcpyDeviceToHost, s));
    B3C_CUDA(cudaMemcpyAsync(&h_out[2], edge_ex + n2, 8, cudaMemcpyDeviceToHost, s));
    B3C_CUDA(cudaStreamSynchronize(s));
    return B3C_OK;
}

int b3c_edges_fill(int32_t n, int32_t row_lo, int32_t n_local, const int64_t *d_indptr, const int32_t *d_indices,
                   const uint32_t *d_counts, void *d_ws, const double *d_vmax, int scale, int32_t *d_edge_u,
                   int32_t *d_edge_v, double *d_edge_w, double *d_scl, void *stream) {
    B3C_REQUIRE(n > 0 && n_local > 0 && row_lo >= 0 && row_lo + n_local <= n, "bad row block");
    B3C_REQUIRE(d_indptr && d_counts && d_ws && d_vmax && d_edge_u && d_edge_v && d_edge_w, "null pointer");
    // the per-contig records b3c_edges_count left in d_ws hold the site counts, x and the gapless ids
    const CompressWs w = compress_layout(n);
    char *ws = (char *)d_ws;
    k_edges_fill<<<row_grid(n_local), ROW_THREADS, 0, (cudaStream_t)stream>>>(
        row_lo, d_indptr, d_indices, d_counts, (const ContigAttr *)(ws + w.o_attr), (const int64_t *)(ws + w.o_seg_ex),
        (const int32_t *)(ws + w.o_seg_row), n_local, (const int64_t *)(ws + w.o_seg_len),
        (const int64_t *)(ws + w.o_edge_ex), d_vmax, scale, d_edge_u, d_edge_v, d_edge_w, d_scl);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

}  // extern "C"
