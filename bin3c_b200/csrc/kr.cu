// Knight-Ruiz balancing (sparse_utils.py:90-224) as fp64 CSR SpMV + fused vector phases.
//
// One persistent cooperative kernel runs the whole Newton/CG iteration with device-side
// control flow: every CTA derives the loop scalars from the same per-chunk partial sums in
// the same order, so all CTAs take identical branches and the host is not involved until the
// scale vector is final.  The same phase functions are exposed one-by-one (b3c_krp_*) for
// the multi-GPU row-block driver, which puts NCCL collectives between them.
//
// SpMV: the non-zeros are cut into fixed tiles of SPMV_TILE entries regardless of row
// boundaries (nnz-balanced, so heavy-tailed contig rows cost nothing extra).  A CTA streams
// a tile's values and column indices with coalesced loads, gathers u[col], parks the products
// in shared memory and reduces them per row.  Rows that straddle tiles leave partial sums that
// a tiny fix-up pass adds in tile order, so the result is deterministic.
//
// Reductions (dot products, min, max) use fixed 1024-row chunks with a fixed tree inside the
// chunk and an in-order sum over chunks: the value does not depend on the grid size or on how
// rows are split over GPUs (row blocks are chunk aligned).
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace b3c {

constexpr int KR_THREADS = 256;
constexpr int KR_WARPS = KR_THREADS / 32;
constexpr int SPMV_NPT = 8;
constexpr int SPMV_TILE = KR_THREADS * SPMV_NPT;     // 2048 non-zeros = 24 KB of matrix per tile
constexpr int CHUNK = 1024;                          // rows per reduction chunk
constexpr int CHUNK_RPT = CHUNK / KR_THREADS;
constexpr int SPTR_CAP = 1024;                       // row pointers of a tile staged in shared memory
constexpr int RED_MAX = 8;                           // values reduced together by one block reduction
constexpr int KR_MIN_CTAS = 4;                       // CTAs per SM the persistent kernel is compiled for

// partial arrays, each n_chunks long
enum { PA = 0, PB, PC, PMIN, PNEGMAX, PG1, PG2, P_COUNT };

struct KRScalars {
    double tol, delta, Delta, rt, stop_tol;
    double rho_km1, rho_km2, rout, rold, eta, inner_tol, alpha, beta, gamma;
    long long n_iter, max_iter, k, outer, n_spmv, zero_diag;
    int status, ymode, ysel, state;
};

// per-phase cycle counters of CTA 0 (work = its own phase time, sync = its wait at the grid barrier)
enum { T_INIT = 0, T_SPMV, T_FIX, T_RESID, T_DIR, T_W, T_STEP, T_UPDATE, T_SCALAR, T_COUNT };
struct KRTimers {
    long long work[T_COUNT], sync[T_COUNT], total;
};

struct KRArgs {
    // local row block [row_lo, row_hi) of an n x n matrix; indptr is local (0-based), columns global
    int32_t n, row_lo, row_hi;
    int64_t nnz;
    const int64_t *indptr;
    const int32_t *indices;
    const double *data;
    // plan
    int64_t n_tiles;
    int32_t *tile_ra;
    int32_t *chunk_t;          // per local chunk: first tile whose last-starting row lies in the chunk
    double *head_part, *tail_part;
    double *dfix;
    // vectors (global row indexing, length n)
    double *x, *v, *rk, *y0, *y1, *p, *Z, *w, *u, *q;
    // partials [P_COUNT][n_chunks]
    double *part;
    int32_t n_chunks;
    KRScalars *ctl;
    KRTimers *timers;
};

// ---- deterministic block reductions ---------------------------------------------------------
// NS sums followed by NM minima, reduced together: warp xor-tree, then the 8 warp results in warp
// order.  Every thread gets the results; the shape is fixed, so the value is reproducible.
template <int NS, int NM>
__device__ __forceinline__ void block_reduce(double (&v)[NS + NM], double *s_red) {
    constexpr int K = NS + NM;
    static_assert(K <= RED_MAX, "too many values");
#pragma unroll
    for (int i = 0; i < K; ++i) v[i] = (i < NS) ? warp_sum(v[i]) : warp_min(v[i]);
    const unsigned w = threadIdx.x >> 5;
    __syncthreads();
    if (lane_id() == 0) {
#pragma unroll
        for (int i = 0; i < K; ++i) s_red[w * K + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < K; ++i) {
        double r = s_red[i];
#pragma unroll
        for (int j = 1; j < KR_WARPS; ++j) r = (i < NS) ? r + s_red[j * K + i] : fmin(r, s_red[j * K + i]);
        v[i] = r;
    }
}

// the same over per-chunk partial arrays: out[i] = reduce(part[ids[i]][0..nc)); identical in every CTA
template <int NS, int NM>
__device__ __forceinline__ void reduce_parts(const double *part, int nc, const int (&ids)[NS + NM],
                                             double (&out)[NS + NM], double *s_red) {
#pragma unroll
    for (int i = 0; i < NS + NM; ++i) out[i] = (i < NS) ? 0.0 : (double)INFINITY;
    for (int j = threadIdx.x; j < nc; j += KR_THREADS) {
#pragma unroll
        for (int i = 0; i < NS + NM; ++i) {
            const double x = part[(int64_t)ids[i] * nc + j];
            out[i] = (i < NS) ? out[i] + x : fmin(out[i], x);
        }
    }
    block_reduce<NS, NM>(out, s_red);
}

// ---- SpMV -------------------------------------------------------------------------------------
// `u` is rewritten between SpMV phases of the same (persistent) launch, so it is read with
// ordinary coherent loads -- never ld.global.nc -- and carries no __restrict__.
__device__ __forceinline__ void spmv_tile(const KRArgs &A, const double *u, int64_t t, double *s_prod,
                                          int *s_ptr) {
    const int64_t base = t * SPMV_TILE;
    const int64_t rem = A.nnz - base;
    const int cnt = (int)(rem < SPMV_TILE ? (rem > 0 ? rem : 0) : SPMV_TILE);
    // 1. stream the tile: all loads are issued before anything is consumed
    double a[SPMV_NPT];
    int c[SPMV_NPT];
#pragma unroll
    for (int k = 0; k < SPMV_NPT; ++k) {
        const int idx = k * KR_THREADS + threadIdx.x;
        if (idx < cnt) {
            a[k] = ld_stream_f64(A.data + base + idx);
            c[k] = ld_stream_s32(A.indices + base + idx);
        } else {
            a[k] = 0.0;
            c[k] = 0;
        }
    }
    // 2. rows that start inside the tile are [ra, rb); their pointers go to shared memory while the
    //    matrix loads are in flight, so the reduction below never waits on global memory
    const int ra = A.tile_ra[t], rb = A.tile_ra[t + 1];
    const int n_rows = rb - ra;
    const bool staged = n_rows < SPTR_CAP;
    if (staged) {
        for (int i = threadIdx.x; i <= n_rows; i += KR_THREADS) {
            const int64_t rel = A.indptr[ra + i] - base;
            s_ptr[i] = (int)(rel > cnt ? cnt + 1 : rel);              // cnt+1 marks "ends beyond this tile"
        }
    }
    // 3. gather u[col], multiply, park the products
#pragma unroll
    for (int k = 0; k < SPMV_NPT; ++k) {
        const int idx = k * KR_THREADS + threadIdx.x;
        if (idx < cnt) s_prod[idx] = a[k] * u[c[k]];
    }
    __syncthreads();
    // 4. segmented reduction: item 0 is the head of a row begun in an earlier tile, item i>0 is row ra+i-1
    const int n_items = n_rows + 1;
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    auto bounds = [&](int it, int &lo, int &hi, bool &complete) {
        int p0, p1;
        if (staged) {
            p0 = (it == 0) ? 0 : s_ptr[it - 1];
            p1 = s_ptr[it == 0 ? 0 : it];
        } else {
            const int64_t r0 = (it == 0) ? 0 : A.indptr[ra + it - 1] - base;
            const int64_t r1 = A.indptr[ra + (it == 0 ? 0 : it)] - base;
            p0 = (int)r0;
            p1 = (int)(r1 > cnt ? cnt + 1 : r1);
        }
        lo = p0;
        hi = p1 > cnt ? cnt : p1;
        complete = p1 <= cnt;
    };
    if (n_items <= 8 * KR_WARPS) {
        for (int it = warp; it < n_items; it += KR_WARPS) {
            int lo, hi;
            bool complete;
            bounds(it, lo, hi, complete);
            double s = 0.0;
            for (int e = lo + (int)lane; e < hi; e += 32) s += s_prod[e];
            s = warp_sum(s);
            if (lane == 0) {
                if (it == 0) A.head_part[t] = s;
                else if (complete) A.q[A.row_lo + ra + it - 1] = s;
                else A.tail_part[t] = s;
            }
        }
    } else {
        for (int it = threadIdx.x; it < n_items; it += KR_THREADS) {
            int lo, hi;
            bool complete;
            bounds(it, lo, hi, complete);
            double s = 0.0;
            for (int e = lo; e < hi; ++e) s += s_prod[e];
            if (it == 0) A.head_part[t] = s;
            else if (complete) A.q[A.row_lo + ra + it - 1] = s;
            else A.tail_part[t] = s;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void phase_spmv(const KRArgs &A, double *s_prod, int *s_ptr) {
    for (int64_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) spmv_tile(A, A.u, t, s_prod, s_ptr);
}

// A row that straddles tiles leaves tail_part in its first tile and head_part in the following
// ones; its value is their sum in tile order.  The straddling row of tile t, if any, is rb-1.
__device__ __forceinline__ void fix_tile(const KRArgs &A, int64_t t) {
    const int ra = A.tile_ra[t], rb = A.tile_ra[t + 1];
    if (rb <= ra) return;
    const int64_t rend = A.indptr[rb];
    if (rend <= (t + 1) * SPMV_TILE) return;
    const int64_t t_last = (rend - 1) / SPMV_TILE;
    double s = A.tail_part[t];
    for (int64_t t2 = t + 1; t2 <= t_last; ++t2) s += A.head_part[t2];
    A.q[A.row_lo + rb - 1] = s;
}

// stand-alone form (b3c_spmv): all tiles
__device__ __forceinline__ void phase_fix(const KRArgs &A) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < A.n_tiles; t += stride) fix_tile(A, t);
}

// fused form: the CTA that owns chunk c patches the straddling rows that fall inside the chunk
// before it reads q, so no separate pass (and no extra grid barrier) is needed
__device__ __forceinline__ void chunk_fixup(const KRArgs &A, int c) {
    const int lc = c - A.row_lo / CHUNK;
    const int t0 = A.chunk_t[lc], t1 = A.chunk_t[lc + 1];
    for (int t = t0 + (int)threadIdx.x; t < t1; t += KR_THREADS) fix_tile(A, t);
    __syncthreads();
}

// ---- vector phases: one CTA per 1024-row chunk, 4 rows per thread -----------------------------------
#define KR_FOR_CHUNKS(c) for (int c = blockIdx.x; c < A.n_chunks; c += gridDim.x)
#define KR_ROW(c, i) ((int64_t)(c) * CHUNK + (i) * KR_THREADS + threadIdx.x)

__device__ __forceinline__ bool chunk_local(const KRArgs &A, int c) {
    const int64_t r0 = (int64_t)c * CHUNK;
    return r0 >= A.row_lo && r0 < A.row_hi;
}

__device__ __forceinline__ void phase_init(const KRArgs &A) {
    KR_FOR_CHUNKS(c) {
        if (!chunk_local(A, c)) continue;
#pragma unroll
        for (int i = 0; i < CHUNK_RPT; ++i) {
            const int64_t r = KR_ROW(c, i);
            if (r < A.row_hi) {
                A.x[r] = 1.0;
                A.u[r] = 1.0;
            }
        }
    }
}

// v = x * (A x), rk = 1 - v, partial rk.rk           (sparse_utils.py:136-139, 196-199)
__device__ __forceinline__ void phase_resid(const KRArgs &A, double *s_red) {
    KR_FOR_CHUNKS(c) {
        double acc = 0.0;
        const bool loc = chunk_local(A, c);
        if (loc) {
            chunk_fixup(A, c);
#pragma unroll
            for (int i = 0; i < CHUNK_RPT; ++i) {
                const int64_t r = KR_ROW(c, i);
                if (r < A.row_hi) {
                    const double xx = A.x[r];
                    double qq = A.q[r];
                    if (A.dfix[r] != 0.0) qq = __dadd_rn(qq, A.u[r]);          // zero diagonal counted as one (Q2)
                    const double vv = __dmul_rn(xx, qq);
                    const double rr = __dsub_rn(1.0, vv);
                    A.v[r] = vv;
                    A.rk[r] = rr;
                    acc = __dadd_rn(acc, __dmul_rn(rr, rr));
                }
            }
        }
        double r[1] = {acc};
        block_reduce<1, 0>(r, s_red);
        if (threadIdx.x == 0) A.part[PA * A.n_chunks + c] = loc ? r[0] : 0.0;
    }
}

// first CG step: Z = rk / v, p = Z, partial rk.Z (Q1); later steps: p = Z + beta p.  u = x * p
__device__ __forceinline__ void phase_dir(const KRArgs &A, bool first, double beta, double *ycur, double *s_red) {
    KR_FOR_CHUNKS(c) {
        double acc = 0.0;
        const bool loc = chunk_local(A, c);
        if (loc) {
#pragma unroll
            for (int i = 0; i < CHUNK_RPT; ++i) {
                const int64_t r = KR_ROW(c, i);
                if (r < A.row_hi) {
                    double pp;
                    if (first) {
                        const double rr = A.rk[r];
                        const double z = __ddiv_rn(rr, A.v[r]);             // sparse_utils.py:158
                        A.Z[r] = z;
                        pp = z;
                        acc = __dadd_rn(acc, __dmul_rn(rr, z));
                        ycur[r] = 1.0;                                      // y[:] = e (sparse_utils.py:150)
                    } else {
                        pp = __dadd_rn(A.Z[r], __dmul_rn(beta, A.p[r]));    // sparse_utils.py:163
                    }
                    A.p[r] = pp;
                    A.u[r] = __dmul_rn(A.x[r], pp);
                }
            }
        }
        if (first) {
            double r[1] = {acc};
            block_reduce<1, 0>(r, s_red);
            if (threadIdx.x == 0) A.part[PB * A.n_chunks + c] = loc ? r[0] : 0.0;
        } else if (threadIdx.x == 0) {
            A.part[PB * A.n_chunks + c] = 0.0;
        }
    }
}

// w = x * (A (x p)) + v * p, partial p.w              (sparse_utils.py:165-166)
__device__ __forceinline__ void phase_w(const KRArgs &A, double *s_red) {
    KR_FOR_CHUNKS(c) {
        double acc = 0.0;
        const bool loc = chunk_local(A, c);
        if (loc) {
            chunk_fixup(A, c);
#pragma unroll
            for (int i = 0; i < CHUNK_RPT; ++i) {
                const int64_t r = KR_ROW(c, i);
                if (r < A.row_hi) {
                    double qq = A.q[r];
                    if (A.dfix[r] != 0.0) qq = __dadd_rn(qq, A.u[r]);
                    const double pp = A.p[r];
                    const double ww = __dadd_rn(__dmul_rn(A.x[r], qq), __dmul_rn(A.v[r], pp));
                    A.w[r] = ww;
                    acc = __dadd_rn(acc, __dmul_rn(pp, ww));
                }
            }
        }
        double r[1] = {acc};
        block_reduce<1, 0>(r, s_red);
        if (threadIdx.x == 0) A.part[PA * A.n_chunks + c] = loc ? r[0] : 0.0;
    }
}

// ap = alpha p, ynew = y + ap, min/max and both clamp factors, and -- speculatively, used only if
// the step is accepted -- rk -= alpha w, Z = rk * v (Q1), partial rk.Z   (sparse_utils.py:167-190)
__device__ __forceinline__ void phase_step(const KRArgs &A, double alpha, double delta, double Delta,
                                           const double *ycur, double *ynew, double *s_red) {
    KR_FOR_CHUNKS(c) {
        double rho = 0.0, mn = INFINITY, nmx = INFINITY, g1 = INFINITY, g2 = INFINITY;
        const bool loc = chunk_local(A, c);
        if (loc) {
#pragma unroll
            for (int i = 0; i < CHUNK_RPT; ++i) {
                const int64_t r = KR_ROW(c, i);
                if (r < A.row_hi) {
                    const double ap = __dmul_rn(alpha, A.p[r]);
                    const double yy = ycur[r];
                    const double yn = __dadd_rn(yy, ap);
                    ynew[r] = yn;
                    mn = fmin(mn, yn);
                    nmx = fmin(nmx, -yn);
                    if (ap < 0.0) g1 = fmin(g1, __ddiv_rn(__dsub_rn(delta, yy), ap));      // :174-175
                    if (yn > Delta) g2 = fmin(g2, __ddiv_rn(__dsub_rn(Delta, yy), ap));    // :180-181
                    const double rr = __dsub_rn(A.rk[r], __dmul_rn(alpha, A.w[r]));       // :186
                    const double z = __dmul_rn(rr, A.v[r]);                               // :189
                    A.rk[r] = rr;
                    A.Z[r] = z;
                    rho = __dadd_rn(rho, __dmul_rn(rr, z));
                }
            }
        }
        double r[5] = {rho, mn, nmx, g1, g2};
        block_reduce<1, 4>(r, s_red);
        if (threadIdx.x == 0) {
            const int nc = A.n_chunks;
            A.part[PC * nc + c] = loc ? r[0] : 0.0;
            A.part[PMIN * nc + c] = r[1];
            A.part[PNEGMAX * nc + c] = r[2];
            A.part[PG1 * nc + c] = r[3];
            A.part[PG2 * nc + c] = r[4];
        }
    }
}

// x *= y (y possibly clamped: y + gamma * alpha p), u = x      (sparse_utils.py:176,182,195)
__device__ __forceinline__ void phase_update(const KRArgs &A, int ymode, double gamma, double alpha,
                                             const double *ycur) {
    KR_FOR_CHUNKS(c) {
        if (!chunk_local(A, c)) continue;
#pragma unroll
        for (int i = 0; i < CHUNK_RPT; ++i) {
            const int64_t r = KR_ROW(c, i);
            if (r < A.row_hi) {
                double yy = 1.0;
                if (ymode >= 1) yy = ycur[r];
                if (ymode == 2) yy = __dadd_rn(yy, __dmul_rn(gamma, __dmul_rn(alpha, A.p[r])));
                const double xx = __dmul_rn(A.x[r], yy);
                A.x[r] = xx;
                A.u[r] = xx;
            }
        }
    }
}

// ---- scalar logic (identical in every thread) -----------------------------------------------------
// after a residual phase: rho = rk.rk; first call initialises, later calls close an outer step
__device__ __forceinline__ void scalar_outer(KRScalars &S, double rho, bool first) {
    const double g = 0.9, etamax = 0.1;
    S.rho_km1 = rho;
    S.rout = rho;
    if (first) {
        S.rold = rho;
        return;
    }
    S.n_iter += S.k + 1;                                  // sparse_utils.py:201
    const double rat = S.rout / S.rold;
    S.rold = S.rout;
    const double res_norm = sqrt(S.rout);
    const double eta_o = S.eta;
    S.eta = g * rat;
    if (g * eta_o * eta_o > 0.1) S.eta = fmax(S.eta, g * eta_o * eta_o);
    S.eta = fmax(fmin(S.eta, etamax), S.stop_tol / res_norm);
}

// after a step phase: accept or clamp.  returns true when the inner loop must stop.
__device__ __forceinline__ bool scalar_decide(KRScalars &S, double ymin, double ymax, double g1, double g2,
                                              double rho_new) {
    if (ymin <= S.delta) {                                // sparse_utils.py:171-177
        S.gamma = (S.delta == 0.0) ? 0.0 : g1;
        S.ymode = 2;
        return true;
    }
    if (ymax >= S.Delta) {                                // sparse_utils.py:179-183
        if (isinf(g2)) S.status = B3C_ERR_TIE;            // no element above Delta: exact tie (Q13)
        S.gamma = isinf(g2) ? 0.0 : g2;
        S.ymode = 2;
        return true;
    }
    S.ymode = 1;                                          // y = ynew
    S.ysel ^= 1;
    S.rho_km2 = S.rho_km1;
    S.rho_km1 = rho_new;
    return false;
}

// ---- the persistent kernel ---------------------------------------------------------------------------
#define KR_PHASE(id, call)                                                     \
    do {                                                                       \
        const long long t0_ = clock64();                                       \
        call;                                                                  \
        const long long t1_ = clock64();                                       \
        grid.sync();                                                           \
        if (timing) {                                                          \
            const long long t2_ = clock64();                                   \
            A.timers->work[id] += t1_ - t0_;                                   \
            A.timers->sync[id] += t2_ - t1_;                                   \
        }                                                                      \
    } while (0)

__global__ void __launch_bounds__(KR_THREADS, KR_MIN_CTAS) k_kr_persistent(KRArgs A) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_prod[SPMV_TILE];
    __shared__ int s_ptr[SPTR_CAP + 1];
    __shared__ double s_red[KR_WARPS * RED_MAX];
    KRScalars S = *A.ctl;
    const int nc = A.n_chunks;
    double *ybuf[2] = {A.y0, A.y1};
    const bool timing = (blockIdx.x == 0 && threadIdx.x == 0);
    const long long t_begin = clock64();
    long long ts_ = 0;

    KR_PHASE(T_INIT, phase_init(A));
    KR_PHASE(T_SPMV, phase_spmv(A, s_prod, s_ptr));
    KR_PHASE(T_RESID, phase_resid(A, s_red));
    S.n_spmv = 1;
    {
        double r[1];
        const int ids[1] = {PA};
        reduce_parts<1, 0>(A.part, nc, ids, r, s_red);
        scalar_outer(S, r[0], true);
    }

    while (S.rout > S.rt && S.n_iter < S.max_iter) {      // sparse_utils.py:146
        S.outer += 1;
        S.k = 0;
        S.ymode = 0;
        S.inner_tol = fmax(S.rout * S.eta * S.eta, S.rt);
        while (S.rho_km1 > S.inner_tol) {                 // sparse_utils.py:154
            S.k += 1;
            const bool first = (S.k == 1);
            if (!first) S.beta = S.rho_km1 / S.rho_km2;
            double *ycur = ybuf[S.ysel], *ynew = ybuf[S.ysel ^ 1];
            KR_PHASE(T_DIR, phase_dir(A, first, S.beta, ycur, s_red));
            KR_PHASE(T_SPMV, phase_spmv(A, s_prod, s_ptr));
            KR_PHASE(T_W, phase_w(A, s_red));
            S.n_spmv += 1;
            ts_ = clock64();
            {
                double r[2];
                const int ids[2] = {PA, PB};
                reduce_parts<2, 0>(A.part, nc, ids, r, s_red);
                if (first) S.rho_km1 = r[1];              // rk.Z of the first step (sparse_utils.py:160)
                S.alpha = S.rho_km1 / r[0];               // rho / p.w (sparse_utils.py:166)
            }
            if (timing) A.timers->work[T_SCALAR] += clock64() - ts_;
            KR_PHASE(T_STEP, phase_step(A, S.alpha, S.delta, S.Delta, ycur, ynew, s_red));
            ts_ = clock64();
            double r[5];
            {
                const int ids[5] = {PC, PMIN, PNEGMAX, PG1, PG2};
                reduce_parts<1, 4>(A.part, nc, ids, r, s_red);
            }
            if (timing) A.timers->work[T_SCALAR] += clock64() - ts_;
            if (scalar_decide(S, r[1], -r[2], r[3], r[4], r[0])) break;
            if (S.k >= S.max_iter + 8) {                  // safety net: the reference's inner loop is unbounded
                S.status = B3C_ERR_NOCONV;
                break;
            }
        }
        if (S.status != 0) break;
        // with ymode 2 the step was not accepted: ycur still holds y; with ymode 1 ysel was flipped
        KR_PHASE(T_UPDATE, phase_update(A, S.ymode, S.gamma, S.alpha, ybuf[S.ysel]));
        KR_PHASE(T_SPMV, phase_spmv(A, s_prod, s_ptr));
        KR_PHASE(T_RESID, phase_resid(A, s_red));
        S.n_spmv += 1;
        double r[1];
        const int ids[1] = {PA};
        reduce_parts<1, 0>(A.part, nc, ids, r, s_red);
        scalar_outer(S, r[0], false);
    }
    if (timing) {
        *A.ctl = S;
        A.timers->total = clock64() - t_begin;
    }
}

// ---- stand-alone kernels (plan, microbench SpMV, host-driven phases) -----------------------------
__global__ void k_tile_plan(int32_t n_local, const int64_t *__restrict__ indptr, int64_t n_tiles,
                            int32_t *__restrict__ tile_ra) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == n_tiles) {
        tile_ra[t] = n_local;
        return;
    }
    const int64_t base = t * SPMV_TILE;
    int lo = 0, hi = n_local;                 // first r in [0, n_local] with indptr[r] >= base
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (indptr[mid] >= base) hi = mid;
        else lo = mid + 1;
    }
    tile_ra[t] = lo;
}

// dfix[r] = 1 where the diagonal entry of (global) row r is absent or zero (sparse_utils.py:110-115)
__global__ void __launch_bounds__(KR_THREADS) k_diag_fix(int32_t row_lo, int32_t row_hi,
                                                         const int64_t *__restrict__ indptr,
                                                         const int32_t *__restrict__ indices,
                                                         const double *__restrict__ data, double *__restrict__ dfix,
                                                         KRScalars *ctl) {
    const unsigned lane = lane_id();
    const int64_t nw = (int64_t)gridDim.x * KR_WARPS;
    unsigned nz = 0;
    for (int64_t lr = (int64_t)blockIdx.x * KR_WARPS + (threadIdx.x >> 5); lr < row_hi - row_lo; lr += nw) {
        const int64_t lo = indptr[lr], hi = indptr[lr + 1];
        const int32_t gr = row_lo + (int32_t)lr;
        double d = 0.0;           // duplicates of the diagonal would be summed by scipy's diagonal()
        for (int64_t e = lo + lane; e < hi; e += 32)
            if (indices[e] == gr) d += data[e];
        d = warp_sum(d);
        if (lane == 0) {
            const bool z = (d == 0.0);
            dfix[gr] = z ? 1.0 : 0.0;
            nz += z ? 1u : 0u;
        }
    }
    if (lane == 0 && nz) atomicAdd((unsigned long long *)&ctl->zero_diag, (unsigned long long)nz);
}

// chunk_t[lc] = first tile t with tile_ra[t+1] > lc*CHUNK, i.e. whose last-starting row is at or
// beyond the chunk's first local row; chunk_t[n_local_chunks] = n_tiles
__global__ void k_chunk_plan(int32_t n_local_chunks, int64_t n_tiles, const int32_t *__restrict__ tile_ra,
                             int32_t *__restrict__ chunk_t) {
    const int lc = blockIdx.x * blockDim.x + threadIdx.x;
    if (lc > n_local_chunks) return;
    if (lc == n_local_chunks) {
        chunk_t[lc] = (int32_t)n_tiles;
        return;
    }
    const int64_t r0 = (int64_t)lc * CHUNK;
    int64_t lo = 0, hi = n_tiles;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)tile_ra[mid + 1] > r0) hi = mid;
        else lo = mid + 1;
    }
    chunk_t[lc] = (int32_t)lo;
}

__global__ void __launch_bounds__(KR_THREADS) k_spmv(KRArgs A) {
    __shared__ double s_prod[SPMV_TILE];
    __shared__ int s_ptr[SPTR_CAP + 1];
    phase_spmv(A, s_prod, s_ptr);
}
__global__ void __launch_bounds__(KR_THREADS) k_spmv_fix(KRArgs A) { phase_fix(A); }

// ---- phase-at-a-time form (multi-GPU row-block driver) ------------------------------------------
// The driver all-reduces u with SUM, so the slices a rank does not own must hold zeros.
__device__ __forceinline__ void zero_nonlocal_u(const KRArgs &A) {
    if (A.row_lo == 0 && A.row_hi == A.n) return;
    KR_FOR_CHUNKS(c) {
        if (chunk_local(A, c)) continue;
#pragma unroll
        for (int i = 0; i < CHUNK_RPT; ++i) {
            const int64_t r = KR_ROW(c, i);
            if (r < A.n) A.u[r] = 0.0;
        }
    }
}

enum { KRP_INIT = 0, KRP_SPMV, KRP_RESID, KRP_DIR, KRP_W, KRP_STEP, KRP_UPDATE };
enum { KRS_OUTER_FIRST = 0, KRS_OUTER, KRS_ALPHA, KRS_DECIDE };
enum { KR_STATE_DONE = 0, KR_STATE_INNER = 1, KR_STATE_UPDATE = 2 };

__global__ void __launch_bounds__(KR_THREADS) k_krp_phase(KRArgs A, int phase) {
    __shared__ double s_prod[SPMV_TILE];
    __shared__ int s_ptr[SPTR_CAP + 1];
    __shared__ double s_red[KR_WARPS * RED_MAX];
    const KRScalars S = *A.ctl;                 // only k_krp_scalar writes the control block
    double *ybuf[2] = {A.y0, A.y1};
    switch (phase) {
        case KRP_INIT:
            phase_init(A);
            zero_nonlocal_u(A);
            break;
        case KRP_SPMV:
            phase_spmv(A, s_prod, s_ptr);
            break;
        case KRP_RESID:
            phase_resid(A, s_red);
            break;
        case KRP_DIR:
            phase_dir(A, S.k == 1, S.beta, ybuf[S.ysel], s_red);
            zero_nonlocal_u(A);
            break;
        case KRP_W:
            phase_w(A, s_red);
            break;
        case KRP_STEP:
            phase_step(A, S.alpha, S.delta, S.Delta, ybuf[S.ysel], ybuf[S.ysel ^ 1], s_red);
            break;
        case KRP_UPDATE:
            phase_update(A, S.ymode, S.gamma, S.alpha, ybuf[S.ysel]);
            zero_nonlocal_u(A);
            break;
        default:
            break;
    }
}

// the loop control of k_kr_persistent, one decision at a time, on the (all-reduced) partials
__global__ void __launch_bounds__(KR_THREADS) k_krp_scalar(KRArgs A, int which) {
    __shared__ double s_red[KR_WARPS * RED_MAX];
    KRScalars S = *A.ctl;
    const int nc = A.n_chunks;
    if (which == KRS_OUTER_FIRST || which == KRS_OUTER) {
        double r[1];
        const int ids[1] = {PA};
        reduce_parts<1, 0>(A.part, nc, ids, r, s_red);
        scalar_outer(S, r[0], which == KRS_OUTER_FIRST);
        S.n_spmv += 1;
        if (S.rout > S.rt && S.n_iter < S.max_iter) {          // sparse_utils.py:146
            S.outer += 1;
            S.k = 0;
            S.ymode = 0;
            S.inner_tol = fmax(S.rout * S.eta * S.eta, S.rt);
            if (S.rho_km1 > S.inner_tol) {                     // sparse_utils.py:154
                S.k = 1;
                S.state = KR_STATE_INNER;
            } else {
                S.state = KR_STATE_UPDATE;
            }
        } else {
            S.state = KR_STATE_DONE;
        }
    } else if (which == KRS_ALPHA) {
        double r[2];
        const int ids[2] = {PA, PB};
        reduce_parts<2, 0>(A.part, nc, ids, r, s_red);
        if (S.k == 1) S.rho_km1 = r[1];
        S.alpha = S.rho_km1 / r[0];
        S.n_spmv += 1;
    } else if (which == KRS_DECIDE) {
        double r[5];
        const int ids[5] = {PC, PMIN, PNEGMAX, PG1, PG2};
        reduce_parts<1, 4>(A.part, nc, ids, r, s_red);
        bool stop = scalar_decide(S, r[1], -r[2], r[3], r[4], r[0]);
        if (!stop && S.k >= S.max_iter + 8) {
            S.status = B3C_ERR_NOCONV;
            stop = true;
        }
        if (S.status != 0) {
            S.state = KR_STATE_DONE;
        } else if (stop) {
            S.state = KR_STATE_UPDATE;
        } else if (S.rho_km1 > S.inner_tol) {
            S.k += 1;
            S.beta = S.rho_km1 / S.rho_km2;
            S.state = KR_STATE_INNER;
        } else {
            S.state = KR_STATE_UPDATE;
        }
    }
    if (threadIdx.x == 0) *A.ctl = S;
}

static std::mutex g_krp_mu;
static std::unordered_map<void *, KRArgs> g_krp;

// ---- workspace ---------------------------------------------------------------------------------------
struct KRLayout {
    int64_t n_tiles, o_tile_ra, o_chunk_t, o_head, o_tail, o_dfix, o_vec, o_part, o_ctl, o_timers, total;
    int32_t n_chunks;
};
static KRLayout kr_layout(int32_t n, int64_t nnz) {
    KRLayout L;
    Carver c;
    L.n_tiles = ceil_div(nnz, SPMV_TILE);
    if (L.n_tiles < 1) L.n_tiles = 1;
    L.n_chunks = (int32_t)ceil_div(n, CHUNK);
    L.o_tile_ra = c.take((L.n_tiles + 1) * 4);
    L.o_chunk_t = c.take(((int64_t)L.n_chunks + 2) * 4);
    L.o_head = c.take(L.n_tiles * 8);
    L.o_tail = c.take(L.n_tiles * 8);
    L.o_dfix = c.take((int64_t)n * 8);
    L.o_vec = c.take((int64_t)n * 8 * 10);
    L.o_part = c.take((int64_t)L.n_chunks * 8 * P_COUNT);
    L.o_ctl = c.take(sizeof(KRScalars));
    L.o_timers = c.take(sizeof(KRTimers));
    L.total = c.cur;
    return L;
}

static void kr_bind(KRArgs &A, const KRLayout &L, char *ws, int32_t n, int32_t row_lo, int32_t row_hi, int64_t nnz,
                    const int64_t *indptr, const int32_t *indices, const double *data) {
    A.n = n;
    A.row_lo = row_lo;
    A.row_hi = row_hi;
    A.nnz = nnz;
    A.indptr = indptr;
    A.indices = indices;
    A.data = data;
    A.n_tiles = L.n_tiles;
    A.tile_ra = (int32_t *)(ws + L.o_tile_ra);
    A.chunk_t = (int32_t *)(ws + L.o_chunk_t);
    A.head_part = (double *)(ws + L.o_head);
    A.tail_part = (double *)(ws + L.o_tail);
    A.dfix = (double *)(ws + L.o_dfix);
    double *vec = (double *)(ws + L.o_vec);
    A.x = vec;
    A.v = vec + (int64_t)n * 1;
    A.rk = vec + (int64_t)n * 2;
    A.y0 = vec + (int64_t)n * 3;
    A.y1 = vec + (int64_t)n * 4;
    A.p = vec + (int64_t)n * 5;
    A.Z = vec + (int64_t)n * 6;
    A.w = vec + (int64_t)n * 7;
    A.u = vec + (int64_t)n * 8;
    A.q = vec + (int64_t)n * 9;
    A.part = (double *)(ws + L.o_part);
    A.n_chunks = L.n_chunks;
    A.ctl = (KRScalars *)(ws + L.o_ctl);
    A.timers = (KRTimers *)(ws + L.o_timers);
}

static int persistent_grid(int *grid_out) {
    static int cached = 0;
    if (!cached) {
        int per_sm = 0, dev = 0, sms = 0;
        B3C_CUDA(cudaGetDevice(&dev));
        B3C_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        B3C_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_kr_persistent, KR_THREADS, 0));
        if (per_sm < 1) {
            set_error("persistent KR kernel does not fit on an SM");
            return B3C_ERR_CUDA;
        }
        if (per_sm > KR_MIN_CTAS) per_sm = KR_MIN_CTAS;
        cached = sms * per_sm;
    }
    *grid_out = cached;
    return B3C_OK;
}

static unsigned spmv_grid(int64_t n_tiles) {
    const int64_t cap = (int64_t)kNumSMs * 8;
    return (unsigned)(n_tiles < cap ? n_tiles : cap);
}

}  // namespace b3c

using namespace b3c;

extern "C" {

int64_t b3c_kr_workspace_bytes(int32_t n, int64_t nnz) {
    if (n <= 0 || nnz < 0) return B3C_ERR_ARG;
    return kr_layout(n, nnz).total;
}

int b3c_kr_run(int32_t n, int64_t nnz, const int64_t *d_indptr, const int32_t *d_indices, const double *d_data,
               double tol, double delta, double Delta, int32_t max_iter, int32_t mode, double *d_x, void *d_ws,
               int64_t ws_bytes, int64_t *h_info, void *stream) {
    B3C_REQUIRE(n > 0 && nnz >= 0 && d_indptr && d_x && d_ws && h_info, "bad arguments");
    B3C_REQUIRE(nnz == 0 || (d_indices && d_data), "null matrix arrays");
    B3C_REQUIRE(mode == 0, "b3c_kr_run: only mode 0 (persistent kernel) is implemented; use b3c_krp_* for phases");
    const KRLayout L = kr_layout(n, nnz);
    if (ws_bytes < L.total) {
        set_error("KR workspace too small: %lld < %lld", (long long)ws_bytes, (long long)L.total);
        return B3C_ERR_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = (char *)d_ws;
    KRArgs A;
    kr_bind(A, L, ws, n, 0, n, nnz, d_indptr, d_indices, d_data);

    KRScalars S;
    memset(&S, 0, sizeof(S));
    S.tol = tol;
    S.delta = delta;
    S.Delta = Delta;
    S.rt = tol * tol;                 // sparse_utils.py:135
    S.stop_tol = tol * 0.5;           // sparse_utils.py:131
    S.eta = 0.1;                      // etamax (sparse_utils.py:129-130)
    S.max_iter = max_iter;
    B3C_CUDA(cudaMemcpyAsync(A.ctl, &S, sizeof(S), cudaMemcpyHostToDevice, s));
    B3C_CUDA(cudaMemsetAsync(A.timers, 0, sizeof(KRTimers), s));

    k_tile_plan<<<(unsigned)ceil_div(L.n_tiles + 1, 256), 256, 0, s>>>(n, d_indptr, L.n_tiles, A.tile_ra);
    B3C_LAUNCH_CHECK();
    k_chunk_plan<<<(unsigned)ceil_div(L.n_chunks + 1, 256), 256, 0, s>>>(L.n_chunks, L.n_tiles, A.tile_ra, A.chunk_t);
    B3C_LAUNCH_CHECK();
    {
        int64_t blocks = ceil_div(n, KR_WARPS);
        if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
        k_diag_fix<<<(unsigned)blocks, KR_THREADS, 0, s>>>(0, n, d_indptr, d_indices, d_data, A.dfix, A.ctl);
        B3C_LAUNCH_CHECK();
    }
    int grid = 0;
    int rc = persistent_grid(&grid);
    if (rc) return rc;
    void *args[] = {&A};
    B3C_CUDA(cudaLaunchCooperativeKernel((void *)k_kr_persistent, dim3(grid), dim3(KR_THREADS), args, 0, s));
    count_launch();
    B3C_CUDA(cudaMemcpyAsync(d_x, A.x, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    B3C_CUDA(cudaMemcpyAsync(&S, A.ctl, sizeof(S), cudaMemcpyDeviceToHost, s));
    KRTimers T;
    B3C_CUDA(cudaMemcpyAsync(&T, A.timers, sizeof(T), cudaMemcpyDeviceToHost, s));
    B3C_CUDA(cudaStreamSynchronize(s));
    h_info[0] = S.n_iter;
    h_info[1] = S.zero_diag;
    h_info[2] = S.outer;
    h_info[3] = S.n_spmv;
    h_info[4] = grid;
    h_info[5] = T.total;
    for (int i = 0; i < T_COUNT; ++i) {
        h_info[6 + i] = T.work[i];
        h_info[6 + T_COUNT + i] = T.sync[i];
    }
    if (S.status == B3C_ERR_TIE) {
        set_error("KR: max(ynew) == Delta with no element above Delta (reference raises ValueError here)");
        return B3C_ERR_TIE;
    }
    if (S.status == B3C_ERR_NOCONV || S.n_iter > max_iter) {
        set_error("matrix balancing failed to converge in %lld iterations", (long long)S.n_iter);
        return B3C_ERR_NOCONV;
    }
    return B3C_OK;
}

int b3c_spmv(int32_t n, int64_t nnz, const int64_t *d_indptr, const int32_t *d_indices, const double *d_data,
             const double *d_u, double *d_y, void *d_ws, int64_t ws_bytes, int32_t prepared, void *stream) {
    B3C_REQUIRE(n > 0 && nnz >= 0 && d_indptr && d_u && d_y && d_ws, "bad arguments");
    const KRLayout L = kr_layout(n, nnz);
    if (ws_bytes < L.total) {
        set_error("SpMV workspace too small: %lld < %lld", (long long)ws_bytes, (long long)L.total);
        return B3C_ERR_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    KRArgs A;
    kr_bind(A, L, (char *)d_ws, n, 0, n, nnz, d_indptr, d_indices, d_data);
    A.u = const_cast<double *>(d_u);
    A.q = d_y;
    if (!prepared) {
        k_tile_plan<<<(unsigned)ceil_div(L.n_tiles + 1, 256), 256, 0, s>>>(n, d_indptr, L.n_tiles, A.tile_ra);
        B3C_LAUNCH_CHECK();
    }
    k_spmv<<<spmv_grid(L.n_tiles), KR_THREADS, 0, s>>>(A);
    B3C_LAUNCH_CHECK();
    k_spmv_fix<<<(unsigned)ceil_div(L.n_tiles, 256), 256, 0, s>>>(A);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

// ---- phase API (multi-GPU row-block driver, bin3c_b200/dist.py) --------------------------------------

int64_t b3c_krp_workspace_bytes(int32_t n, int64_t nnz_local) {
    if (n <= 0 || nnz_local < 0) return B3C_ERR_ARG;
    return kr_layout(n, nnz_local).total;
}

int b3c_krp_setup(int32_t n, int32_t row_lo, int32_t row_hi, int64_t nnz_local, const int64_t *d_indptr,
                  const int32_t *d_indices, const double *d_data, double tol, double delta, double Delta,
                  int32_t max_iter, void *d_ws, int64_t ws_bytes, int64_t *h_offsets, void *stream) {
    B3C_REQUIRE(n > 0 && 0 <= row_lo && row_lo < row_hi && row_hi <= n, "bad row block [%d,%d) of %d", row_lo, row_hi, n);
    B3C_REQUIRE(row_lo % CHUNK == 0 && (row_hi % CHUNK == 0 || row_hi == n), "row blocks must be %d-row aligned", CHUNK);
    B3C_REQUIRE(d_indptr && d_ws && h_offsets && nnz_local >= 0, "bad arguments");
    const KRLayout L = kr_layout(n, nnz_local);
    if (ws_bytes < L.total) {
        set_error("KR workspace too small: %lld < %lld", (long long)ws_bytes, (long long)L.total);
        return B3C_ERR_CAPACITY;
    }
    cudaStream_t s = (cudaStream_t)stream;
    KRArgs A;
    kr_bind(A, L, (char *)d_ws, n, row_lo, row_hi, nnz_local, d_indptr, d_indices, d_data);
    KRScalars S;
    memset(&S, 0, sizeof(S));
    S.tol = tol;
    S.delta = delta;
    S.Delta = Delta;
    S.rt = tol * tol;
    S.stop_tol = tol * 0.5;
    S.eta = 0.1;
    S.max_iter = max_iter;
    S.state = KR_STATE_INNER;
    B3C_CUDA(cudaMemcpyAsync(A.ctl, &S, sizeof(S), cudaMemcpyHostToDevice, s));
    B3C_CUDA(cudaMemsetAsync(A.timers, 0, sizeof(KRTimers), s));
    B3C_CUDA(cudaMemsetAsync(A.part, 0, (size_t)L.n_chunks * 8 * P_COUNT, s));
    const int32_t n_local = row_hi - row_lo;
    const int32_t n_local_chunks = (int32_t)ceil_div(n_local, CHUNK);
    k_tile_plan<<<(unsigned)ceil_div(L.n_tiles + 1, 256), 256, 0, s>>>(n_local, d_indptr, L.n_tiles, A.tile_ra);
    B3C_LAUNCH_CHECK();
    k_chunk_plan<<<(unsigned)ceil_div(n_local_chunks + 1, 256), 256, 0, s>>>(n_local_chunks, L.n_tiles, A.tile_ra,
                                                                            A.chunk_t);
    B3C_LAUNCH_CHECK();
    {
        int64_t blocks = ceil_div(n_local, KR_WARPS);
        if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
        k_diag_fix<<<(unsigned)blocks, KR_THREADS, 0, s>>>(row_lo, row_hi, d_indptr, d_indices, d_data, A.dfix, A.ctl);
        B3C_LAUNCH_CHECK();
    }
    h_offsets[0] = L.o_vec + (int64_t)n * 8 * 8;        // u: float64[n]
    h_offsets[1] = L.o_vec;                             // x: float64[n]
    h_offsets[2] = L.o_part;                            // partials: float64[7][n_chunks]
    h_offsets[3] = L.n_chunks;
    h_offsets[4] = L.o_ctl;
    std::lock_guard<std::mutex> g(g_krp_mu);
    g_krp[d_ws] = A;
    return B3C_OK;
}

static int krp_get(void *ws, KRArgs *A) {
    std::lock_guard<std::mutex> g(g_krp_mu);
    auto it = g_krp.find(ws);
    if (it == g_krp.end()) {
        set_error("workspace %p has no KR plan: call b3c_krp_setup first", ws);
        return B3C_ERR_ARG;
    }
    *A = it->second;
    return B3C_OK;
}

int b3c_krp_phase(void *d_ws, int32_t phase, void *stream) {
    KRArgs A;
    int rc = krp_get(d_ws, &A);
    if (rc) return rc;
    B3C_REQUIRE(phase >= KRP_INIT && phase <= KRP_UPDATE, "unknown phase %d", phase);
    unsigned grid;
    if (phase == KRP_SPMV) grid = spmv_grid(A.n_tiles);
    else grid = (unsigned)(A.n_chunks < kNumSMs * 4 ? A.n_chunks : kNumSMs * 4);
    k_krp_phase<<<grid, KR_THREADS, 0, (cudaStream_t)stream>>>(A, phase);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_krp_scalar(void *d_ws, int32_t which, void *stream) {
    KRArgs A;
    int rc = krp_get(d_ws, &A);
    if (rc) return rc;
    B3C_REQUIRE(which >= KRS_OUTER_FIRST && which <= KRS_DECIDE, "unknown scalar step %d", which);
    k_krp_scalar<<<1, KR_THREADS, 0, (cudaStream_t)stream>>>(A, which);
    B3C_LAUNCH_CHECK();
    return B3C_OK;
}

int b3c_krp_state(void *d_ws, int64_t *h_state, void *stream) {
    KRArgs A;
    int rc = krp_get(d_ws, &A);
    if (rc) return rc;
    B3C_REQUIRE(h_state != nullptr, "null h_state");
    KRScalars S;
    B3C_CUDA(cudaMemcpyAsync(&S, A.ctl, sizeof(S), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    B3C_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    h_state[0] = S.state;
    h_state[1] = S.status;
    h_state[2] = S.n_iter;
    h_state[3] = S.k;
    h_state[4] = S.outer;
    h_state[5] = S.n_spmv;
    h_state[6] = S.zero_diag;
    return B3C_OK;
}

}  // extern "C"
