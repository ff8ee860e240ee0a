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

// partial arrays, each n_chunks long
enum { PA = 0, PB, PC, PMIN, PNEGMAX, PG1, PG2, P_COUNT };

struct KRScalars {
    double tol, delta, Delta, rt, stop_tol;
    double rho_km1, rho_km2, rout, rold, eta, inner_tol, alpha, beta, gamma;
    long long n_iter, max_iter, k, outer, n_spmv, zero_diag;
    int status, ymode, ysel, state;
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
    double *head_part, *tail_part;
    double *dfix;
    // vectors (global row indexing, length n)
    double *x, *v, *rk, *y0, *y1, *p, *Z, *w, *u, *q;
    // partials [P_COUNT][n_chunks]
    double *part;
    int32_t n_chunks;
    KRScalars *ctl;
};

// ---- deterministic block reductions ---------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double *s_red) {
    v = warp_sum(v);
    const unsigned w = threadIdx.x >> 5;
    __syncthreads();
    if (lane_id() == 0) s_red[w] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < KR_WARPS; ++i) r += s_red[i];
    return r;
}
__device__ __forceinline__ double block_min(double v, double *s_red) {
    v = warp_min(v);
    const unsigned w = threadIdx.x >> 5;
    __syncthreads();
    if (lane_id() == 0) s_red[w] = v;
    __syncthreads();
    double r = s_red[0];
#pragma unroll
    for (int i = 1; i < KR_WARPS; ++i) r = fmin(r, s_red[i]);
    return r;
}

// every thread of every CTA gets the same value: fixed strided accumulation + fixed tree
__device__ __forceinline__ double reduce_sum(const double *part, int n, double *s_red) {
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += KR_THREADS) a += part[i];
    return block_sum(a, s_red);
}
__device__ __forceinline__ double reduce_min(const double *part, int n, double *s_red) {
    double a = INFINITY;
    for (int i = threadIdx.x; i < n; i += KR_THREADS) a = fmin(a, part[i]);
    return block_min(a, s_red);
}

// ---- SpMV -------------------------------------------------------------------------------------
// `u` is rewritten between SpMV phases of the same (persistent) launch, so it is read with
// ordinary coherent loads -- never ld.global.nc -- and carries no __restrict__.
__device__ __forceinline__ void spmv_tile(const KRArgs &A, const double *u, int64_t t, double *s_prod) {
    const int64_t base = t * SPMV_TILE;
    const int64_t rem = A.nnz - base;
    const int cnt = (int)(rem < SPMV_TILE ? (rem > 0 ? rem : 0) : SPMV_TILE);
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
#pragma unroll
    for (int k = 0; k < SPMV_NPT; ++k) {
        const int idx = k * KR_THREADS + threadIdx.x;
        if (idx < cnt) s_prod[idx] = a[k] * u[c[k]];
    }
    __syncthreads();
    const int ra = A.tile_ra[t], rb = A.tile_ra[t + 1];
    const int64_t end = base + cnt;
    const int n_items = (rb - ra) + 1;            // item 0 = head segment of a row begun earlier
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    if (n_items <= 8 * KR_WARPS) {
        for (int it = warp; it < n_items; it += KR_WARPS) {
            int64_t lo, hi, rend = 0;
            if (it == 0) {
                lo = base;
                hi = min(A.indptr[ra], end);
            } else {
                lo = A.indptr[ra + it - 1];
                rend = A.indptr[ra + it];
                hi = min(rend, end);
            }
            double s = 0.0;
            for (int64_t e = lo + lane; e < hi; e += 32) s += s_prod[(int)(e - base)];
            s = warp_sum(s);
            if (lane == 0) {
                if (it == 0) A.head_part[t] = s;
                else if (rend <= end) A.q[A.row_lo + ra + it - 1] = s;
                else A.tail_part[t] = s;
            }
        }
    } else {
        for (int it = threadIdx.x; it < n_items; it += KR_THREADS) {
            int64_t lo, hi, rend = 0;
            if (it == 0) {
                lo = base;
                hi = min(A.indptr[ra], end);
            } else {
                lo = A.indptr[ra + it - 1];
                rend = A.indptr[ra + it];
                hi = min(rend, end);
            }
            double s = 0.0;
            for (int64_t e = lo; e < hi; ++e) s += s_prod[(int)(e - base)];
            if (it == 0) A.head_part[t] = s;
            else if (rend <= end) A.q[A.row_lo + ra + it - 1] = s;
            else A.tail_part[t] = s;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void phase_spmv(const KRArgs &A, double *s_prod) {
    for (int64_t t = blockIdx.x; t < A.n_tiles; t += gridDim.x) spmv_tile(A, A.u, t, s_prod);
}

// rows that straddle tiles: tail of the first tile + heads of the following ones, in tile order
__device__ __forceinline__ void phase_fix(const KRArgs &A) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < A.n_tiles; t += stride) {
        const int ra = A.tile_ra[t], rb = A.tile_ra[t + 1];
        if (rb <= ra) continue;
        const int64_t rend = A.indptr[rb];            // end of the last row that starts in this tile
        if (rend <= (t + 1) * SPMV_TILE) continue;
        const int64_t t_last = (rend - 1) / SPMV_TILE;
        double s = A.tail_part[t];
        for (int64_t t2 = t + 1; t2 <= t_last; ++t2) s += A.head_part[t2];
        A.q[A.row_lo + rb - 1] = s;
    }
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
        acc = block_sum(acc, s_red);
        if (threadIdx.x == 0) A.part[PA * A.n_chunks + c] = loc ? acc : 0.0;
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
            acc = block_sum(acc, s_red);
            if (threadIdx.x == 0) A.part[PB * A.n_chunks + c] = loc ? acc : 0.0;
        }
    }
}

// w = x * (A (x p)) + v * p, partial p.w              (sparse_utils.py:165-166)
__device__ __forceinline__ void phase_w(const KRArgs &A, double *s_red) {
    KR_FOR_CHUNKS(c) {
        double acc = 0.0;
        const bool loc = chunk_local(A, c);
        if (loc) {
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
        acc = block_sum(acc, s_red);
        if (threadIdx.x == 0) A.part[PA * A.n_chunks + c] = loc ? acc : 0.0;
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
        rho = block_sum(rho, s_red);
        mn = block_min(mn, s_red);
        nmx = block_min(nmx, s_red);
        g1 = block_min(g1, s_red);
        g2 = block_min(g2, s_red);
        if (threadIdx.x == 0) {
            const int nc = A.n_chunks;
            A.part[PC * nc + c] = loc ? rho : 0.0;
            A.part[PMIN * nc + c] = mn;
            A.part[PNEGMAX * nc + c] = nmx;
            A.part[PG1 * nc + c] = g1;
            A.part[PG2 * nc + c] = g2;
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
__global__ void __launch_bounds__(KR_THREADS) k_kr_persistent(KRArgs A) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_prod[SPMV_TILE];
    __shared__ double s_red[KR_WARPS];
    KRScalars S = *A.ctl;
    const int nc = A.n_chunks;
    double *ybuf[2] = {A.y0, A.y1};

    phase_init(A);
    grid.sync();
    phase_spmv(A, s_prod);
    grid.sync();
    phase_fix(A);
    grid.sync();
    phase_resid(A, s_red);
    grid.sync();
    S.n_spmv = 1;
    scalar_outer(S, reduce_sum(A.part + PA * nc, nc, s_red), true);

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
            phase_dir(A, first, S.beta, ycur, s_red);
            grid.sync();
            phase_spmv(A, s_prod);
            grid.sync();
            phase_fix(A);
            grid.sync();
            phase_w(A, s_red);
            grid.sync();
            S.n_spmv += 1;
            if (first) S.rho_km1 = reduce_sum(A.part + PB * nc, nc, s_red);
            const double pw = reduce_sum(A.part + PA * nc, nc, s_red);
            S.alpha = S.rho_km1 / pw;
            phase_step(A, S.alpha, S.delta, S.Delta, ycur, ynew, s_red);
            grid.sync();
            const double ymin = reduce_min(A.part + PMIN * nc, nc, s_red);
            const double ymax = -reduce_min(A.part + PNEGMAX * nc, nc, s_red);
            const double g1 = reduce_min(A.part + PG1 * nc, nc, s_red);
            const double g2 = reduce_min(A.part + PG2 * nc, nc, s_red);
            const double rho_new = reduce_sum(A.part + PC * nc, nc, s_red);
            if (scalar_decide(S, ymin, ymax, g1, g2, rho_new)) break;
            if (S.k >= S.max_iter + 8) {                  // safety net: the reference's inner loop is unbounded
                S.status = B3C_ERR_NOCONV;
                break;
            }
        }
        if (S.status != 0) break;
        // with ymode 2 the step was not accepted: ycur still holds y; with ymode 1 ysel was flipped
        phase_update(A, S.ymode, S.gamma, S.alpha, ybuf[S.ysel]);
        grid.sync();
        phase_spmv(A, s_prod);
        grid.sync();
        phase_fix(A);
        grid.sync();
        phase_resid(A, s_red);
        grid.sync();
        S.n_spmv += 1;
        scalar_outer(S, reduce_sum(A.part + PA * nc, nc, s_red), false);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *A.ctl = S;
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

__global__ void __launch_bounds__(KR_THREADS) k_spmv(KRArgs A) {
    __shared__ double s_prod[SPMV_TILE];
    phase_spmv(A, s_prod);
}
__global__ void __launch_bounds__(KR_THREADS) k_spmv_fix(KRArgs A) { phase_fix(A); }

// ---- workspace ---------------------------------------------------------------------------------------
struct KRLayout {
    int64_t n_tiles, o_tile_ra, o_head, o_tail, o_dfix, o_vec, o_part, o_ctl, total;
    int32_t n_chunks;
};
static KRLayout kr_layout(int32_t n, int64_t nnz) {
    KRLayout L;
    Carver c;
    L.n_tiles = ceil_div(nnz, SPMV_TILE);
    if (L.n_tiles < 1) L.n_tiles = 1;
    L.n_chunks = (int32_t)ceil_div(n, CHUNK);
    L.o_tile_ra = c.take((L.n_tiles + 1) * 4);
    L.o_head = c.take(L.n_tiles * 8);
    L.o_tail = c.take(L.n_tiles * 8);
    L.o_dfix = c.take((int64_t)n * 8);
    L.o_vec = c.take((int64_t)n * 8 * 10);
    L.o_part = c.take((int64_t)L.n_chunks * 8 * P_COUNT);
    L.o_ctl = c.take(sizeof(KRScalars));
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
        if (per_sm > 4) per_sm = 4;
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

    k_tile_plan<<<(unsigned)ceil_div(L.n_tiles + 1, 256), 256, 0, s>>>(n, d_indptr, L.n_tiles, A.tile_ra);
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
    B3C_CUDA(cudaStreamSynchronize(s));
    h_info[0] = S.n_iter;
    h_info[1] = S.zero_diag;
    h_info[2] = S.outer;
    h_info[3] = S.n_spmv;
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

}  // extern "C"
