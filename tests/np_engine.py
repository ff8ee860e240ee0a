"""
TEST INFRASTRUCTURE.  A NumPy stand-in for bin3c_b200.dist.CudaEngine so the HOST logic of the
multi-GPU driver (row splits, key routing plan, collectives between KR phases, loop control
hand-off) can run on CPU ranks with the gloo backend.  It follows the same phase contract as the
kernels (include/bin3c_b200.h, "Row-block phase API") but is not the product and is never
imported by it.
"""
import numpy as np
import scipy.sparse as sp
import torch

from bin3c_b200 import synth
from bin3c_b200.dist import (CHUNK, KRP_INIT, KRP_SPMV, KRP_RESID, KRP_DIR, KRP_W, KRP_STEP, KRP_UPDATE,
                             KRS_OUTER_FIRST, KRS_OUTER, KRS_ALPHA, KRS_DECIDE, STATE_DONE, STATE_INNER,
                             STATE_UPDATE, PA, PB, PC, PMIN, PNEGMAX, PG1, PG2)


class Block(object):
    def __init__(self, m, row_lo, n_total):
        self.m = m.tocsr()              # n_local x n_total
        self.row_lo = row_lo
        self.n = m.shape[0]
        self.n_total = n_total

    @property
    def nnz(self):
        return self.m.nnz


class NumpyEngine(object):

    def __init__(self, tid2idx, lengths, sites, pair_capacity):
        self.lut = np.asarray(tid2idx, dtype=np.int64)
        self.lengths = np.asarray(lengths)
        self.sites = np.asarray(sites)
        self.n = len(lengths)
        self.b = max(1, int(np.ceil(np.log2(max(self.n, 2)))))

    # ---- accumulation ---------------------------------------------------------------------------
    def classify(self, records):
        rec = records.numpy().view(np.uint64) if isinstance(records, torch.Tensor) else np.asarray(records, np.uint64)
        ti, tj, ok = synth.unpack_pairs(rec)
        inr = (ti < len(self.lut)) & (tj < len(self.lut))
        ii = np.where(inr, self.lut[np.minimum(ti, len(self.lut) - 1)], -1)
        jj = np.where(inr, self.lut[np.minimum(tj, len(self.lut) - 1)], -1)
        incl = (ii >= 0) & (jj >= 0)
        good = incl & ok
        self.counters = dict(accepted=int(good.sum()), ref_excluded=int((~incl).sum()),
                             poor_match=int((incl & ~ok).sum()))
        a, c = np.minimum(ii[good], jj[good]), np.maximum(ii[good], jj[good])
        d = a == c
        self.diag = torch.from_numpy(np.bincount(a[d], minlength=self.n).astype(np.int32))
        self.keys = (a[~d].astype(np.int64) << self.b) | c[~d].astype(np.int64)

    def row_hist(self):
        h = np.bincount(self.keys >> self.b, minlength=self.n) + np.bincount(self.keys & ((1 << self.b) - 1),
                                                                           minlength=self.n)
        return torch.from_numpy(h.astype(np.int64))

    def diag_counts(self):
        return self.diag

    def route(self, splits):
        i, j = self.keys >> self.b, self.keys & ((1 << self.b) - 1)
        directed = np.concatenate([(i << self.b) | j, (j << self.b) | i])
        owner = np.searchsorted(np.asarray(splits)[1:], np.concatenate([i, j]), side='right')
        order = np.argsort(owner, kind='stable')
        G = len(splits) - 1
        return torch.from_numpy(directed[order]), np.bincount(owner, minlength=G).tolist(), self.counters

    def build_block(self, keys, row_lo, row_hi):
        k = keys.numpy()
        u, cnt = np.unique(k, return_counts=True)
        r, c = u >> self.b, u & ((1 << self.b) - 1)
        assert r.size == 0 or (r.min() >= row_lo and r.max() < row_hi)
        d = self.diag.numpy()[row_lo:row_hi]
        rows = np.concatenate([r - row_lo, np.flatnonzero(d)])
        cols = np.concatenate([c, np.flatnonzero(d) + row_lo])
        vals = np.concatenate([cnt, d[d > 0]]).astype(np.uint32)
        m = sp.coo_matrix((vals, (rows, cols)), shape=(row_hi - row_lo, self.n)).tocsr()
        m.sort_indices()
        return Block(m, row_lo, self.n)

    # ---- mask / norm ------------------------------------------------------------------------------
    def block_mask(self, blk, min_len, min_sig):
        m = blk.m.tocoo()
        off = (m.row + blk.row_lo) != m.col
        sig = np.zeros(blk.n, dtype=np.uint32)
        np.maximum.at(sig, m.row[off], m.data[off])
        ln = self.lengths[blk.row_lo:blk.row_lo + blk.n]
        return torch.from_numpy(((ln >= min_len) & (sig >= min_sig)).astype(np.uint8))

    def new_mask(self):
        return torch.zeros(self.n, dtype=torch.uint8)

    def site_norm(self, blk):
        s = self.sites.astype(np.float64)
        s[s == 0] = 1
        m = blk.m.tocoo()
        data = m.data.astype(np.float64) * (1.0 / (s[m.row + blk.row_lo] * s[m.col]))
        return Block(sp.coo_matrix((data, (m.row, m.col)), shape=m.shape), blk.row_lo, self.n)

    # ---- KR phases ---------------------------------------------------------------------------------
    def kr_setup(self, blk, tol, delta, Delta, max_iter):
        n = self.n
        self.blk = blk
        self.lo, self.hi = blk.row_lo, blk.row_lo + blk.n
        a = blk.m.tocsr()
        dg = np.asarray(a[np.arange(blk.n), np.arange(self.lo, self.hi)]).ravel()
        self.dfix = (dg == 0)
        self.A = a
        self.nc = -(-n // CHUNK)
        self.u = torch.zeros(n, dtype=torch.float64)
        self.x = torch.zeros(n, dtype=torch.float64)
        self.part = torch.zeros(7, self.nc, dtype=torch.float64)
        z = lambda: np.zeros(n)
        self.v, self.rk, self.p, self.Z, self.w, self.q = z(), z(), z(), z(), z(), z()
        self.y = [z(), z()]
        self.S = dict(tol=tol, delta=delta, Delta=Delta, rt=tol ** 2, stop_tol=tol * 0.5, eta=0.1, max_iter=max_iter,
                      rho_km1=0.0, rho_km2=0.0, rout=0.0, rold=0.0, inner_tol=0.0, alpha=0.0, beta=0.0, gamma=0.0,
                      n_iter=0, k=0, outer=0, n_spmv=0, zero_diag=int(self.dfix.sum()), status=0, ymode=0, ysel=0,
                      state=STATE_INNER)

    def _chunks(self, vals, ident=0.0, op=np.add):
        out = np.full(self.nc, ident)
        for c in range(self.nc):
            lo, hi = max(c * CHUNK, self.lo), min((c + 1) * CHUNK, self.hi)
            if lo < hi:
                out[c] = op.reduce(vals[lo - self.lo:hi - self.lo]) if len(vals) else ident
        return out

    def _qq(self):
        q = self.q[self.lo:self.hi].copy()
        q[self.dfix] += self.u.numpy()[self.lo:self.hi][self.dfix]
        return q

    def kr_phase(self, ph):
        lo, hi, S = self.lo, self.hi, self.S
        x, u = self.x.numpy(), self.u.numpy()
        sl = slice(lo, hi)
        part = self.part.numpy()
        if ph == KRP_INIT:
            u[:] = 0
            x[sl] = 1.0
            u[sl] = 1.0
        elif ph == KRP_SPMV:
            self.q[sl] = self.A.dot(u)
        elif ph == KRP_RESID:
            self.v[sl] = x[sl] * self._qq()
            self.rk[sl] = 1 - self.v[sl]
            part[PA] = self._chunks(self.rk[sl] * self.rk[sl])
        elif ph == KRP_DIR:
            if S['k'] == 1:
                self.Z[sl] = self.rk[sl] / self.v[sl]
                self.p[sl] = self.Z[sl]
                part[PB] = self._chunks(self.rk[sl] * self.Z[sl])
                self.y[S['ysel']][sl] = 1.0
            else:
                self.p[sl] = self.Z[sl] + S['beta'] * self.p[sl]
                part[PB] = 0.0
            u[:] = 0
            u[sl] = x[sl] * self.p[sl]
        elif ph == KRP_W:
            self.w[sl] = x[sl] * self._qq() + self.v[sl] * self.p[sl]
            part[PA] = self._chunks(self.p[sl] * self.w[sl])
        elif ph == KRP_STEP:
            ycur, ynew = self.y[S['ysel']], self.y[S['ysel'] ^ 1]
            ap = S['alpha'] * self.p[sl]
            yn = ycur[sl] + ap
            ynew[sl] = yn
            with np.errstate(divide='ignore', invalid='ignore'):
                g1 = np.where(ap < 0, (S['delta'] - ycur[sl]) / ap, np.inf)
                g2 = np.where(yn > S['Delta'], (S['Delta'] - ycur[sl]) / ap, np.inf)
            self.rk[sl] = self.rk[sl] - S['alpha'] * self.w[sl]
            self.Z[sl] = self.rk[sl] * self.v[sl]
            part[PC] = self._chunks(self.rk[sl] * self.Z[sl])
            part[PMIN] = self._chunks(yn, np.inf, np.minimum)
            part[PNEGMAX] = self._chunks(-yn, np.inf, np.minimum)
            part[PG1] = self._chunks(g1, np.inf, np.minimum)
            part[PG2] = self._chunks(g2, np.inf, np.minimum)
        elif ph == KRP_UPDATE:
            yy = np.ones(hi - lo)
            if S['ymode'] >= 1:
                yy = self.y[S['ysel']][sl].copy()
            if S['ymode'] == 2:
                yy = yy + S['gamma'] * (S['alpha'] * self.p[sl])
            x[sl] = x[sl] * yy
            u[:] = 0
            u[sl] = x[sl]

    def kr_scalar(self, which):
        S = self.S
        part = self.part.numpy()
        if which in (KRS_OUTER_FIRST, KRS_OUTER):
            rho = part[PA].sum()
            S['rho_km1'] = S['rout'] = rho
            if which == KRS_OUTER_FIRST:
                S['rold'] = rho
            else:
                S['n_iter'] += S['k'] + 1
                rat = S['rout'] / S['rold']
                S['rold'] = S['rout']
                eta_o = S['eta']
                S['eta'] = 0.9 * rat
                if 0.9 * eta_o ** 2 > 0.1:
                    S['eta'] = max(S['eta'], 0.9 * eta_o ** 2)
                S['eta'] = max(min(S['eta'], 0.1), S['stop_tol'] / np.sqrt(S['rout']))
            S['n_spmv'] += 1
            if S['rout'] > S['rt'] and S['n_iter'] < S['max_iter']:
                S['outer'] += 1
                S['k'] = 0
                S['ymode'] = 0
                S['inner_tol'] = max(S['rout'] * S['eta'] ** 2, S['rt'])
                if S['rho_km1'] > S['inner_tol']:
                    S['k'] = 1
                    S['state'] = STATE_INNER
                else:
                    S['state'] = STATE_UPDATE
            else:
                S['state'] = STATE_DONE
        elif which == KRS_ALPHA:
            if S['k'] == 1:
                S['rho_km1'] = part[PB].sum()
            S['alpha'] = S['rho_km1'] / part[PA].sum()
            S['n_spmv'] += 1
        elif which == KRS_DECIDE:
            ymin, ymax = part[PMIN].min(), -part[PNEGMAX].min()
            g1, g2, rho_new = part[PG1].min(), part[PG2].min(), part[PC].sum()
            stop = True
            if ymin <= S['delta']:
                S['gamma'] = 0.0 if S['delta'] == 0 else g1
                S['ymode'] = 2
            elif ymax >= S['Delta']:
                if np.isinf(g2):
                    S['status'] = -6
                S['gamma'] = 0.0 if np.isinf(g2) else g2
                S['ymode'] = 2
            else:
                stop = False
                S['ymode'] = 1
                S['ysel'] ^= 1
                S['rho_km2'] = S['rho_km1']
                S['rho_km1'] = rho_new
            if S['status'] != 0:
                S['state'] = STATE_DONE
            elif stop:
                S['state'] = STATE_UPDATE
            elif S['rho_km1'] > S['inner_tol']:
                S['k'] += 1
                S['beta'] = S['rho_km1'] / S['rho_km2']
                S['state'] = STATE_INNER
            else:
                S['state'] = STATE_UPDATE

    def kr_state(self):
        return {k: self.S[k] for k in ('state', 'status', 'n_iter', 'k', 'outer', 'n_spmv', 'zero_diag')}

    # ---- scaling + edges ------------------------------------------------------------------------------
    def kr_apply(self, blk, x):
        xx = x.numpy()
        m = blk.m.tocoo()
        data = xx[m.row + blk.row_lo] * (m.data * xx[m.col])
        return Block(sp.coo_matrix((data, (m.row, m.col)), shape=m.shape), blk.row_lo, self.n)

    def compress_edges(self, blk, mask, reduce_max, scale=True):
        mk = mask.numpy().astype(bool)
        newidx = np.where(mk, np.cumsum(mk) - 1, -1)
        m = blk.m.tocsr()
        m.sort_indices()
        m = m.tocoo()
        gr = m.row + blk.row_lo
        keep = mk[gr] & mk[m.col]
        vmax = torch.tensor([m.data[keep].max() if keep.any() else 0.0], dtype=torch.float64)
        reduce_max(vmax)
        scl = 1.0 / float(vmax[0]) if scale else 1.0
        e = keep & (m.col >= gr)
        return dict(u=torch.from_numpy(newidx[gr[e]].astype(np.int32)), v=torch.from_numpy(newidx[m.col[e]].astype(np.int32)),
                    w=torch.from_numpy(m.data[e] * scl), scl=torch.tensor([scl], dtype=torch.float64),
                    n_edges=int(e.sum()), n_accepted=int(mk.sum()))

    def synchronize(self):
        pass
