"""TEST INFRASTRUCTURE ONLY -- a numpy stand-in for the three downstream entry points of the C ABI
(``ds_set_scores`` / ``ds_corr`` / ``ds_group_stats`` of ``scdrs_b200._native.Context``), so that the HOST side of
``scdrs_b200/downstream.py`` (cell alignment, group layout, sub-graph, quantile positions, p-values, frames) is
covered by the CPU test suite.  The product never imports this file."""
import numpy as np
from scipy.stats import rankdata


class NumpyDownstreamContext:
    def ds_set_scores(self, frame, cols=None):
        frame = np.asarray(frame)
        if cols is not None:
            frame = frame[:, cols]
        self.f32 = frame.dtype == np.float32
        self.S = np.asarray(frame, dtype=np.float64)
        self.ds_shape = frame.shape

    def ds_corr(self, variables):
        v = np.atleast_2d(variables)
        return np.array([[np.corrcoef(x, self.S[:, j])[0, 1] for j in range(self.S.shape[1])] for x in v])

    def ds_group_stats(self, order, grp_ptr, q_lo=None, q_gamma=None, geary_mode=0, graph=None):
        n_group, ncol = len(grp_ptr) - 1, self.S.shape[1]
        quant = np.full((n_group, ncol), np.nan) if q_lo is not None else None
        total = denom = None
        M = self.S[order].copy()
        for g in range(n_group):
            a, b = grp_ptr[g], grp_ptr[g + 1]
            blk = M[a:b].copy()
            srt = np.sort(blk, axis=0)
            if quant is not None and b > a:
                lo = q_lo[g]
                hi = min(lo + 1, b - a - 1)
                d = srt[hi] - srt[lo]
                if self.f32:
                    d = d.astype(np.float32).astype(np.float64)
                quant[g] = srt[hi] - d * (1 - q_gamma[g]) if q_gamma[g] >= 0.5 else srt[lo] + d * q_gamma[g]
            if geary_mode == 1:
                for j in range(ncol):
                    M[a:b, j] = srt[:, 0][rankdata(blk[:, j], method="ordinal").astype(np.int64) - 1]
        if geary_mode:
            indptr, nbr, w = graph
            total, denom = np.zeros((n_group, ncol)), np.zeros((n_group, ncol))
            for g in range(n_group):
                a, b = grp_ptr[g], grp_ptr[g + 1]
                for p in range(a, b):
                    e = slice(indptr[p], indptr[p + 1])
                    total[g] += (w[e][:, None] * (M[p][None, :] - M[nbr[e]]) ** 2).sum(axis=0)
                if b > a:
                    denom[g] = ((M[a:b] - M[a:b].mean(axis=0)) ** 2).sum(axis=0)
        return quant, total, denom


class NumpyStatsContext:
    """numpy stand-in for the per-gene / per-cell reductions of the C ABI that ``scdrs_b200/pp.py`` calls
    (``gene_moments``, ``gene_expm1_implicit``, ``gene_stats_exact``, ``gene_cell_stats``): dense float64 arithmetic
    on the context's own cells, so that the HOST side of ``preprocess`` -- the algebra that turns moments into
    mean / variance, the covariate regression, and the sums over the ranks of a sharded atlas -- runs on CPU."""

    TILE = 32

    def __init__(self, X):
        from scipy import sparse

        self.A = np.asarray(X.toarray() if sparse.issparse(X) else X, dtype=np.float64)
        self.stored = self.A != 0
        self.n_cell, self.n_gene = self.A.shape
        self.C = self.B = self.mu = None

    def set_cov(self, cov_mat=None, cov_beta=None, gene_mean=None):
        self.C, self.B, self.mu = cov_mat, cov_beta, gene_mean

    def _w(self, cell_weight):
        return np.ones(self.n_cell) if cell_weight is None else np.asarray(cell_weight, dtype=np.float64)

    def gene_moments(self, mode, cov_mat=None, cell_weight=None):
        w = self._w(cell_weight)[:, None]
        if mode == 1:
            T = np.where(self.stored, np.expm1(self.A), 0.0)
            return np.stack([(w * T).sum(axis=0), (w * T * T).sum(axis=0)], axis=1)
        cols = [(w * self.A).sum(axis=0), (w * self.A * self.A).sum(axis=0)]
        if cov_mat is not None:
            cols += [(w * self.A * np.asarray(cov_mat)[:, [j]]).sum(axis=0) for j in range(np.asarray(cov_mat).shape[1])]
        return np.stack(cols, axis=1)

    def _corrected(self):
        return self.A + self.C @ self.B.T + self.mu if self.C is not None else self.A

    def gene_expm1_implicit(self, shift, cell_weight=None):
        w = self._w(cell_weight)[:, None]
        T = np.expm1(self._corrected()) - np.asarray(shift)[None, :]
        return np.stack([(w * T).sum(axis=0), (w * T * T).sum(axis=0)], axis=1)

    def gene_stats_exact(self, transform, genes, gene_sum, gene_sumsq, cell_weight=None, center=False, denom=0.0,
                         mean_in=None):
        tiles = np.unique(np.asarray(genes) // self.TILE)
        idx = (tiles[:, None] * self.TILE + np.arange(self.TILE)[None, :]).reshape(-1)
        idx = idx[idx < self.n_gene]
        V = self._corrected()[:, idx]
        V = np.expm1(V) if transform == 1 else V
        w = self._w(cell_weight)[:, None]
        s = (w * V).sum(axis=0)
        if center:
            m = s / denom if mean_in is None else np.asarray(mean_in)[idx]
            q = (w * (V - m) ** 2).sum(axis=0)
        else:
            q = (w * V * V).sum(axis=0)
        gene_sum[idx], gene_sumsq[idx] = s, q
        return idx

    def gene_cell_stats(self, transform, cell_weight=None, want_gene=True, want_cell=True, center=False, denom=0.0):
        V = self._corrected()
        V = np.expm1(V) if transform == 1 else V
        return None, None, V.sum(axis=1), (V * V).sum(axis=1)


class NumpyStatsState:
    """What ``scdrs_b200._device.get_state`` returns, backed by ``NumpyStatsContext``."""

    def __init__(self, adata):
        self.ctx = NumpyStatsContext(adata.X)

    def ensure_matrix(self, adata):
        return self.ctx

    def ensure_cov(self, adata, implicit):
        if not implicit:
            self.ctx.set_cov(None)
            return
        from scdrs_b200.pp import _cov_arrays

        self.ctx.set_cov(*_cov_arrays(adata))
