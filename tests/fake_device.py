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
