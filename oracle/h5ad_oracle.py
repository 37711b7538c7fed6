"""TEST INFRASTRUCTURE ONLY -- host (numpy / scipy.sparse) restatement of the checks, filters and
normalisation that ``scdrs.util.load_h5ad`` applies to a raw count matrix (scdrs/util.py:72-91 of the
reference), used to check ``scdrs_b200.util.preprocess_counts`` (device path).  Only ``tests/`` may
import this module.

The reference calls scanpy (absent from this image, not under /root/reference; pinned as
``scanpy>=1.6.0``, setup.py:29): ``sc.pp.filter_cells(adata, min_genes)``, ``sc.pp.filter_genes(adata,
min_cells)``, ``sc.pp.normalize_per_cell(adata, counts_per_cell_after=1e4)``, ``sc.pp.log1p(adata)``.
Their published behaviour is restated here: cells with fewer than ``min_genes`` expressed genes are
removed (``obs["n_genes"]``), then genes expressed in fewer than ``min_cells`` remaining cells
(``var["n_cells"]``); ``normalize_per_cell`` removes cells whose total count is below 1
(``min_counts=1``), stores the totals in ``obs["n_counts"]`` and scales every cell to 1e4; ``log1p`` is the
natural logarithm.  No reference test pins these steps (the CLI tests run with both flags False,
tests/test_CLI.py:36-38): PARITY UNPINNED beyond this restatement.
"""
import numpy as np
from scipy import sparse


def preprocess_counts(X, flag_filter_data=False, flag_raw_count=True, min_genes=250, min_cells=50):
    """Returns (X_out csr float32, keep_cell bool[n_cell], keep_gene bool[n_gene], n_genes, n_cells, n_counts)
    with the per-cell / per-gene arrays indexed like the ORIGINAL matrix."""
    X = sparse.csr_matrix(X)
    if np.isnan(X.sum()):
        raise ValueError("NaN")
    if (X < 0).sum() > 0:
        raise ValueError("negative")
    n_cell, n_gene = X.shape
    keep_cell = np.ones(n_cell, bool)
    keep_gene = np.ones(n_gene, bool)
    n_genes = np.diff(X.indptr).astype(np.int64)
    n_cells = np.zeros(n_gene, np.int64)
    n_counts = np.zeros(n_cell, np.float32)
    if flag_filter_data:
        keep_cell = n_genes >= min_genes
        Xc = sparse.csc_matrix(X[keep_cell])
        n_cells = np.diff(Xc.indptr).astype(np.int64)
        keep_gene = n_cells >= min_cells
    Y = X[keep_cell][:, keep_gene]
    if flag_raw_count:
        counts = np.asarray(Y.sum(axis=1)).reshape(-1).astype(np.float32)
        n_counts[np.flatnonzero(keep_cell)] = counts
        ok = counts >= 1
        Y = sparse.csr_matrix(Y[ok], dtype=np.float32, copy=True)
        idx = np.flatnonzero(keep_cell)
        keep_cell = np.zeros(n_cell, bool)
        keep_cell[idx[ok]] = True
        scale = (np.float32(1e4) / counts[ok]).astype(np.float32)
        Y.data *= np.repeat(scale, np.diff(Y.indptr))
        Y.data = np.log1p(Y.data)
    return sparse.csr_matrix(Y, dtype=np.float32), keep_cell, keep_gene, n_genes, n_cells, n_counts
