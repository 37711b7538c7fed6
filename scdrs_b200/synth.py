"""Seeded synthetic inputs of the shapes BASELINE.json names.

Host-side helpers shared by ``bench.py`` and the tests.  Expression is TMS-FACS-like: a
log-normalised-count CSR matrix (float32 values, int32 column ids, int64 row pointers) in
which gene g is expressed in a cell with probability ``q_g`` (wide range, mean ``density``)
and expressed values are ``log1p(Gamma(1, 3))``; covariates are ``const, n_genes, sex_male,
age``; traits are sets of distinct genes with unit or MAGMA-z-like weights.
"""
import numpy as np
import pandas as pd
from scipy import sparse

from .anndata_lite import AnnData


def gene_detection_prob(n_gene, density, rng):
    p = rng.gamma(0.5, 1.0, size=n_gene) + 1e-3
    q = p / p.mean() * density
    # squash so that no gene is expressed everywhere while keeping the mean density
    return np.minimum(q, 0.95)


def make_csr(n_cell, n_gene, density=0.10, seed=20261019, chunk=4096):
    """CSR (data f32, indices i32, indptr i64) built in row chunks of Bernoulli(q_g) draws."""
    rng = np.random.default_rng(seed)
    q = gene_detection_prob(n_gene, density, rng).astype(np.float32)
    datas, idxs, counts = [], [], []
    for start in range(0, n_cell, chunk):
        m = min(chunk, n_cell - start)
        mask = rng.random((m, n_gene), dtype=np.float32) < q
        r, c = np.nonzero(mask)
        counts.append(np.bincount(r, minlength=m))
        idxs.append(c.astype(np.int32))
        datas.append(np.log1p(rng.gamma(1.0, 3.0, size=c.shape[0])).astype(np.float32))
    indptr = np.zeros(n_cell + 1, dtype=np.int64)
    np.cumsum(np.concatenate(counts), out=indptr[1:])
    return sparse.csr_matrix(
        (np.concatenate(datas), np.concatenate(idxs), indptr), shape=(n_cell, n_gene)
    )


def make_covariates(X, seed=7):
    """const, n_genes, sex_male, age (experiments/job.compute_score/readme.md:5 of the reference)."""
    rng = np.random.default_rng(seed)
    n_cell = X.shape[0]
    n_genes = np.diff(X.indptr).astype(np.float64)
    return pd.DataFrame(
        {
            "const": np.ones(n_cell),
            "n_genes": n_genes,
            "sex_male": rng.integers(0, 2, size=n_cell).astype(np.float64),
            "age": rng.choice([3.0, 18.0, 24.0], size=n_cell),
        }
    )


def make_adata(n_cell, n_gene, density=0.10, seed=20261019, n_group=0):
    X = make_csr(n_cell, n_gene, density, seed)
    obs = pd.DataFrame(index=["cell_%d" % i for i in range(n_cell)])
    if n_group:
        rng = np.random.default_rng(seed + 1)
        size = 1.0 / np.arange(1, n_group + 1) ** 1.2
        size = np.maximum(size, 0.012 * size.max())  # groups stay >= 1 % of the largest
        obs["cell_type"] = rng.choice(
            ["ct_%03d" % i for i in range(n_group)], size=n_cell, p=size / size.sum()
        )
    var = pd.DataFrame(index=["gene_%05d" % i for i in range(n_gene)])
    adata = AnnData(X, obs, var)
    cov = make_covariates(X)
    cov.index = obs.index
    return adata, cov


def make_traits(var_names, n_trait=74, n_gene_per_trait=1000, weighted=False, seed=1000):
    """dict trait -> (gene_list, weight_list); weights 1 or clip(3 + Exp(1.2), max=10)."""
    out = {}
    var_names = np.asarray(var_names)
    for t in range(n_trait):
        rng = np.random.default_rng(seed + t)
        k = min(n_gene_per_trait, len(var_names))
        genes = var_names[rng.choice(len(var_names), size=k, replace=False)]
        if weighted:
            w = np.minimum(3.0 + rng.exponential(1.2, size=k), 10.0)
        else:
            w = np.ones(k)
        out["trait_%02d" % t] = (list(genes), list(w))
    return out
