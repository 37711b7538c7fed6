"""Golden vectors for the downstream analyses, from the UNMODIFIED reference sources.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_downstream.py
Writes tests/golden/ref_downstream_case.npz: seeded score frames, cell groups, a synthetic symmetric
neighbour graph (scanpy's ``pp.neighbors`` is absent; any weighted graph exercises the same code) and what
the reference's downstream_group_analysis / downstream_corr_analysis / downstream_gene_analysis / gearys_c /
test_gearysc return for them (statsmodels' multipletests replaced by the Benjamini-Hochberg restatement of
oracle/downstream_oracle.py).
"""
import os
import sys
import warnings

import numpy as np
import pandas as pd
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
warnings.filterwarnings("ignore")

from oracle import downstream_oracle, ref_shim  # noqa: E402
from scdrs_b200.anndata_lite import AnnData  # noqa: E402


def make_case(n_cell=240, n_gene=60, n_ctrl=25, n_group=4, k=8, seed=5, dtype=np.float64):
    rng = np.random.default_rng(seed)
    X = sparse.random(n_cell, n_gene, density=0.3, random_state=seed, format="csr", dtype=np.float32)
    X.data = np.log1p(X.data * 5).astype(np.float32)
    names = ["c%04d" % i for i in rng.permutation(n_cell)]            # not sorted: alignment is exercised
    labels = rng.integers(0, n_group, n_cell)
    labels[:3] = n_group                                              # a tiny group
    obs = pd.DataFrame({"grp": ["g%d" % v for v in labels], "v1": rng.normal(size=n_cell), "v2": rng.gamma(2.0, size=n_cell)}, index=names)
    latent = rng.normal(size=n_cell)
    norm = (latent + 0.5 * obs["v1"].values + 0.3 * rng.normal(size=n_cell)).astype(dtype)
    ctrl = rng.normal(size=(n_cell, n_ctrl)).astype(dtype)
    ctrl[::7, 3] = ctrl[0, 3]                                         # ties inside a control column
    # symmetric weighted k-nearest-neighbour graph on (latent, v1)
    pts = np.stack([latent, obs["v1"].values], axis=1)
    d = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(axis=2)
    np.fill_diagonal(d, np.inf)
    nbr = np.argsort(d, axis=1)[:, :k]
    w = np.exp(-np.take_along_axis(d, nbr, axis=1))
    G = sparse.csr_matrix((w.reshape(-1), (np.repeat(np.arange(n_cell), k), nbr.reshape(-1))), shape=(n_cell, n_cell))
    G = ((G + G.T) / 2).tocsr().astype(np.float32)
    p = rng.uniform(size=n_cell) ** 3
    cols = {"norm_score": norm, "pval": p.astype(dtype)}
    for i in range(n_ctrl):
        cols["ctrl_norm_score_%d" % i] = ctrl[:, i]
    df = pd.DataFrame(cols, index=names)
    adata = AnnData(X, obs, pd.DataFrame(index=["gene%d" % i for i in range(n_gene)]), obsp={"connectivities": G})
    return adata, df


def main():
    pp, method = ref_shim.load()
    method.scdrs.pp = pp            # method.py reaches _get_mean_var through the package (method.py:1153)
    method.multipletests = lambda p, method="fdr_bh": (None, downstream_oracle.fdr_bh(p))
    # scipy >= 1.1x returns float ranks for method="ordinal"; the reference indexes with them (method.py:978),
    # which only older scipy allowed.  Same ranks, integer dtype:
    import scipy.stats

    _rankdata = scipy.stats.rankdata
    scipy.stats.rankdata = lambda v, method="average": _rankdata(v, method=method).astype(np.int64)
    adata, df = make_case()
    grp = method.downstream_group_analysis(adata, df, ["grp"])["grp"]
    corr = method.downstream_corr_analysis(adata, df, ["v1", "v2"])
    gene = method.downstream_gene_analysis(adata, df)
    order = sorted(adata.obs_names)
    sub = adata[order]
    rls = method.test_gearysc(sub, df.loc[order], "grp")
    rls_ctrl = method.test_gearysc(sub, df.loc[order], "grp", opt="control")
    gc = method.gearys_c(adata, df["norm_score"].values)
    G = adata.obsp["connectivities"]
    np.savez_compressed(
        os.path.join(HERE, "ref_downstream_case.npz"),
        X_data=adata.X.data, X_indices=adata.X.indices, X_indptr=adata.X.indptr, shape=np.array(adata.shape),
        obs_names=np.array(adata.obs_names, dtype="U"), var_names=np.array(adata.var_names, dtype="U"),
        grp=np.array(adata.obs["grp"], dtype="U"), v1=adata.obs["v1"].values, v2=adata.obs["v2"].values,
        G_data=G.data, G_indices=G.indices, G_indptr=G.indptr,
        score_columns=np.array(df.columns, dtype="U"), score_values=df.values,
        grp_index=np.array(grp.index, dtype="U"), grp_columns=np.array(grp.columns, dtype="U"), grp_values=grp.values.astype(np.float64),
        corr_values=corr.values.astype(np.float64),
        gene_index=np.array(gene.index, dtype="U"), gene_values=gene.values.astype(np.float64),
        rls_index=np.array(rls.index, dtype="U"), rls_columns=np.array(rls.columns, dtype="U"), rls_values=rls.values.astype(np.float64),
        rls_ctrl_values=rls_ctrl.loc[rls.index].values.astype(np.float64), gearys_c=np.float64(gc),
    )
    print(grp, corr, gene.head(), rls, gc, sep="\n")


if __name__ == "__main__":
    main()
