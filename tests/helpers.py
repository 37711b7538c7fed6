"""Shared test helpers: golden-case loading and small comparisons."""
import glob
import os

import numpy as np
import pandas as pd
from scipy import sparse

from scdrs_b200.anndata_lite import AnnData, read_h5ad

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOY = os.path.join(GOLDEN, "toy")


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN, "ref_case_*.npz")))


def load_case(path, with_param=True):
    """Rebuild (adata, cov, case dict) of one tests/golden/ref_case_*.npz file."""
    z = np.load(path, allow_pickle=False)
    shape = tuple(int(s) for s in z["shape"])
    X = sparse.csr_matrix((z["X_data"], z["X_indices"], z["X_indptr"]), shape=shape)
    obs = pd.DataFrame(index=[str(s) for s in z["obs_names"]])
    if z["cell_type"].shape[0]:
        obs["cell_type"] = [str(s) for s in z["cell_type"]]
    var = pd.DataFrame(index=[str(s) for s in z["var_names"]])
    adata = AnnData(X, obs, var)
    cov = pd.DataFrame(z["cov_values"], index=obs.index, columns=[str(s) for s in z["cov_columns"]])
    use_cov = bool(z["use_cov"])
    if with_param:
        gs = pd.DataFrame(
            {
                "mean": z["gs_mean"], "var": z["gs_var"], "var_tech": z["gs_var_tech"],
                "ct_mean": z["gs_ct_mean"], "ct_var": z["gs_ct_var"], "ct_var_tech": z["gs_ct_var_tech"],
                "mean_var": [str(s) for s in z["gs_mean_var"]],
            },
            index=var.index,
        )
        param = {"FLAG_SPARSE": True, "FLAG_COV": use_cov, "FLAG_ADJ_PROP": bool(z["cell_type"].shape[0]),
                 "GENE_STATS": gs,
                 "CELL_STATS": pd.DataFrame({"mean": z["cs_mean"], "var": z["cs_var"]}, index=obs.index)}
        if use_cov:
            cols = [str(s) for s in z["cov_mat_columns"]]
            param["COV_MAT"] = pd.DataFrame(z["cov_mat"], index=obs.index, columns=cols)
            param["COV_BETA"] = pd.DataFrame(z["cov_beta"], index=var.index, columns=cols)
            param["COV_GENE_MEAN"] = pd.Series(z["cov_gene_mean"], index=var.index)
        adata.uns["SCDRS_PARAM"] = param
    case = dict(
        gene_list=[str(s) for s in z["gene_list"]], gene_weight=list(z["gene_weight"]),
        ctrl_match_key=str(z["ctrl_match_key"]), weight_opt=str(z["weight_opt"]), n_ctrl=int(z["n_ctrl"]),
        use_cov=use_cov, adj_prop="cell_type" if z["cell_type"].shape[0] else None,
        ref=pd.DataFrame(z["ref_values"], index=obs.index, columns=[str(s) for s in z["ref_columns"]]),
        ref_ctrl_idx=z["ref_ctrl_idx"], ref_ctrl_weight=z["ref_ctrl_weight"], z=z,
    )
    return adata, cov, case


def load_toy():
    """The reference's bundled toy data set and golden score files (tests/golden/toy)."""
    adata = read_h5ad(os.path.join(TOY, "toydata_mouse.h5ad"))
    df_cov = pd.read_csv(os.path.join(TOY, "toydata_mouse.cov"), sep="\t", index_col=0)
    df_gs = pd.read_csv(os.path.join(TOY, "toydata_mouse.gs"), sep="\t")
    df_gs.index = df_gs["TRAIT"]
    ref = {
        "REF_NOCOV": pd.read_csv(os.path.join(TOY, "toydata_gs_mouse.ref_Ctrl20_CovNone.score.gz"), sep="\t", index_col=0),
        "REF_COV": pd.read_csv(os.path.join(TOY, "toydata_gs_mouse.ref_Ctrl20_CovConstCovariate.score.gz"), sep="\t", index_col=0),
        "REF_COV_FULL": pd.read_csv(os.path.join(TOY, "toydata_gs_mouse.ref_Ctrl20_CovConstCovariate.full_score.gz"), sep="\t", index_col=0),
    }
    stats = {
        "nocov": pd.read_csv(os.path.join(TOY, "toydata_mouse.gene_stats.nocov.tsv"), sep="\t", index_col=0, dtype={"mean_var": str}),
        "cov": pd.read_csv(os.path.join(TOY, "toydata_mouse.gene_stats.cov.tsv"), sep="\t", index_col=0, dtype={"mean_var": str}),
    }
    return adata, df_cov, df_gs, ref, stats


def max_rel(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def load_downstream_case():
    """tests/golden/ref_downstream_case.npz (made by make_golden_downstream.py from the live reference):
    (adata with obsp["connectivities"], df_full_score, dict of the reference's outputs)."""
    z = np.load(os.path.join(GOLDEN, "ref_downstream_case.npz"), allow_pickle=False)
    shape = tuple(int(s) for s in z["shape"])
    X = sparse.csr_matrix((z["X_data"], z["X_indices"], z["X_indptr"]), shape=shape)
    names = [str(s) for s in z["obs_names"]]
    obs = pd.DataFrame({"grp": [str(s) for s in z["grp"]], "v1": z["v1"], "v2": z["v2"]}, index=names)
    G = sparse.csr_matrix((z["G_data"], z["G_indices"], z["G_indptr"]), shape=(shape[0], shape[0]))
    adata = AnnData(X, obs, pd.DataFrame(index=[str(s) for s in z["var_names"]]), obsp={"connectivities": G})
    df = pd.DataFrame(z["score_values"], index=names, columns=[str(s) for s in z["score_columns"]])
    ref = {
        "group": pd.DataFrame(z["grp_values"], index=[str(s) for s in z["grp_index"]], columns=[str(s) for s in z["grp_columns"]]),
        "corr": z["corr_values"],
        "gene": pd.DataFrame(z["gene_values"], index=[str(s) for s in z["gene_index"]], columns=["CORR", "RANK"]),
        "rls": pd.DataFrame(z["rls_values"], index=[str(s) for s in z["rls_index"]], columns=[str(s) for s in z["rls_columns"]]),
        "rls_ctrl": pd.DataFrame(z["rls_ctrl_values"], index=[str(s) for s in z["rls_index"]], columns=[str(s) for s in z["rls_columns"]]),
        "gearys_c": float(z["gearys_c"]),
    }
    return adata, df, ref


def make_downstream_inputs(n_cell, n_ctrl, n_group, k=10, seed=0, dtype=np.float64, n_gene=0):
    """Seeded score frame + cell groups + symmetric random-neighbour graph for the downstream analyses."""
    rng = np.random.default_rng(seed)
    names = ["c%07d" % i for i in rng.permutation(n_cell)]
    labels = np.minimum((rng.pareto(1.2, n_cell) * 1.5).astype(np.int64), n_group - 1)
    obs = pd.DataFrame({"grp": ["g%03d" % v for v in labels], "v1": rng.normal(size=n_cell)}, index=names)
    norm = (rng.normal(size=n_cell) + 0.4 * obs["v1"].values).astype(dtype)
    ctrl = rng.normal(size=(n_cell, n_ctrl)).astype(dtype)
    ctrl[::5, 0] = 0.25                                                  # ties
    # neighbours mostly inside the own group (sorted by label, nearby positions), some anywhere
    by_label = np.argsort(labels, kind="stable")
    rank = np.empty(n_cell, np.int64)
    rank[by_label] = np.arange(n_cell)
    off = rng.integers(1, 40, size=(n_cell, k)) * rng.choice([-1, 1], size=(n_cell, k))
    nbr = by_label[np.clip(rank[:, None] + off, 0, n_cell - 1)]
    far = rng.random((n_cell, k)) < 0.15
    nbr[far] = rng.integers(0, n_cell, far.sum())
    w = rng.uniform(0.05, 1.0, size=(n_cell, k))
    G = sparse.csr_matrix((w.reshape(-1), (np.repeat(np.arange(n_cell), k), nbr.reshape(-1))), shape=(n_cell, n_cell))
    G.setdiag(0)
    G.eliminate_zeros()
    G = ((G + G.T) / 2).tocsr().astype(np.float32)
    cols = {"norm_score": norm, "pval": (rng.uniform(size=n_cell) ** 3).astype(dtype)}
    for i in range(n_ctrl):
        cols["ctrl_norm_score_%d" % i] = ctrl[:, i]
    df = pd.DataFrame(cols, index=names)
    if n_gene:
        X = sparse.random(n_cell, n_gene, density=0.1, random_state=seed, format="csr", dtype=np.float32)
    else:
        X = sparse.csr_matrix((n_cell, 1), dtype=np.float32)
    adata = AnnData(X, obs, None, obsp={"connectivities": G})
    return adata, df
