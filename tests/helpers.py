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
