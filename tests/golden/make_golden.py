"""Generate the committed golden vectors from the UNMODIFIED reference sources.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Writes tests/golden/ref_case_*.npz.  Each file holds seeded synthetic inputs, the SCDRS_PARAM the
reference's own ``preprocess`` produced for them (with scdrs_b200.loess1d standing in for the absent
skmisc.loess; that LOESS is pinned separately against the toy TSV fixtures) and the outputs of the
reference's ``score_cell`` / ``_select_ctrl_geneset`` / ``_correct_background`` for them.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402
from scdrs_b200 import synth  # noqa: E402
from scdrs_b200.loess1d import loess  # noqa: E402

CASES = [
    # name, n_cell, n_gene, density, use_cov, n_group, weighted, key, weight_opt, n_ctrl, n_trait_gene, seed
    ("sparse_cov_vs", 300, 1200, 0.15, True, 0, False, "mean_var", "vs", 30, 120, 11),
    ("sparse_nocov_weighted", 257, 1100, 0.12, False, 0, True, "mean_var", "vs", 25, 100, 12),
    ("sparse_cov_adjprop_invstd_meankey", 320, 1300, 0.10, True, 6, True, "mean", "inv_std", 20, 90, 13),
    ("sparse_nocov_uniform", 200, 1000, 0.2, False, 0, False, "mean_var", "uniform", 40, 60, 14),
]


def main():
    pp, method = ref_shim.load(loess_impl=loess)
    for (name, n_cell, n_gene, dens, use_cov, n_group, weighted, key, wopt, n_ctrl, n_tg, seed) in CASES:
        adata, cov = synth.make_adata(n_cell, n_gene, density=dens, seed=seed, n_group=n_group)
        pp.preprocess(adata, cov=cov if use_cov else None, adj_prop="cell_type" if n_group else None)
        param = adata.uns["SCDRS_PARAM"]
        traits = synth.make_traits(adata.var_names, n_trait=1, n_gene_per_trait=n_tg, weighted=weighted, seed=seed)
        genes, gw = traits["trait_00"]
        df = method.score_cell(adata, genes, gene_weight=gw, ctrl_match_key=key, n_ctrl=n_ctrl, weight_opt=wopt,
                               return_ctrl_raw_score=True, return_ctrl_norm_score=True)
        df_gene = param["GENE_STATS"].loc[adata.var_names].copy()
        df_gene["gene"] = df_gene.index
        glist = sorted(set(genes) & set(df_gene["gene"]))
        w_of = dict(zip(genes, gw))
        ctrl_list, ctrl_w = method._select_ctrl_geneset(df_gene, glist, [w_of[g] for g in glist], key, n_ctrl, 200, 0)
        col_of = {g: i for i, g in enumerate(adata.var_names)}
        ctrl_idx = np.array([[col_of[g] for g in ctrl_list[i]] for i in range(n_ctrl)], dtype=np.int32)
        ctrl_wt = np.array([ctrl_w[i] for i in range(n_ctrl)], dtype=np.float64)
        X = adata.X.tocsr()
        gs = param["GENE_STATS"]
        out = dict(
            X_data=X.data.astype(np.float32), X_indices=X.indices.astype(np.int32), X_indptr=X.indptr.astype(np.int64),
            shape=np.array(X.shape), obs_names=np.array(adata.obs_names, dtype="U"), var_names=np.array(adata.var_names, dtype="U"),
            use_cov=use_cov, cov_values=cov.values, cov_columns=np.array(cov.columns, dtype="U"),
            cell_type=np.array(adata.obs["cell_type"], dtype="U") if n_group else np.array([], dtype="U"),
            gene_list=np.array(genes, dtype="U"), gene_weight=np.array(gw, dtype=np.float64),
            ctrl_match_key=key, weight_opt=wopt, n_ctrl=n_ctrl,
            gs_mean=gs["mean"].values.astype(np.float64), gs_var=gs["var"].values.astype(np.float64),
            gs_var_tech=gs["var_tech"].values.astype(np.float64), gs_ct_mean=gs["ct_mean"].values.astype(np.float64),
            gs_ct_var=gs["ct_var"].values.astype(np.float64), gs_ct_var_tech=gs["ct_var_tech"].values.astype(np.float64),
            gs_mean_var=np.array(gs["mean_var"].astype(str), dtype="U"),
            cs_mean=param["CELL_STATS"]["mean"].values.astype(np.float64),
            cs_var=param["CELL_STATS"]["var"].values.astype(np.float64),
            ref_columns=np.array(df.columns, dtype="U"), ref_values=df.values.astype(np.float32),
            ref_ctrl_idx=ctrl_idx, ref_ctrl_weight=ctrl_wt,
        )
        if use_cov:
            out["cov_mat"] = param["COV_MAT"].values.astype(np.float64)
            out["cov_mat_columns"] = np.array(param["COV_MAT"].columns, dtype="U")
            out["cov_beta"] = param["COV_BETA"].values.astype(np.float64)
            out["cov_gene_mean"] = param["COV_GENE_MEAN"].values.astype(np.float64)
        path = os.path.join(HERE, "ref_case_%s.npz" % name)
        np.savez_compressed(path, **out)
        print(name, "->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
