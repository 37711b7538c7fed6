"""CPU tests that pin the oracle (numpy restatement) before anything is compared with it:
the reference's golden score files for the toy data, its own known-answer tests, and outputs of the
unmodified reference sources committed under tests/golden/ (plus a live comparison when
/root/reference is present)."""
import os

import numpy as np
import pandas as pd
import pytest

import helpers
from oracle import ref_shim, scdrs_oracle as orc
from scdrs_b200.loess1d import loess

COLS = ["raw_score", "norm_score", "mc_pval", "pval", "nlog10_pval", "zscore"]


def _toy_scored(cov, dense):
    adata, df_cov, df_gs, ref, stats = helpers.load_toy()
    if dense:
        adata.X = adata.X.toarray()
    orc.preprocess(adata, cov=df_cov if cov else None, loess_impl=loess)
    # like the reference's tests (tests/test_method_score_cell_main.py:93-98): golden bins
    key = "cov" if cov else "nocov"
    adata.uns["SCDRS_PARAM"]["GENE_STATS"]["mean_var"] = stats[key]["mean_var"]
    genes = df_gs.loc["toydata_gs_mouse", "GENESET"].split(",")
    df = orc.score_cell(adata, genes, ctrl_match_key="mean_var", n_ctrl=20, weight_opt="vs",
                        return_ctrl_norm_score=True)
    return df, ref, stats, adata


@pytest.mark.parametrize("cov", [False, True])
@pytest.mark.parametrize("dense", [False, True])
def test_oracle_reproduces_reference_golden_scores(cov, dense):
    df, ref, _, _ = _toy_scored(cov, dense)
    gold = ref["REF_COV" if cov else "REF_NOCOV"]
    for c in COLS:
        # the reference's own bar is rtol=1e-2 (tests/test_method_score_cell_main.py:78); files hold 8 digits
        assert np.allclose(df[c].values, gold[c].values, rtol=2e-5, atol=2e-5), c


def test_oracle_reproduces_reference_golden_ctrl_norm_scores():
    df, ref, _, _ = _toy_scored(True, False)
    gold = ref["REF_COV_FULL"]
    cols = ["ctrl_norm_score_%d" % i for i in range(20)]
    assert np.allclose(df[cols].values, gold[cols].values, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("key", ["nocov", "cov"])
def test_oracle_gene_stats_match_reference_tsv(key):
    adata, df_cov, _, _, stats = helpers.load_toy()
    orc.preprocess(adata, cov=df_cov if key == "cov" else None, loess_impl=loess)
    gs, gold = adata.uns["SCDRS_PARAM"]["GENE_STATS"], stats[key]
    for c in ["mean", "var", "ct_mean", "ct_var"]:
        assert np.allclose(gs[c].values.astype(float), gold[c].values, rtol=2e-5, atol=1e-9), c
    # the TSVs were written by an older reference version: their ct_mean / ct_var differ from a
    # fresh run in the 6th digit and the LOESS trend amplifies that; the LOESS itself is pinned
    # input -> output in tests/test_loess.py
    for c in ["ct_var_tech", "var_tech"]:
        assert np.allclose(gs[c].values.astype(float), gold[c].values, rtol=5e-4, atol=1e-9), c
    assert (gs["mean_var"].astype(str).values == gold["mean_var"].values).all()


def test_known_answer_empi_null():
    # tests/test_method_score_cell_other.py:449-458 of the reference
    p = orc.get_p_from_empi_null([0, 1], [0.5, 0.6, 0.7])
    assert np.allclose(p, [1, 0.25])


@pytest.mark.parametrize("path", helpers.golden_cases())
def test_oracle_matches_committed_reference_outputs(path):
    adata, cov, case = helpers.load_case(path)
    df, info = orc.score_cell(adata, case["gene_list"], gene_weight=case["gene_weight"],
                              ctrl_match_key=case["ctrl_match_key"], n_ctrl=case["n_ctrl"],
                              weight_opt=case["weight_opt"], return_ctrl_raw_score=True,
                              return_ctrl_norm_score=True, return_internals=True)
    assert np.array_equal(info["ctrl_cols"], case["ref_ctrl_idx"])          # control sets bit-identical
    ref = case["ref"]
    assert list(df.columns) == list(ref.columns)
    assert np.allclose(df.values, ref.values, rtol=1e-6, atol=1e-6)
    for c in ["mc_pval", "pval"]:
        assert np.array_equal(df[c].values, ref[c].values), c                # integer ranks identical


@pytest.mark.parametrize("path", helpers.golden_cases())
def test_oracle_preprocess_matches_committed_reference_stats(path):
    adata, cov, case = helpers.load_case(path, with_param=False)
    orc.preprocess(adata, cov=cov if case["use_cov"] else None, adj_prop=case["adj_prop"], loess_impl=loess)
    gs, z = adata.uns["SCDRS_PARAM"]["GENE_STATS"], case["z"]
    tol = 1e-9 if case["use_cov"] else 2e-5   # the reference's normal sparse mode accumulates in float32
    for c in ["mean", "var", "var_tech", "ct_mean", "ct_var", "ct_var_tech"]:
        assert np.allclose(gs[c].values.astype(float), z["gs_" + c], rtol=tol, atol=1e-12), c
    assert (gs["mean_var"].astype(str).values == z["gs_mean_var"]).all()


@pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present on this box")
def test_oracle_matches_live_reference():
    from scdrs_b200 import synth

    pp, method = ref_shim.load(loess_impl=loess)
    adata, cov = synth.make_adata(400, 2000, density=0.1, seed=21)
    pp.preprocess(adata, cov=cov)
    genes, w = synth.make_traits(adata.var_names, 1, 150, weighted=True, seed=5)["trait_00"]
    ref = method.score_cell(adata, genes, gene_weight=w, n_ctrl=40, return_ctrl_norm_score=True)
    mine = orc.score_cell(adata, genes, gene_weight=w, n_ctrl=40, return_ctrl_norm_score=True)
    assert np.array_equal(ref.values, mine.values)


# ---- downstream analyses: oracle/downstream_oracle.py --------------------
from oracle import downstream_oracle as dso  # noqa: E402


def _toy_downstream():
    adata, _, _, ref, _ = helpers.load_toy()
    full = ref["REF_COV_FULL"]
    ctrl = full[[c for c in full.columns if c.startswith("ctrl_norm_score")]].values
    pre = os.path.join(helpers.TOY, "toydata_gs_mouse.ref_Ctrl20_CovConstCovariate.")
    gold = {k: pd.read_csv(pre + k, sep="\t", index_col=0) for k in ("scdrs_cell_corr", "scdrs_gene", "scdrs_group.cell_type")}
    return adata, full, ctrl, gold


def test_downstream_oracle_reproduces_reference_golden_files():
    """The reference's own outputs for the toy data (scdrs/data/*.scdrs_*): every column that does not need
    scanpy's neighbour graph."""
    adata, full, ctrl, gold = _toy_downstream()
    assert list(full.index) == list(adata.obs_names)
    vcols = list(gold["scdrs_cell_corr"].index)
    got = dso.corr_analysis([adata.obs[v].values.astype(np.float64) for v in vcols], full["norm_score"].values, ctrl)
    np.testing.assert_allclose(got, gold["scdrs_cell_corr"].values, rtol=2e-6)
    gene = dso.gene_analysis(adata.X, adata.var_names, full["norm_score"].values)
    g = gold["scdrs_gene"]
    np.testing.assert_allclose(gene.loc[g.index, "CORR"].values, g["CORR"].values, rtol=1e-5, atol=1e-7)
    # ranks differ only inside runs of tied correlations (the reference sorts with an unstable quicksort)
    differ = gene.loc[g.index, "RANK"].values != g["RANK"].values
    assert differ.mean() < 0.01
    np.testing.assert_allclose(gene["CORR"].values[g["RANK"].values[differ].astype(int)], g["CORR"].values[differ], rtol=1e-5, atol=1e-7)
    grp = dso.group_analysis(None, adata.obs["cell_type"].values, full["norm_score"].values, ctrl, full["pval"].values)
    gg = gold["scdrs_group.cell_type"]
    for col in ["n_cell", "n_ctrl", "assoc_mcp", "assoc_mcz", "n_fdr_0.05", "n_fdr_0.1", "n_fdr_0.2"]:
        np.testing.assert_allclose(grp.loc[gg.index, col].values.astype(float), gg[col].values, rtol=2e-6, err_msg=col)


def test_downstream_oracle_matches_committed_reference_outputs():
    adata, df, ref = helpers.load_downstream_case()
    order = sorted(adata.obs_names)
    sub = adata[order]
    dfo = df.loc[order]
    ctrl = dfo[[c for c in df.columns if c.startswith("ctrl_norm_score")]].values
    G = sub.obsp["connectivities"]
    grp = dso.group_analysis(G, sub.obs["grp"].values, dfo["norm_score"].values, ctrl, dfo["pval"].values)
    np.testing.assert_allclose(grp.loc[ref["group"].index].values.astype(float), ref["group"].values, rtol=1e-5, equal_nan=True)
    rls = dso.test_gearysc(G, sub.obs["grp"].values, dfo["norm_score"].values, ctrl)
    np.testing.assert_allclose(rls.loc[ref["rls"].index].values, ref["rls"].values, rtol=1e-9, equal_nan=True)
    rls = dso.test_gearysc(G, sub.obs["grp"].values, dfo["norm_score"].values, ctrl, opt="control")
    np.testing.assert_allclose(rls.loc[ref["rls"].index].values, ref["rls_ctrl"].values, rtol=1e-9, equal_nan=True)
    corr = dso.corr_analysis([sub.obs["v1"].values, sub.obs["v2"].values], dfo["norm_score"].values, ctrl)
    np.testing.assert_allclose(corr, ref["corr"], rtol=1e-6)
    gene = dso.gene_analysis(sub.X, adata.var_names, dfo["norm_score"].values)
    np.testing.assert_allclose(gene.loc[ref["gene"].index].values, ref["gene"].values, rtol=1e-5, atol=1e-7)
    gc = dso.gearys_c(adata.obsp["connectivities"], df["norm_score"].values)
    assert abs(gc - ref["gearys_c"]) < 1e-12


def test_gearys_c_oracle_against_the_double_loop():
    """tests/test_method_downstream.py:14-24 of the reference: the naive O(N^2) definition."""
    adata, df = helpers.make_downstream_inputs(60, 2, 3, k=5, seed=3)
    W = adata.obsp["connectivities"].toarray().astype(np.float64)
    x = np.arange(60, dtype=np.float64)
    c = 0.0
    for i in range(60):
        for j in range(60):
            c += W[i, j] * (x[i] - x[j]) ** 2
    c = c * 59 / 2 / W.sum() / np.sum((x - x.mean()) ** 2)
    assert np.allclose(dso.gearys_c(adata.obsp["connectivities"], x), c)
