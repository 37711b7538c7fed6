"""Parity of the device path of the downstream analyses (scdrs_b200/downstream.py, csrc/downstream.cu) with
the reference's golden files for the toy data, with outputs of the unmodified reference committed under
tests/golden/ and with the oracle (oracle/downstream_oracle.py) on larger seeded inputs.  Counts are exact;
floating-point statistics within 1e-5 relative (Monte-Carlo p-values are ratios of exact counts)."""
import os

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

import helpers
from oracle import downstream_oracle as dso

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def method(native_lib):
    import scdrs_b200

    return scdrs_b200.method


def _ctrl_cols(df):
    return [c for c in df.columns if c.startswith("ctrl_norm_score")]


def test_committed_reference_outputs(method):
    adata, df, ref = helpers.load_downstream_case()
    grp = method.downstream_group_analysis(adata, df, ["grp"])["grp"]
    assert list(grp.columns) == list(ref["group"].columns) and grp.index.name == "group"
    assert all(dt == np.float32 for dt in grp.dtypes)
    got = grp.loc[ref["group"].index].values.astype(np.float64)
    np.testing.assert_allclose(got, ref["group"].values, rtol=RTOL, equal_nan=True)
    for col in ["n_cell", "n_ctrl", "assoc_mcp", "hetero_mcp", "n_fdr_0.05", "n_fdr_0.1", "n_fdr_0.2"]:      # counts: exact
        j = list(grp.columns).index(col)
        np.testing.assert_array_equal(got[:, j].astype(np.float32), ref["group"].values[:, j].astype(np.float32))

    corr = method.downstream_corr_analysis(adata, df, ["v1", "v2"])
    np.testing.assert_allclose(corr.values.astype(np.float64), ref["corr"], rtol=RTOL)

    gene = method.downstream_gene_analysis(adata, df)
    np.testing.assert_allclose(gene.loc[ref["gene"].index, "CORR"].values, ref["gene"]["CORR"].values, rtol=RTOL, atol=1e-7)
    assert (gene.loc[ref["gene"].index, "RANK"].values == ref["gene"]["RANK"].values).all()

    order = sorted(adata.obs_names)
    sub, dfo = adata[order], df.loc[order]
    for opt, key in (("control_distribution_match", "rls"), ("control", "rls_ctrl")):
        rls = method.test_gearysc(sub, dfo, "grp", opt=opt)
        assert list(rls.index) == list(ref[key].index) and list(rls.columns) == list(ref[key].columns)
        np.testing.assert_allclose(rls.values, ref[key].values, rtol=1e-9, equal_nan=True)
    gc = method.gearys_c(adata, df["norm_score"].values)
    assert abs(gc - ref["gearys_c"]) < 1e-12 * abs(ref["gearys_c"]) + 1e-15


def test_toy_golden_files(method):
    """scdrs/data/toydata_gs_mouse.ref_Ctrl20_CovConstCovariate.scdrs_{cell_corr,gene,group.cell_type} of the
    reference (tests/test_CLI.py:175-190 compares at rtol 1e-3); the hetero_* columns need scanpy's graph and are
    checked against the oracle on a stand-in graph instead."""
    adata, _, _, ref, _ = helpers.load_toy()
    full = ref["REF_COV_FULL"]
    pre = os.path.join(helpers.TOY, "toydata_gs_mouse.ref_Ctrl20_CovConstCovariate.")
    gold = {k: pd.read_csv(pre + k, sep="\t", index_col=0) for k in ("scdrs_cell_corr", "scdrs_gene", "scdrs_group.cell_type")}
    corr = method.downstream_corr_analysis(adata, full, list(gold["scdrs_cell_corr"].index))
    np.testing.assert_allclose(corr.values, gold["scdrs_cell_corr"].values, rtol=2e-6)
    gene = method.downstream_gene_analysis(adata, full)
    g = gold["scdrs_gene"]
    np.testing.assert_allclose(gene.loc[g.index, "CORR"].values, g["CORR"].values, rtol=RTOL, atol=1e-7)
    differ = gene.loc[g.index, "RANK"].values != g["RANK"].values          # ties only (unstable sort in the reference)
    assert differ.mean() < 0.01
    dense = helpers.AnnData(adata.X.toarray(), adata.obs, adata.var)
    gene_dense = method.downstream_gene_analysis(dense, full)
    np.testing.assert_allclose(gene_dense.loc[g.index, "CORR"].values, g["CORR"].values, rtol=RTOL, atol=2e-7)

    rng = np.random.default_rng(0)
    n = adata.shape[0]
    W = sparse.random(n, n, density=0.3, random_state=1, format="csr", dtype=np.float64)
    W.setdiag(0)
    W.eliminate_zeros()
    adata.obsp["connectivities"] = ((W + W.T) / 2).tocsr()
    grp = method.downstream_group_analysis(adata, full, ["cell_type"])["cell_type"]
    gg = gold["scdrs_group.cell_type"]
    for col in ["n_cell", "n_ctrl", "assoc_mcp", "assoc_mcz", "n_fdr_0.05", "n_fdr_0.1", "n_fdr_0.2"]:
        np.testing.assert_allclose(grp.loc[gg.index, col].values.astype(float), gg[col].values, rtol=2e-6, err_msg=col)
    ctrl = full[_ctrl_cols(full)].values
    want = dso.group_analysis(adata.obsp["connectivities"], adata.obs["cell_type"].values, full["norm_score"].values, ctrl, full["pval"].values)
    np.testing.assert_allclose(grp.loc[want.index].values.astype(float), want.values.astype(float), rtol=RTOL)
    del rng


@pytest.mark.parametrize("dtype,n_cell,n_ctrl,n_group", [(np.float64, 6000, 40, 25), (np.float32, 3000, 70, 7),
                                                         (np.float64, 900, 1100, 3)])
def test_group_statistics_against_oracle(method, dtype, n_cell, n_ctrl, n_group):
    adata, df = helpers.make_downstream_inputs(n_cell, n_ctrl, n_group, seed=n_cell, dtype=dtype)
    got = method.downstream_group_analysis(adata, df, ["grp"])["grp"]
    order = sorted(adata.obs_names)
    sub, dfo = adata[order], df.loc[order]
    ctrl = dfo[_ctrl_cols(df)].values
    want = dso.group_analysis(sub.obsp["connectivities"], sub.obs["grp"].values, dfo["norm_score"].values, ctrl, dfo["pval"].values)
    assert list(got.index) == list(want.index)
    np.testing.assert_allclose(got.values.astype(float), want.values.astype(float), rtol=RTOL, equal_nan=True)
    for col in ["n_cell", "assoc_mcp", "hetero_mcp", "n_fdr_0.1"]:
        np.testing.assert_array_equal(got[col].values, want[col].values)
    # Geary's C itself, to 1e-10: sums of the same fp64 terms in a different order
    rls = method.test_gearysc(sub, dfo, "grp")
    want_rls = dso.test_gearysc(sub.obsp["connectivities"], sub.obs["grp"].values, dfo["norm_score"].values, ctrl)
    np.testing.assert_allclose(rls.loc[want_rls.index].values, want_rls.values, rtol=1e-9, equal_nan=True)


def test_gearysc_permutation_option_uses_numpy_global_rng(method):
    adata, df = helpers.make_downstream_inputs(400, 6, 3, seed=9)
    ctrl = df[_ctrl_cols(df)].values
    np.random.seed(4)
    got = method.test_gearysc(adata, df, "grp", opt="permutation")
    np.random.seed(4)
    want = dso.test_gearysc(adata.obsp["connectivities"], adata.obs["grp"].values, df["norm_score"].values, ctrl, opt="permutation")
    np.testing.assert_allclose(got.loc[want.index].values, want.values, rtol=1e-9, equal_nan=True)
    with pytest.raises(NotImplementedError):
        method.test_gearysc(adata, df, "grp", opt="nope")


def test_gearys_c_against_the_double_loop(method):
    """tests/test_method_downstream.py:115-128 of the reference."""
    adata, _ = helpers.make_downstream_inputs(50, 1, 2, k=6, seed=1)
    W = adata.obsp["connectivities"].toarray().astype(np.float64)
    x = np.arange(50)
    c = 0.0
    for i in range(50):
        for j in range(50):
            c += W[i, j] * (x[i] - x[j]) ** 2
    c = c * 49 / 2 / np.sum(W) / np.sum((x - np.mean(x)) ** 2)
    assert np.allclose(method.gearys_c(adata, x), c)


def test_corr_analysis_against_oracle_and_numpy(method):
    adata, df = helpers.make_downstream_inputs(20000, 200, 5, seed=2, dtype=np.float32)
    adata.obs["v2"] = np.arange(adata.shape[0], dtype=np.float64) % 17
    got = method.downstream_corr_analysis(adata, df, ["v1", "v2"])
    order = sorted(adata.obs_names)
    dfo = df.loc[order]
    want = dso.corr_analysis([adata.obs.loc[order, "v1"].values, adata.obs.loc[order, "v2"].values],
                             dfo["norm_score"].values.astype(np.float64), dfo[_ctrl_cols(df)].values.astype(np.float64))
    np.testing.assert_allclose(got.values.astype(float), want, rtol=RTOL)
    assert list(got.columns) == ["n_ctrl", "corr_mcp", "corr_mcz"] and list(got.index) == ["v1", "v2"]
    with pytest.raises(AssertionError, match="Missing `var_cols`"):
        method.downstream_corr_analysis(adata, df, ["nope"])


def test_gene_analysis_on_a_cell_subset(method):
    adata, df = helpers.make_downstream_inputs(3000, 2, 3, seed=6, n_gene=300)
    sub_df = df.iloc[::3]
    got = method.downstream_gene_analysis(adata, sub_df)
    order = sorted(sub_df.index)
    want = dso.gene_analysis(adata[order].X, adata.var_names, sub_df.loc[order, "norm_score"].values)
    np.testing.assert_allclose(got.loc[want.index, "CORR"].values, want["CORR"].values, rtol=RTOL, atol=1e-7)
    assert (got.loc[want.index, "RANK"].values == want["RANK"].values).mean() > 0.99


def test_pearson_corr_input_variants(method):
    """tests/test_method_downstream.py:212-291 of the reference: dense / sparse / 1-D combinations."""
    np.random.seed(0)
    mat_X = np.random.randn(3, 2)
    mat_Y = np.random.randn(3, 2)
    Xs, Ys = sparse.csr_matrix(mat_X), sparse.csr_matrix(mat_Y)
    true = np.array([[np.corrcoef(mat_X[:, i], mat_Y[:, j])[0, 1] for j in range(2)] for i in range(2)])
    tol = dict(rtol=1e-5, atol=1e-6)
    assert np.allclose(method._pearson_corr(mat_X, mat_Y), true, **tol)
    assert np.allclose(method._pearson_corr(Xs, Ys), true, **tol)
    assert np.allclose(method._pearson_corr(Xs, mat_Y), true, **tol)
    assert np.allclose(method._pearson_corr(mat_X, Ys), true, **tol)
    assert np.allclose(method._pearson_corr(mat_X[:, 0], mat_Y[:, 0]), true[0, 0], **tol)
    assert np.allclose(method._pearson_corr(mat_X[:, 0], mat_Y), true[0, :], **tol)
    assert np.allclose(method._pearson_corr(mat_X, mat_Y[:, 0]), true[:, 0], **tol)
    assert np.allclose(method._pearson_corr(mat_X[:, 0], Ys), true[0, :], **tol)
    assert np.allclose(method._pearson_corr(mat_X, Ys[:, 0]), true[:, 0], **tol)
    assert np.allclose(method._pearson_corr_sparse(mat_X, mat_Y), true, **tol)


def test_group_analysis_argument_checks(method):
    adata, df = helpers.make_downstream_inputs(100, 3, 2, seed=8)
    with pytest.raises(AssertionError, match="Missing `group_cols`"):
        method.downstream_group_analysis(adata, df, ["nope"])
    adata.obsp.pop("connectivities")
    with pytest.raises(AssertionError, match="Expect `connectivities`"):
        method.downstream_group_analysis(adata, df, ["grp"])


def test_cli_perform_downstream(native_lib, tmp_path):
    """tests/test_CLI.py:137-190 of the reference (corr and gene analyses; the toy .h5ad carries no neighbour
    graph, so the group analysis must ask for one)."""
    from scdrs_b200 import cli

    args = ["perform-downstream", "--h5ad_file", os.path.join(helpers.TOY, "toydata_mouse.h5ad"),
            "--score-file", os.path.join(helpers.TOY, "@.full_score.gz"), "--flag-filter-data", "False",
            "--flag-raw-count", "False", "--knn-n-neighbors", "15", "--knn-n-pcs", "20", "--out-folder", str(tmp_path)]
    cli.main(args + ["--corr-analysis", "causal_variable,non_causal_variable,covariate"])
    cli.main(args + ["--gene-analysis"])
    prefix = "toydata_gs_mouse.ref_Ctrl20_CovConstCovariate"
    for suffix in ["scdrs_gene", "scdrs_cell_corr"]:
        df_res = pd.read_csv(os.path.join(str(tmp_path), "%s.%s" % (prefix, suffix)), sep="\t", index_col=0)
        df_ref = pd.read_csv(os.path.join(helpers.TOY, "%s.%s" % (prefix, suffix)), sep="\t", index_col=0)
        if suffix == "scdrs_gene":
            np.testing.assert_allclose(df_res.loc[df_ref.index, "CORR"].values, df_ref["CORR"].values, rtol=1e-5, atol=1e-7)
        else:
            assert list(df_res.index) == list(df_ref.index)
            np.testing.assert_allclose(df_res.values, df_ref.values, rtol=1e-5)
    with pytest.raises(ValueError, match="connectivities"):
        cli.main(args + ["--group-analysis", "cell_type"])
