"""Parity of the CUDA path (through the C ABI) with the oracle and with the reference's golden
vectors.  Integer outputs (control indices, rank counts) must be bit-exact; floating-point outputs
must agree within 1e-5 (BASELINE.json north_star), tolerance written in each test."""
import os

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

import helpers
from oracle import scdrs_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-5     # north_star: raw, normalised and z-scores within 1e-5 relative in fp32
ATOL = 1e-5


@pytest.fixture(scope="module", params=["auto", "f64"])
def sc(native_lib, request):
    """The package with the raw-score scatter in its default arithmetic (fixed point where it
    applies) and forced to fp64 accumulators: every parity test runs against both kernels."""
    import scdrs_b200
    from scdrs_b200 import _device

    _device.set_spmm_mode(request.param)
    yield scdrs_b200
    _device.set_spmm_mode("auto")


@pytest.fixture(scope="module")
def ctx(native_lib):
    from scdrs_b200 import _native

    c = _native.Context(0)
    yield c
    c.close()


def _check_counts_self_consistent(df, n_ctrl):
    """mc / pooled counts must be EXACTLY what numpy gets from the returned float32 scores."""
    norm = df["norm_score"].values
    ctrl = df[["ctrl_norm_score_%d" % i for i in range(n_ctrl)]].values
    mc = (ctrl >= norm[:, None]).sum(axis=1)
    assert np.array_equal(df["mc_pval"].values, ((1 + mc) / (1 + n_ctrl)).astype(np.float32))
    ge = orc.null_ge_count(norm, ctrl.reshape(-1))
    p = ((ge + 1) / (ctrl.size + 1)).astype(np.float32)
    assert np.array_equal(df["pval"].values, p)


def _compare_with_reference(df, ref, n_ctrl, n_cell):
    """Float columns within tolerance; p-value columns equal except where ties straddle it."""
    for c in ["raw_score", "norm_score"]:
        assert np.allclose(df[c].values, ref[c].values, rtol=RTOL, atol=ATOL), c
    cr = ["ctrl_raw_score_%d" % i for i in range(n_ctrl)]
    cn = ["ctrl_norm_score_%d" % i for i in range(n_ctrl)]
    if cr[0] in ref and cr[0] in df:
        assert np.allclose(df[cr].values, ref[cr].values, rtol=RTOL, atol=ATOL)
    if cn[0] in ref and cn[0] in df:
        assert np.allclose(df[cn].values, ref[cn].values, rtol=RTOL, atol=ATOL)
    n_null = n_cell * n_ctrl
    # Rank counts are exact given the scores (checked by _check_counts_self_consistent).  Against the
    # float64 reference a rank may move by the number of null values that lie within the float32
    # rounding of the trait score: about n_null * density * 2e-7 |t|, i.e. a relative change of the
    # p-value of order hazard(t) * 1e-7 |t|.  Bound it at 1e-4 relative and report the count.
    p_got, p_ref = df["pval"].values.astype(np.float64), ref["pval"].values.astype(np.float64)
    r_got, r_ref = np.round(p_got * (n_null + 1)), np.round(p_ref * (n_null + 1))
    d_rank = np.abs(r_got - r_ref)
    assert (d_rank <= 1 + 1e-4 * r_ref).all(), (d_rank.max(), np.abs(p_got / p_ref - 1).max())
    n_diff = int((d_rank > 0).sum())
    d_mc = np.abs(df["mc_pval"].values.astype(np.float64) - ref["mc_pval"].values) * (1 + n_ctrl)
    assert d_mc.max() <= 1.01 and (d_mc > 0.5).sum() <= max(1, int(n_null * 1e-5)), (d_mc > 0.5).sum()
    same = d_rank == 0
    assert np.allclose(df["zscore"].values[same], ref["zscore"].values[same], rtol=RTOL, atol=ATOL)
    assert np.allclose(df["nlog10_pval"].values[same], ref["nlog10_pval"].values[same], rtol=RTOL, atol=ATOL)
    print("cells whose pooled rank differs from the float64 reference: %d / %d (max |d rank| %d)" % (
        n_diff, n_cell, int(d_rank.max())))
    return n_diff


@pytest.mark.parametrize("path", helpers.golden_cases())
def test_score_cell_matches_reference_golden_outputs(sc, path):
    adata, cov, case = helpers.load_case(path)
    df = sc.score_cell(adata, case["gene_list"], gene_weight=case["gene_weight"],
                       ctrl_match_key=case["ctrl_match_key"], n_ctrl=case["n_ctrl"],
                       weight_opt=case["weight_opt"], return_ctrl_raw_score=True, return_ctrl_norm_score=True)
    ref = case["ref"]
    assert list(df.columns) == list(ref.columns) and df.values.dtype == np.float32
    assert list(df.index) == list(ref.index)
    _check_counts_self_consistent(df, case["n_ctrl"])
    _compare_with_reference(df, ref, case["n_ctrl"], adata.shape[0])


@pytest.mark.parametrize("cov", [False, True])
@pytest.mark.parametrize("dense", [False, True])
def test_score_cell_toy_data_golden_files(sc, cov, dense):
    """The reference's own integration tests (tests/test_method_score_cell_main.py:82-177)."""
    adata, df_cov, df_gs, ref, stats = helpers.load_toy()
    genes = df_gs.loc["toydata_gs_mouse", "GENESET"].split(",")
    genes = sorted(set(genes) & set(adata.var_names))
    if dense:
        adata.X = adata.X.toarray()
    sc.preprocess(adata, cov=df_cov if cov else None, n_mean_bin=20, n_var_bin=20, copy=False)
    key = "cov" if cov else "nocov"
    adata.uns["SCDRS_PARAM"]["GENE_STATS"]["mean_var"] = stats[key]["mean_var"]
    df = sc.score_cell(adata, genes, ctrl_match_key="mean_var", n_ctrl=20, weight_opt="vs",
                       return_ctrl_norm_score=True)
    gold = ref["REF_COV" if cov else "REF_NOCOV"]
    for c in ["raw_score", "norm_score", "mc_pval", "pval"]:
        assert np.allclose(df[c].values, gold[c].values, rtol=1e-2, equal_nan=True), c   # the reference's bar
    # and the tight bar: the golden files carry 8 significant digits
    for c in ["raw_score", "norm_score", "mc_pval", "pval", "nlog10_pval", "zscore"]:
        assert np.allclose(df[c].values, gold[c].values, rtol=5e-5, atol=5e-5), c
    if cov:
        cols = ["ctrl_norm_score_%d" % i for i in range(20)]
        assert np.allclose(df[cols].values, ref["REF_COV_FULL"][cols].values, rtol=5e-5, atol=5e-5)
    _check_counts_self_consistent(df, 20)


@pytest.mark.parametrize("use_cov,weighted,key,wopt,n_ctrl", [
    (True, False, "mean_var", "vs", 100),
    (False, True, "mean_var", "vs", 64),
    (True, True, "mean", "inv_std", 33),
    (False, False, "mean_var", "uniform", 8),
])
def test_score_cell_matches_oracle_on_synthetic(sc, use_cov, weighted, key, wopt, n_ctrl):
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(1500, 3000, density=0.1, seed=31)
    orc.preprocess(adata, cov=cov if use_cov else None, loess_impl=loess)
    genes, w = synth.make_traits(adata.var_names, 1, 200, weighted=weighted, seed=77)["trait_00"]
    ref, info = orc.score_cell(adata, genes, gene_weight=w, ctrl_match_key=key, n_ctrl=n_ctrl, weight_opt=wopt,
                               return_ctrl_raw_score=True, return_ctrl_norm_score=True, return_internals=True)
    df = sc.score_cell(adata, genes, gene_weight=w, ctrl_match_key=key, n_ctrl=n_ctrl, weight_opt=wopt,
                       return_ctrl_raw_score=True, return_ctrl_norm_score=True)
    _check_counts_self_consistent(df, n_ctrl)
    _compare_with_reference(df, ref, n_ctrl, adata.shape[0])


def test_score_cell_single_control_set_is_all_nan_like_the_reference(sc):
    """n_ctrl = 1: the cell-wise std of one control is 0 and the reference returns NaN scores."""
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, _ = synth.make_adata(300, 1200, density=0.1, seed=6)
    orc.preprocess(adata, cov=None, loess_impl=loess)
    genes, _ = synth.make_traits(adata.var_names, 1, 50, seed=4)["trait_00"]
    ref = orc.score_cell(adata, genes, n_ctrl=1)
    df = sc.score_cell(adata, genes, n_ctrl=1)
    assert np.allclose(df["raw_score"].values, ref["raw_score"].values, rtol=RTOL, atol=ATOL)
    assert np.isnan(ref["norm_score"].values).all() and np.isnan(df["norm_score"].values).all()


def test_score_cell_zero_rows_get_the_minimum(sc):
    """Cells whose raw score is exactly 0 (scdrs/method.py:521-523, 581-583), plus empty rows."""
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, _ = synth.make_adata(400, 1500, density=0.08, seed=5)
    X = adata.X.tolil()
    X[3, :] = 0
    X[17, :] = 0
    adata.X = sparse.csr_matrix(X.tocsr())
    adata.X.eliminate_zeros()
    orc.preprocess(adata, cov=None, loess_impl=loess)
    genes, w = synth.make_traits(adata.var_names, 1, 40, seed=3)["trait_00"]
    ref = orc.score_cell(adata, genes, n_ctrl=25, return_ctrl_norm_score=True, return_ctrl_raw_score=True)
    df = sc.score_cell(adata, genes, n_ctrl=25, return_ctrl_norm_score=True, return_ctrl_raw_score=True)
    assert df["raw_score"].iloc[3] == 0 and df["raw_score"].iloc[17] == 0
    assert (ref["raw_score"] == 0).sum() >= 2
    _check_counts_self_consistent(df, 25)
    _compare_with_reference(df, ref, 25, 400)
    assert df["norm_score"].iloc[3] == df["norm_score"].min()


def test_compute_raw_score_matches_numpy(sc):
    """_compute_raw_score for sparse x cov x weight options (tests/test_method_score_cell_other.py:64-281)."""
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    for use_cov in (False, True):
        adata, cov = synth.make_adata(300, 800, density=0.2, seed=9)
        orc.preprocess(adata, cov=cov if use_cov else None, loess_impl=loess)
        param = adata.uns["SCDRS_PARAM"]
        genes = list(adata.var_names[[5, 17, 200, 301, 644]])
        gw = [1.0, 2.0, 0.5, 3.0, 1.5]
        for wopt in ("uniform", "vs", "inv_std"):
            v_raw, v_w = sc.method._compute_raw_score(adata, genes, gw, wopt)
            idx = adata.var_names.get_indexer(genes)
            w = orc.score_weights(param["GENE_STATS"], idx, gw, wopt)
            kw = {}
            if use_cov:
                kw = dict(cov_mat=param["COV_MAT"].values, cov_beta=param["COV_BETA"].values,
                          gene_mean=param["COV_GENE_MEAN"].values)
            want = orc.compute_raw_score(adata.X.tocsc(), idx, w, **kw)
            assert np.allclose(v_w, w, rtol=1e-12)
            assert np.allclose(v_raw, want, rtol=1e-6, atol=1e-7), (use_cov, wopt)


def test_correct_background_finds_planted_cell(sc):
    """_correct_background under perturbations (tests/test_method_score_cell_other.py:390-446)."""
    rng = np.random.default_rng(0)
    n_cell, n_ctrl = 200, 100
    for config in range(5):
        v_raw = rng.normal(size=n_cell) * 0.1 + 5
        mat = rng.normal(size=(n_cell, n_ctrl)) * 0.1 + 5
        v_raw[5] += 1.0
        ratio = np.ones(n_ctrl)
        if config == 1:
            mat = mat + rng.normal(size=n_ctrl)              # gene-set-wise shifts
        if config == 2:
            v_raw += np.linspace(0, 3, n_cell); mat += np.linspace(0, 3, n_cell)[:, None]   # cell-wise shifts
        if config == 3:
            ratio = rng.uniform(0.5, 2, size=n_ctrl); mat = 5 + (mat - 5) * np.sqrt(ratio)
        if config == 4:
            v_raw[10:20] = 0; mat[10:20, ::3] = 0            # zero masking
        got_t, got_c = sc.method._correct_background(v_raw, mat, ratio)
        want_t, want_c = orc.correct_background(v_raw, mat, ratio)
        assert got_t[5] > 3
        assert np.allclose(got_t, want_t, rtol=RTOL, atol=ATOL), config
        assert np.allclose(got_c, want_c, rtol=RTOL, atol=ATOL), config


def test_get_p_from_empi_null_known_answer_and_exact_counts(sc, ctx):
    # tests/test_method_score_cell_other.py:449-458
    assert np.allclose(sc.method._get_p_from_empi_null([0, 1], [0.5, 0.6, 0.7]), [1, 0.25])
    rng = np.random.default_rng(1)
    for dt in (np.float32, np.float64):
        for n_t, n_null in [(1, 1), (7, 0), (1000, 100000), (4097, 300001)]:
            t = rng.normal(size=n_t).astype(dt)
            null = rng.normal(size=n_null).astype(dt)
            if n_null > 10:
                null[:5] = t[0]                      # exact ties count as >=
                null[5:9] = [np.inf, -np.inf, 0.0, -0.0]
                t[-1] = 0.0
            got = ctx.empi_null_count(t, null)
            assert np.array_equal(got, orc.null_ge_count(t, null)), (dt, n_t, n_null)
    # heavy duplication and a constant query vector
    t = np.repeat(np.float32([0.5, -1.0, 0.5]), 400)
    null = np.round(rng.normal(size=50000), 1).astype(np.float32)
    assert np.array_equal(ctx.empi_null_count(t, null), orc.null_ge_count(t, null))
    t = np.full(300, 2.0, np.float32)
    assert np.array_equal(ctx.empi_null_count(t, null), orc.null_ge_count(t, null))


def test_full_size_properties(sc, ctx):
    """Size-independent properties at a larger shape: linearity of the raw score in the weights,
    rank counts consistent with the returned scores, idempotence of a repeated call."""
    from scdrs_b200 import synth

    X = synth.make_csr(20000, 5000, density=0.1, seed=2)
    ctx.set_csr_host(X.indptr, X.indices, X.data, X.shape[1])
    rng = np.random.default_rng(4)
    genes = np.stack([rng.choice(5000, size=300, replace=False) for _ in range(3)]).astype(np.int32)
    genes[2] = genes[0]
    w = rng.uniform(0.1, 1, size=(3, 300))
    w[2] = 2.0 * w[0]
    raw = ctx.compute_raw_score(genes, w)
    assert np.allclose(raw[:, 2], 2.0 * raw[:, 0], rtol=1e-6, atol=1e-9)                  # linearity
    want = np.asarray(X[:, genes[1]] @ w[1]).reshape(-1)
    assert np.allclose(raw[:, 1], want, rtol=1e-6, atol=1e-7)
    raw2 = ctx.compute_raw_score(genes, w)
    assert np.array_equal(raw, raw2)                                                      # bitwise reproducible


@pytest.mark.parametrize("exact_order", [False, True])
@pytest.mark.parametrize("path", helpers.golden_cases())
def test_preprocess_matches_reference_golden_stats(sc, path, exact_order, monkeypatch):
    """Default: parallel per-gene moments (csrc/stats_par.cu), numerically constant genes recomputed in numpy's
    order.  SCDRS_B200_EXACT_ORDER=1: every gene through the sequential kernels that add cells in numpy's order."""
    monkeypatch.setenv("SCDRS_B200_EXACT_ORDER", "1" if exact_order else "0")
    adata, cov, case = helpers.load_case(path, with_param=False)
    sc.preprocess(adata, cov=cov if case["use_cov"] else None, adj_prop=case["adj_prop"])
    param, z = adata.uns["SCDRS_PARAM"], case["z"]
    assert set(param) >= {"FLAG_SPARSE", "FLAG_COV", "FLAG_ADJ_PROP", "GENE_STATS", "CELL_STATS"}
    assert list(param["GENE_STATS"].columns) == ["mean", "var", "var_tech", "ct_mean", "ct_var", "ct_var_tech", "mean_var"]
    gs = param["GENE_STATS"]
    # reference normal sparse mode accumulates in float32; with covariates its float32-accumulated COV_GENE_MEAN (up to
    # 5e-7 off the exact mean at these sizes) is part of every corrected value, and only the sequential path repeats it
    tol = (1e-7 if exact_order else 2e-6) if case["use_cov"] else 2e-5
    for c in ["mean", "var", "ct_mean", "ct_var", "var_tech", "ct_var_tech"]:
        assert np.allclose(gs[c].values.astype(float), z["gs_" + c], rtol=max(tol, 1e-5 if "tech" in c else tol), atol=1e-10), c
    n_bin_diff = int((gs["mean_var"].astype(str).values != z["gs_mean_var"]).sum())
    assert n_bin_diff <= (0 if (case["use_cov"] and exact_order) else 5), n_bin_diff
    assert np.allclose(param["CELL_STATS"]["mean"].values.astype(float), z["cs_mean"], rtol=1e-5, atol=1e-9)
    assert np.allclose(param["CELL_STATS"]["var"].values.astype(float), z["cs_var"], rtol=1e-4, atol=1e-9)
    if case["use_cov"]:
        assert np.allclose(param["COV_BETA"].values, z["cov_beta"], rtol=1e-9, atol=1e-12)
        if exact_order:
            # float32 gene means in scipy's own operation order: identical bits
            assert np.array_equal(param["COV_GENE_MEAN"].values.astype(np.float64), z["cov_gene_mean"])
        else:
            # the reference ACCUMULATES the mean in float32 (scipy keeps the dtype); the parallel path sums in float64
            # and rounds once: the difference is the reference's own accumulation error
            assert np.allclose(param["COV_GENE_MEAN"].values.astype(np.float64), z["cov_gene_mean"], rtol=2e-6, atol=0)


def test_compute_stats_modes_agree(sc):
    """sparse + cov must equal dense + cov (tests/test_pp.py:115-161 of the reference)."""
    from scdrs_b200 import synth

    a_sp, cov = synth.make_adata(500, 1000, density=0.15, seed=8)
    a_dn = a_sp.copy()
    a_dn.X = a_dn.X.toarray()
    sc.preprocess(a_sp, cov=cov)
    sc.preprocess(a_dn, cov=cov)
    assert a_sp.uns["SCDRS_PARAM"]["FLAG_SPARSE"] and not a_dn.uns["SCDRS_PARAM"]["FLAG_SPARSE"]
    g1, g2 = a_sp.uns["SCDRS_PARAM"]["GENE_STATS"], a_dn.uns["SCDRS_PARAM"]["GENE_STATS"]
    for c in ["mean", "var", "ct_mean", "ct_var"]:
        assert np.allclose(g1[c].values.astype(float), g2[c].values.astype(float), rtol=1e-4, atol=1e-6), c
    c1, c2 = a_sp.uns["SCDRS_PARAM"]["CELL_STATS"], a_dn.uns["SCDRS_PARAM"]["CELL_STATS"]
    assert np.allclose(c1["mean"].values.astype(float), c2["mean"].values.astype(float), rtol=1e-4, atol=1e-6)
    assert np.allclose(c1["var"].values.astype(float), c2["var"].values.astype(float), rtol=1e-3, atol=1e-6)


def test_cli_compute_score_on_toy_data(sc, tmp_path):
    """tests/test_CLI.py:10-71 of the reference: mouse and human gene sets give identical p-values and
    the cov run reproduces the golden score file."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for species, gs in [("mouse", "toydata_mouse.gs"), ("human", "toydata_human.gs")]:
        cmd = [sys.executable, os.path.join(root, "bin", "scdrs"), "compute-score",
               "--h5ad-file", os.path.join(helpers.TOY, "toydata_mouse.h5ad"), "--h5ad-species", "mouse",
               "--gs-file", os.path.join(helpers.TOY, gs), "--gs-species", species,
               "--cov-file", os.path.join(helpers.TOY, "toydata_mouse.cov"), "--ctrl-match-opt", "mean_var",
               "--n-ctrl", "20", "--flag-filter-data", "False", "--weight-opt", "vs", "--flag-raw-count", "False",
               "--flag-return-ctrl-raw-score", "False", "--flag-return-ctrl-norm-score", "True",
               "--out-folder", str(tmp_path)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "FDR<0.1 cells" in r.stdout
    mouse = pd.read_csv(tmp_path / "toydata_gs_mouse.score.gz", sep="\t", index_col=0)
    human = pd.read_csv(tmp_path / "toydata_gs_human.score.gz", sep="\t", index_col=0)
    assert np.allclose(mouse["pval"].values, human["pval"].values)
    gold = helpers.load_toy()[3]["REF_COV"]
    assert list(mouse.columns) == list(gold.columns) and list(mouse.index) == list(gold.index)
    # the reference's own CLI test skips norm_score / pval: live mean_var bins may differ (tests/test_CLI.py:58-69)
    assert np.allclose(mouse["raw_score"].values, gold["raw_score"].values, rtol=1e-2)
    full = pd.read_csv(tmp_path / "toydata_gs_mouse.full_score.gz", sep="\t", index_col=0)
    assert full.shape == (30, 26)


def test_phases_on_two_shards_equal_single_context(native_lib):
    """The multi-GPU phase API with the collectives done by hand: two contexts on one GPU, each
    holding half of the cells, must reproduce the single-context result (counts exactly)."""
    import torch

    from scdrs_b200 import _native, synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(1001, 2000, density=0.1, seed=12)
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    param = adata.uns["SCDRS_PARAM"]
    genes, w = synth.make_traits(adata.var_names, 1, 150, weighted=True, seed=5)["trait_00"]
    _, info = orc.score_cell(adata, genes, gene_weight=w, n_ctrl=50, return_internals=True)
    idx = adata.var_names.get_indexer(sorted(set(genes)))
    w_of = dict(zip(genes, w))
    tgw = np.array([w_of[g] for g in sorted(set(genes))])
    gs = param["GENE_STATS"]

    def make_ctx(lo, hi):
        X = adata.X[lo:hi].tocsr()
        c = _native.Context(0)
        c.set_csr_host(X.indptr, X.indices, X.data, X.shape[1])
        c.set_gene_stats(gs["var"].values.astype(float), gs["var_tech"].values.astype(float))
        c.set_cov(param["COV_MAT"].values[lo:hi], param["COV_BETA"].values, param["COV_GENE_MEAN"].values)
        return c

    n = adata.shape[0]
    full = make_ctx(0, n).score_trait(idx, tgw, info["ctrl_cols"], info["ctrl_gw_raw"], "vs", True, want_ctrl_norm=True)
    shards = [(0, 400), (400, n)]
    ctxs = [make_ctx(lo, hi) for lo, hi in shards]
    ncol = 51
    dev = torch.device("cuda", 0)
    t = lambda m, dt=torch.float64: torch.zeros(m, dtype=dt, device=dev)
    colsum = [t(ncol) for _ in ctxs]
    torch.cuda.synchronize()      # the contexts run on their own streams
    for c, cs in zip(ctxs, colsum):
        c.trait_raw(idx, tgw, info["ctrl_cols"], info["ctrl_gw_raw"], "vs", True, cs.data_ptr()); c.sync()
    tot = colsum[0] + colsum[1]
    col2, cmin = [t(ncol) for _ in ctxs], [t(ncol) for _ in ctxs]
    torch.cuda.synchronize()
    for c, a, b in zip(ctxs, col2, cmin):
        c.trait_rowstats(tot.data_ptr(), n, a.data_ptr(), b.data_ptr()); c.sync()
    col2t, cmint = col2[0] + col2[1], torch.minimum(cmin[0], cmin[1])
    q = [t(hi - lo, torch.float32) for lo, hi in shards]
    torch.cuda.synchronize()
    for c, qq in zip(ctxs, q):
        c.trait_queries(col2t.data_ptr(), cmint.data_ptr(), qq.data_ptr()); c.sync()
    q_all = torch.cat(q)
    hist = [t(n + 1, torch.int64) for _ in ctxs]
    torch.cuda.synchronize()
    for c, h in zip(ctxs, hist):
        c.trait_count(q_all.data_ptr(), n, h.data_ptr()); c.sync()
    hist_t = hist[0] + hist[1]
    torch.cuda.synchronize()
    outs = [c.trait_finish(hist_t.data_ptr(), n, lo, n * 50, 50, want_ctrl_norm=True) for c, (lo, hi) in zip(ctxs, shards)]
    for key in ["ge_count", "mc_count"]:
        assert np.array_equal(np.concatenate([o[key] for o in outs]), full[key]), key
    for key in ["raw_score", "norm_score", "ctrl_norm"]:
        assert np.allclose(np.concatenate([o[key] for o in outs]), full[key], rtol=1e-6, atol=1e-6), key


def test_many_control_sets_shape(sc):
    """n_ctrl = 5000 (config 5's control count) on a small matrix: 40 KB of accumulators per cell."""
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(150, 1500, density=0.1, seed=15)
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    genes, w = synth.make_traits(adata.var_names, 1, 60, weighted=True, seed=8)["trait_00"]
    ref = orc.score_cell(adata, genes, gene_weight=w, n_ctrl=5000)
    df = sc.score_cell(adata, genes, gene_weight=w, n_ctrl=5000)
    _compare_with_reference(df, ref, 5000, 150)


def test_full_size_config2_properties(sc, ctx):
    """BASELINE config 2 at full size (110,096 x 20,000, n_ctrl = 1000): properties that do not need
    the CPU reference -- counts consistent with the returned scores, second alignment (every control
    column has mean 0), cell-wise standardisation (unit spread per cell), bitwise reproducibility."""
    import torch

    import bench

    class Args:
        n_cell, n_gene, density, n_trait, n_trait_gene = 110096, 20000, 0.10, 1, 1000

    adata, traits, _ = bench.build_workload(Args, 0, torch.device("cuda", 0))
    genes, w = traits["trait_00"]
    df = sc.score_cell(adata, genes, gene_weight=w, n_ctrl=1000, return_ctrl_norm_score=True)
    n_ctrl = 1000
    norm = df["norm_score"].values
    ctrl = df[["ctrl_norm_score_%d" % i for i in range(n_ctrl)]].values
    assert np.isfinite(ctrl).all() and np.isfinite(norm).all()
    # integer outputs against numpy on the returned float32 scores
    mc = (ctrl >= norm[:, None]).sum(axis=1)
    assert np.array_equal(df["mc_pval"].values, ((1 + mc) / (1 + n_ctrl)).astype(np.float32))
    ge = orc.null_ge_count(norm, ctrl.reshape(-1))
    assert np.array_equal(df["pval"].values, ((ge + 1) / (ctrl.size + 1)).astype(np.float32))
    # second gene-set alignment: column means are 0 (scdrs/method.py:564-565)
    assert np.abs(ctrl.mean(axis=0, dtype=np.float64)).max() < 1e-5
    assert abs(norm.mean(dtype=np.float64)) < 1e-5
    # cell-wise standardisation (scdrs/method.py:544-548) up to the second alignment's shifts
    sd = ctrl.std(axis=1, dtype=np.float64)
    assert np.abs(sd - 1).max() < 0.05
    # idempotence / bitwise reproducibility of the whole device path
    df2 = sc.score_cell(adata, genes, gene_weight=w, n_ctrl=1000, return_ctrl_norm_score=True)
    assert np.array_equal(df.values, df2.values)


def test_get_mean_var_matches_numpy(sc):
    """_get_mean_var / _get_mean_var_implicit_cov_corr (tests/test_pp.py:384-458, tests/test_pp_adj_prop.py:246-457)."""
    from scdrs_b200 import pp, synth
    from scdrs_b200.anndata_lite import AnnData

    rng = np.random.default_rng(3)
    X = synth.make_csr(400, 300, density=0.2, seed=4)
    D = X.toarray().astype(np.float64)
    w = rng.uniform(0.5, 2.0, size=400)
    for axis in (0, 1):
        m, v = pp._get_mean_var(X, axis=axis)
        assert np.allclose(m, D.mean(axis=axis), rtol=1e-6) and np.allclose(v, D.var(axis=axis), rtol=1e-5, atol=1e-9)
        m, v = pp._get_mean_var(D, axis=axis)
        assert np.allclose(m, D.mean(axis=axis), rtol=1e-6) and np.allclose(v, D.var(axis=axis), rtol=1e-5, atol=1e-9)
    m, v = pp._get_mean_var(X, axis=0, weights=w)
    wm = np.average(D, axis=0, weights=w)
    assert np.allclose(m, wm, rtol=1e-6) and np.allclose(v, np.average(D * D, axis=0, weights=w) - wm ** 2, rtol=1e-5, atol=1e-9)
    # implicit covariate correction with a random beta (tests/test_pp.py:419-458)
    adata = AnnData(X)
    cov = pd.DataFrame({"const": 1.0, "c1": rng.normal(size=400)}, index=adata.obs_names)
    beta = pd.DataFrame(rng.normal(size=(300, 2)) * 0.1, index=adata.var_names, columns=cov.columns)
    mu = pd.Series(rng.uniform(0, 1, size=300), index=adata.var_names)
    adata.uns["SCDRS_PARAM"] = {"FLAG_SPARSE": True, "FLAG_COV": True, "COV_MAT": cov, "COV_BETA": beta, "COV_GENE_MEAN": mu}
    T = D + cov.values @ beta.values.T + mu.values
    for axis in (0, 1):
        m, v = pp._get_mean_var_implicit_cov_corr(adata, axis=axis)
        assert np.allclose(m, T.mean(axis=axis), rtol=1e-8) and np.allclose(v, T.var(axis=axis), rtol=1e-6, atol=1e-10)
    m, v = pp._get_mean_var_implicit_cov_corr(adata, axis=0, transform_func=np.expm1)
    E = np.expm1(T)
    assert np.allclose(m, E.mean(axis=0), rtol=1e-8) and np.allclose(v, E.var(axis=0), rtol=1e-6)
    m, v = pp._get_mean_var_implicit_cov_corr(adata, axis=0, weights=w)
    wm = np.average(T, axis=0, weights=w)
    assert np.allclose(m, wm, rtol=1e-8) and np.allclose(v, np.average(T * T, axis=0, weights=w) - wm ** 2, rtol=1e-6, atol=1e-10)


def test_gene_moment_kernels_against_numpy(ctx):
    """scdrs_gene_moments / scdrs_gene_expm1_implicit against dense float64 numpy; several gene tiles (n_acc = 18
    accumulators per gene leave ~1400 genes per shared-memory tile), ragged rows, an empty row, weights; bitwise
    reproducible."""
    from scdrs_b200 import synth

    rng = np.random.default_rng(5)
    X = synth.make_csr(3000, 4100, density=0.08, seed=6).tolil()
    X[17, :] = 0
    X = sparse.csr_matrix(X.tocsr())
    X.eliminate_zeros()
    D = X.toarray().astype(np.float64)
    ctx.set_csr_host(X.indptr, X.indices, X.data, X.shape[1])
    w = rng.uniform(0.5, 2.0, size=3000)
    for k in (0, 4, 16):
        C = rng.normal(size=(3000, k))
        for wt in (None, w):
            ww = np.ones(3000) if wt is None else wt
            got = ctx.gene_moments(0, C if k else None, wt)
            want = np.stack([(ww[:, None] * D).sum(0), (ww[:, None] * D * D).sum(0)]
                            + [(ww[:, None] * D * C[:, [j]]).sum(0) for j in range(k)], axis=1)
            assert got.shape == want.shape and np.allclose(got, want, rtol=1e-12, atol=1e-9), (k, wt is None)
            assert np.array_equal(got, ctx.gene_moments(0, C if k else None, wt))
    E = np.where(D != 0, np.expm1(D), 0.0)
    got = ctx.gene_moments(1, None, w)
    assert np.allclose(got, np.stack([(w[:, None] * E).sum(0), (w[:, None] * E * E).sum(0)], axis=1), rtol=1e-12)
    C = np.column_stack([np.ones(3000), rng.normal(size=(3000, 3))])
    B = rng.normal(size=(4100, 4)) * 0.05
    mu = rng.uniform(0, 0.5, size=4100)
    ctx.set_cov(C, B, mu)
    V = np.expm1(D + C @ B.T + mu)
    shift = np.expm1(mu)
    for wt in (None, w):
        ww = np.ones(3000) if wt is None else wt
        got = ctx.gene_expm1_implicit(shift, wt)
        want = np.stack([(ww[:, None] * (V - shift)).sum(0), (ww[:, None] * (V - shift) ** 2).sum(0)], axis=1)
        assert np.allclose(got, want, rtol=1e-10, atol=1e-9), wt is None
        assert np.array_equal(got, ctx.gene_expm1_implicit(shift, wt))
    ctx.set_cov(None)


def test_fast_and_exact_order_statistics_agree(sc, monkeypatch):
    """The parallel moments against the sequential numpy-ordered kernels on the same data: all four preprocess modes
    plus adj_prop weights; a numerically constant gene (every cell expresses the same value) must come out of the
    fast path with the exact-order bits (it is detected and recomputed sequentially)."""
    from scdrs_b200 import synth

    for use_cov in (True, False):
        for adj in (None, "cell_type"):
            res = {}
            for exact in ("1", "0"):
                monkeypatch.setenv("SCDRS_B200_EXACT_ORDER", exact)
                adata, cov = synth.make_adata(800, 1500, density=0.12, seed=41, n_group=7)
                X = adata.X.tolil()
                X[:, 3] = 0.7                       # constant, stored everywhere
                adata.X = sparse.csr_matrix(X.tocsr())
                sc.preprocess(adata, cov=cov if use_cov else None, adj_prop=adj)
                res[exact] = adata.uns["SCDRS_PARAM"]
                sc.clear_device_cache()
            g1, g0 = res["1"]["GENE_STATS"], res["0"]["GENE_STATS"]
            for c in ["mean", "var", "ct_mean", "ct_var"]:
                a, b = g0[c].values.astype(float), g1[c].values.astype(float)
                # with covariates the sequential path repeats scipy's float32 accumulation of COV_GENE_MEAN (1e-7 at
                # 800 cells), which enters every corrected value; without, both paths sum the same doubles
                tol = 2e-6 if use_cov else 1e-9
                assert np.allclose(np.delete(a, 3), np.delete(b, 3), rtol=tol, atol=1e-12), (use_cov, adj, c)
            if adj is None and not use_cov:      # the constant gene: same bits as the sequential path
                for c in ["mean", "var", "ct_mean", "ct_var"]:
                    assert g0[c].values[3] == g1[c].values[3], (use_cov, c, g0[c].values[3], g1[c].values[3])
            if adj is None and use_cov:          # ... up to the float32 rounding of its COV_GENE_MEAN
                assert abs(g0["var"].values[3]) < 1e-12 and abs(g1["var"].values[3]) < 1e-12
            assert (g0["mean_var"].astype(str).values != g1["mean_var"].astype(str).values).sum() <= 4
            c0, c1 = res["0"]["CELL_STATS"], res["1"]["CELL_STATS"]
            assert np.allclose(c0.values.astype(float), c1.values.astype(float), rtol=2e-6 if use_cov else 1e-9, atol=1e-12)


def test_preprocess_param_schema_four_modes(sc):
    """SCDRS_PARAM keys and flags for sparse/dense x cov/nocov (tests/test_pp.py:10-112)."""
    from scdrs_b200 import synth

    for dense in (False, True):
        for use_cov in (False, True):
            adata, cov = synth.make_adata(300, 400, density=0.2, seed=2)
            if dense:
                adata.X = adata.X.toarray()
            sc.preprocess(adata, cov=cov if use_cov else None)
            p = adata.uns["SCDRS_PARAM"]
            assert p["FLAG_SPARSE"] == (not dense) and p["FLAG_COV"] == use_cov and p["FLAG_ADJ_PROP"] is False
            assert ("COV_MAT" in p) == (use_cov and not dense) and ("COV_BETA" in p) == (use_cov and not dense)
            assert list(p["CELL_STATS"].columns) == ["mean", "var"] and p["GENE_STATS"].shape == (400, 7)
            if use_cov and not dense:
                assert p["COV_BETA"].shape == (400, 4) and p["COV_GENE_MEAN"].shape == (400,)
    adata, cov = synth.make_adata(300, 400, density=0.2, seed=2)
    with pytest.raises(AssertionError, match="cov does not match"):
        sc.preprocess(adata, cov=cov.set_index(pd.Index(["x%d" % i for i in range(300)])))
    with pytest.raises(AssertionError, match="adj_prop"):
        sc.preprocess(adata, adj_prop="nope")


# ---- the two arithmetics of the raw-score scatter (spmm_score.cu / spmm_score_fx.cu) ---------------
def _score_in_mode(mode, adata, genes, w, **kw):
    import scdrs_b200
    from scdrs_b200 import _device

    _device.set_spmm_mode(mode)
    try:
        df = scdrs_b200.score_cell(adata, genes, gene_weight=w, **kw)
        kind = _device._STATES[id(adata)].ctx.last_spmm_kind()
    finally:
        _device.set_spmm_mode("auto")
    return df, kind


@pytest.mark.parametrize("n_gene,n_trait_gene,weighted", [
    (4000, 300, False),    # ~75 sets per gene: lists of two and three blocks
    (4000, 300, True),
    (400, 120, False),     # ~300 sets per gene: lists beyond the matching's 128 entries (greedy placement)
    (400, 120, True),
])
def test_fixed_point_scatter_against_fp64_and_oracle(native_lib, n_gene, n_trait_gene, weighted):
    """n_ctrl = 1000: the fixed-point kernel (integer sums, two accumulator words per set) against the
    fp64 kernel and the oracle; it must also be bitwise reproducible although the gene-major lists are
    built with atomics (integer sums do not depend on the order)."""
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    n_cell, n_ctrl = 600, 1000
    adata, cov = synth.make_adata(n_cell, n_gene, density=0.1, seed=5)
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    genes, w = synth.make_traits(adata.var_names, 1, n_trait_gene, weighted=weighted, seed=9)["trait_00"]
    kw = dict(n_ctrl=n_ctrl, return_ctrl_raw_score=True, return_ctrl_norm_score=True)
    fx, kind = _score_in_mode("fixed", adata, genes, w, **kw)
    assert kind[0] == "fixed" and kind[1] > 0
    fx2, _ = _score_in_mode("fixed", adata, genes, w, **kw)
    assert np.array_equal(fx.values, fx2.values)
    f64, kind64 = _score_in_mode("f64", adata, genes, w, **kw)
    assert kind64[0] == "f64"
    cr = ["ctrl_raw_score_%d" % i for i in range(n_ctrl)]
    cn = ["ctrl_norm_score_%d" % i for i in range(n_ctrl)]
    # quantisation: 2^-26 of the largest entry per item (an absolute error: a raw score made of one
    # small entry sees it relative to the largest) -> far inside the 1e-5 tolerance
    top = float(np.abs(f64[cr].values).max())
    assert np.allclose(fx[cr].values, f64[cr].values, rtol=1e-6, atol=2e-7 * top)
    assert np.allclose(fx["raw_score"].values, f64["raw_score"].values, rtol=1e-6, atol=2e-7 * top)
    assert np.abs(fx[cn].values - f64[cn].values).max() < 2e-5
    ref = orc.score_cell(adata, genes, gene_weight=w, **kw)
    _check_counts_self_consistent(fx, n_ctrl)
    _compare_with_reference(fx, ref, n_ctrl, n_cell)


def test_fixed_point_scatter_column_tiles_and_outlier_fallback(native_lib):
    """1 + n_ctrl > 1024 is scattered in column tiles of 1024 sets (here two) and must agree with the
    fp64 kernel and the oracle.  One stored value 1e5 times the others would cost the fixed-point scale
    17 bits: auto mode then runs the fp64 kernel (and "fixed" refuses)."""
    from scdrs_b200 import _device, synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(300, 1500, density=0.1, seed=6)
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    genes, w = synth.make_traits(adata.var_names, 1, 100, weighted=False, seed=10)["trait_00"]
    kw = dict(n_ctrl=1100, return_ctrl_raw_score=True, return_ctrl_norm_score=True)
    fx, kind = _score_in_mode("auto", adata, genes, w, **kw)
    assert kind[0] == "fixed"
    f64, kind = _score_in_mode("f64", adata, genes, w, **kw)
    assert kind[0] == "f64"
    cr = ["ctrl_raw_score_%d" % i for i in range(1100)]
    top = float(np.abs(f64[cr].values).max())
    assert np.allclose(fx[cr].values, f64[cr].values, rtol=1e-6, atol=2e-7 * top)
    _check_counts_self_consistent(fx, 1100)
    _compare_with_reference(fx, orc.score_cell(adata, genes, gene_weight=w, **kw), 1100, 300)
    df, kind = _score_in_mode("auto", adata, genes, w, n_ctrl=50)
    assert kind[0] == "fixed"
    X = adata.X.copy()
    X.data[7] *= 1e5
    adata.X = X
    _device.clear_device_cache()
    ref = orc.score_cell(adata, genes, gene_weight=w, n_ctrl=50)
    df, kind = _score_in_mode("auto", adata, genes, w, n_ctrl=50)
    assert kind[0] == "f64"
    _compare_with_reference(df, ref, 50, 300)
    with pytest.raises(RuntimeError, match="fixed-point scatter requested"):
        _score_in_mode("fixed", adata, genes, w, n_ctrl=50)


def test_score_traits_equal_score_cell(native_lib):
    """score_traits (host work of upcoming traits in worker threads, frames built off the critical
    path): every frame must equal, bit for bit, what score_cell returns for that trait alone, in
    .gs order, with binary and weighted traits mixed."""
    import scdrs_b200
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(900, 2500, density=0.1, seed=21)
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    traits = synth.make_traits(adata.var_names, 5, 150, weighted=False, seed=40)
    traits.update({"w_" + k: v for k, v in synth.make_traits(adata.var_names, 4, 120, weighted=True, seed=50).items()})
    seen = []
    res = scdrs_b200.score_traits(adata, traits, n_ctrl=200, return_ctrl_norm_score=True,
                                  on_result=lambda t, df: seen.append((t, df)))
    assert res == list(traits) and [t for t, _ in seen] == list(traits)
    for t, df in seen:
        genes, w = traits[t]
        one = scdrs_b200.score_cell(adata, genes, gene_weight=w, n_ctrl=200, return_ctrl_norm_score=True)
        assert np.array_equal(df.values, one.values, equal_nan=True), t
    # inputs uploaded again into the buffers the context already owns: same frames
    scdrs_b200.invalidate_uploads()
    again = scdrs_b200.score_traits(adata, traits, n_ctrl=200, return_ctrl_norm_score=True)
    for t, df in seen:
        assert np.array_equal(df.values, again[t].values, equal_nan=True), t
    # an error inside a worker thread surfaces in the caller
    bad = dict(traits)
    bad["broken"] = (["not_a_gene"], [1.0])
    with pytest.raises(Exception):
        scdrs_b200.score_traits(adata, bad, n_ctrl=20)
    # ... and leaves the context usable (the trait in flight was collected and dropped)
    ok = scdrs_b200.score_traits(adata, {k: traits[k] for k in list(traits)[:2]}, n_ctrl=200, return_ctrl_norm_score=True)
    for t in ok:
        assert np.array_equal(ok[t].values, again[t].values, equal_nan=True), t


def test_categorical_custom_match_key(sc):
    """A custom match key with few levels takes the reference's categorical branch (groupby instead of
    qcut, scdrs/method.py:284-285); no variance ratio then (:186-195)."""
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(500, 1200, density=0.15, seed=61)
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    gs = adata.uns["SCDRS_PARAM"]["GENE_STATS"]
    gs["mykey"] = pd.qcut(gs["mean"], 12, labels=False, duplicates="drop").astype(float)
    genes, w = synth.make_traits(adata.var_names, 1, 150, weighted=True, seed=13)["trait_00"]
    for n_ctrl in (37, 1000):
        ref = orc.score_cell(adata, genes, gene_weight=w, ctrl_match_key="mykey", n_ctrl=n_ctrl,
                             return_ctrl_raw_score=True, return_ctrl_norm_score=True)
        df = sc.score_cell(adata, genes, gene_weight=w, ctrl_match_key="mykey", n_ctrl=n_ctrl,
                           return_ctrl_raw_score=True, return_ctrl_norm_score=True)
        _check_counts_self_consistent(df, n_ctrl)
        _compare_with_reference(df, ref, n_ctrl, adata.shape[0])


def test_full_size_config4_two_kernels_agree(native_lib):
    """BASELINE config 4 size on one GPU (1,000,000 x 20,000, 2.0e9 stored entries, 16 GB of CSR): the
    fixed-point and the fp64 scatter are independent implementations (different lists, different
    arithmetic); at this size they must still agree far inside the tolerance, rank counts included."""
    import torch

    import bench
    import scdrs_b200

    free, _total = torch.cuda.mem_get_info(0)
    if free < 60 * 2 ** 30:
        pytest.skip("needs 60 GB of free device memory")

    class Args:
        n_cell, n_gene, density, n_trait, n_trait_gene = 1000000, 20000, 0.10, 1, 1000

    adata, traits, info = bench.build_workload(Args, 0, torch.device("cuda", 0))
    assert info["nnz"] > 1.9e9
    genes, w = traits["trait_00"]
    fx, kind = _score_in_mode("fixed", adata, genes, w, n_ctrl=1000)
    assert kind[0] == "fixed"
    f64, kind = _score_in_mode("f64", adata, genes, w, n_ctrl=1000)
    assert kind[0] == "f64"
    assert np.isfinite(fx["norm_score"].values).all()
    assert np.allclose(fx["raw_score"].values, f64["raw_score"].values, rtol=1e-6, atol=0)
    assert np.abs(fx["norm_score"].values - f64["norm_score"].values).max() < 1e-5
    # Monte-Carlo counts move only where a control score ties the trait score within the kernels' difference
    assert (fx["mc_pval"].values != f64["mc_pval"].values).mean() < 1e-3
    # pooled ranks: 1e9 null values, a 2e-7 shift of a score moves its rank by ~1e2 of 1e9
    rel = np.abs(fx["pval"].values.astype(np.float64) - f64["pval"].values) / f64["pval"].values
    assert rel.max() < 1e-3
    scdrs_b200.clear_device_cache()


def test_full_size_config5_shard_two_kernels_agree(native_lib):
    """One GPU's share of BASELINE config 5 (4M x 30k over 8 GPUs: 500,000 cells x 30,000 genes, 1.5e9 stored
    entries, n_ctrl = 5000 -> five column tiles of the fixed-point scatter, 2.5e9 pooled null values per shard):
    same size-independent check as for config 4."""
    import torch

    import bench
    import scdrs_b200

    free, _total = torch.cuda.mem_get_info(0)
    if free < 100 * 2 ** 30:
        pytest.skip("needs 100 GB of free device memory")

    class Args:
        n_cell, n_gene, density, n_trait, n_trait_gene = 500000, 30000, 0.10, 1, 1000

    adata, traits, info = bench.build_workload(Args, 0, torch.device("cuda", 0))
    assert info["nnz"] > 1.4e9
    genes, w = traits["trait_00"]
    fx, kind = _score_in_mode("fixed", adata, genes, w, n_ctrl=5000)
    assert kind[0] == "fixed"
    f64, kind = _score_in_mode("f64", adata, genes, w, n_ctrl=5000)
    assert kind[0] == "f64"
    assert np.isfinite(fx["norm_score"].values).all()
    assert np.allclose(fx["raw_score"].values, f64["raw_score"].values, rtol=1e-6, atol=0)
    assert np.abs(fx["norm_score"].values - f64["norm_score"].values).max() < 1e-5
    assert (fx["mc_pval"].values != f64["mc_pval"].values).mean() < 1e-3
    rel = np.abs(fx["pval"].values.astype(np.float64) - f64["pval"].values) / f64["pval"].values
    assert rel.max() < 1e-3
    scdrs_b200.clear_device_cache()


# ---- the step before the path: load_h5ad's checks, filters and normalisation on the device -----------
def _raw_counts(n_cell, n_gene, seed):
    rng = np.random.default_rng(seed)
    q = rng.gamma(0.5, 1.0, n_gene)
    q = np.minimum(0.6, 0.08 * q / q.mean())
    mask = rng.random((n_cell, n_gene)) < q
    mask[rng.random(n_cell) < 0.05] &= rng.random(n_gene) < 0.1        # some nearly empty cells
    vals = rng.poisson(2.0, size=(n_cell, n_gene)).astype(np.float32) + 1
    X = sparse.csr_matrix(np.where(mask, vals, 0).astype(np.float32))
    X.eliminate_zeros()
    return X


@pytest.mark.parametrize("flt,raw", [(True, True), (True, False), (False, True), (False, False)])
def test_preprocess_counts_matches_host_restatement(native_lib, flt, raw):
    """scdrs_preprocess_counts against oracle/h5ad_oracle.py: survivors, entry counts and column ids exact,
    float32 values within 2e-6 (device log1pf vs numpy); the processed matrix stays resident, so
    preprocess / score_cell upload nothing."""
    from oracle import h5ad_oracle
    from scdrs_b200 import _device, util
    from scdrs_b200.anndata_lite import AnnData

    X = _raw_counts(1500, 900, 3)
    obs = pd.DataFrame({"tag": np.arange(1500)}, index=["c%d" % i for i in range(1500)])
    var = pd.DataFrame({"gtag": np.arange(900)}, index=["g%d" % i for i in range(900)])
    Y, keep_c, keep_g, n_genes, n_cells, n_counts = h5ad_oracle.preprocess_counts(
        X, flag_filter_data=flt, flag_raw_count=raw, min_genes=40, min_cells=25)
    adata = util.preprocess_counts(AnnData(X.copy(), obs, var), flag_filter_data=flt, flag_raw_count=raw,
                                   min_genes=40, min_cells=25)
    assert adata.shape == Y.shape
    assert np.array_equal(adata.obs["tag"].values, np.flatnonzero(keep_c))
    assert np.array_equal(adata.var["gtag"].values, np.flatnonzero(keep_g))
    got = adata.X.tocsr()
    assert np.array_equal(got.indptr, Y.indptr) and np.array_equal(got.indices, Y.indices)
    assert np.allclose(got.data, Y.data, rtol=2e-6, atol=0)
    if flt:
        assert 0 < keep_c.sum() < 1500 and 0 < keep_g.sum() < 900
        assert np.array_equal(adata.obs["n_genes"].values, n_genes[keep_c])
        assert np.array_equal(adata.var["n_cells"].values, n_cells[keep_g])
    if raw:
        assert np.array_equal(adata.obs["n_counts"].values, n_counts[keep_c])     # integer counts: exact
    st = _device.get_state(adata)
    scdrs_mod = __import__("scdrs_b200")
    scdrs_mod.preprocess(adata)
    assert st.h2d_bytes == 0, "the matrix was uploaded again"
    _device.clear_device_cache()


def test_preprocess_counts_rejects_nan_and_negative_like_the_reference(native_lib):
    from scdrs_b200 import util
    from scdrs_b200.anndata_lite import AnnData

    X = _raw_counts(200, 150, 4)
    for bad_value, msg in ((np.nan, "NaN"), (-1.0, "negative")):
        B = X.copy()
        B.data[17] = bad_value
        with pytest.raises(ValueError, match=msg):
            util.preprocess_counts(AnnData(B), flag_filter_data=False, flag_raw_count=False)


def test_load_h5ad_toy_file_device_path(native_lib):
    from scdrs_b200 import util

    adata = util.load_h5ad(os.path.join(helpers.TOY, "toydata_mouse.h5ad"), flag_filter_data=True,
                           flag_raw_count=True, min_genes=500, min_cells=5)
    assert "n_genes" in adata.obs and "n_cells" in adata.var and "n_counts" in adata.obs
    assert np.allclose(np.expm1(adata.X.toarray()).sum(axis=1), 1e4, rtol=1e-4)
    assert (adata.obs["n_genes"] >= 500).all()


def test_pipelined_submit_collect_equals_blocking_call(native_lib):
    """scdrs_score_trait_submit / _collect (two traits in flight, pinned staging) against scdrs_score_trait:
    identical bytes, control matrices included; a third submit without a collect is refused."""
    import scdrs_b200
    from scdrs_b200 import _device, method, synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(700, 2000, density=0.1, seed=5)
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    traits = synth.make_traits(adata.var_names, 3, 100, weighted=True, seed=9)
    state = _device.get_state(adata)
    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, True)
    tab = method._gene_table(adata, state)
    ctx.set_gene_stats(tab.var, tab.var_tech)
    preps = [method._prepare_trait(adata, state, g, w, "mean_var", 60, 200, 0) for g, w in traits.values()]
    args = lambda p: (p["trait_cols"], p["trait_gw"], p["ctrl_cols"], p["ctrl_gw"], "vs", True)   # noqa: E731
    want = [ctx.score_trait(*args(p), want_ctrl_raw=True, want_ctrl_norm=True) for p in preps]
    t0 = ctx.score_trait_submit(*args(preps[0]), want_ctrl_raw=True, want_ctrl_norm=True)
    t1 = ctx.score_trait_submit(*args(preps[1]), want_ctrl_raw=True, want_ctrl_norm=True)
    with pytest.raises(RuntimeError, match="slots"):
        ctx.score_trait_submit(*args(preps[2]))
    got0 = ctx.score_trait_collect(t0)
    t2 = ctx.score_trait_submit(*args(preps[2]), want_ctrl_norm=True)
    got = [got0, ctx.score_trait_collect(t1), ctx.score_trait_collect(t2)]
    for g, w_ in zip(got, want):
        for key in ("raw_score", "norm_score", "mc_count", "ge_count", "ctrl_norm"):
            assert np.array_equal(g[key], w_[key], equal_nan=True), key
    assert np.array_equal(got[0]["ctrl_raw"], want[0]["ctrl_raw"], equal_nan=True) and got[2]["ctrl_raw"] is None
    with pytest.raises(RuntimeError, match="nothing was submitted"):
        ctx.score_trait_collect(t0)
    scdrs_b200.clear_device_cache()


def test_upload_from_pinned_host_memory_round_trips(native_lib):
    """Large uploads take two routes -- staged through pinned chunks (ordinary numpy arrays) or one direct copy
    (arrays that already live in page-locked memory): both must put the same bytes on the device."""
    import torch

    from scdrs_b200 import _native

    rng = np.random.default_rng(0)
    n_cell, n_gene, per = 9000, 3000, 1000                      # 9e6 entries: 36 MB per array, above the threshold
    indptr = np.arange(n_cell + 1, dtype=np.int64) * per
    indices = np.sort(rng.integers(0, n_gene, (n_cell, per)).astype(np.int32), axis=1).reshape(-1)
    data = rng.random(n_cell * per, dtype=np.float32)
    pinned = [torch.from_numpy(a).pin_memory() for a in (indptr, indices, data)]
    for src in ((indptr, indices, data), tuple(t.numpy() for t in pinned)):
        c = _native.Context(0)
        c.set_csr_host(src[0], src[1], src[2], n_gene)
        ip, ix, da = c.get_csr(n_cell * per)
        assert np.array_equal(ip, indptr) and np.array_equal(ix, indices) and np.array_equal(da, data)
        c.close()


def test_text_output_writes_the_same_score_files(native_lib, tmp_path):
    """score_traits(text_output=True): the control columns are formatted on the GPU (csrc/score_format.cu) and only
    copied by the writer threads.  The .full_score.gz files must hold, after gunzip, exactly the bytes of the float
    path (whose text is pandas' own, tests/test_host.py), over more traits than there are pinned slots."""
    import gzip

    import scdrs_b200
    from scdrs_b200 import _native, synth, util
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(1100, 2500, density=0.1, seed=23)
    X = adata.X.tolil()
    X[7, :] = 0                      # a cell whose scores take the zero rule
    adata.X = sparse.csr_matrix(X.tocsr())
    adata.X.eliminate_zeros()
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    traits = synth.make_traits(adata.var_names, 5, 120, weighted=True, seed=61)
    n_ctrl = 130

    def run(text, folder):
        os.makedirs(folder, exist_ok=True)
        seen = []

        def write(trait, df):
            seen.append(trait)
            path = os.path.join(folder, trait + ".full_score.gz")
            if "ctrl_norm_text" in df.attrs:
                dt = df.iloc[:, 0:6].attrs["ctrl_norm_text"]            # slicing must not copy the pinned block
                assert dt is df.attrs["ctrl_norm_text"]
                assert text and df.shape[1] == 6 and dt.array.shape == (1100, n_ctrl, 16)
                _native.write_score_gz_slots(path, df, dt.columns, dt.array)
            else:
                assert not text and df.shape[1] == 6 + n_ctrl
                util.write_score_file(path, df)

        scdrs_b200.score_traits(adata, traits, n_ctrl=n_ctrl, return_ctrl_norm_score=True, on_result=write,
                                text_output=text)
        assert seen == list(traits)

    run(True, str(tmp_path / "text"))
    run(False, str(tmp_path / "float"))
    for t in traits:
        a = gzip.open(str(tmp_path / "text" / (t + ".full_score.gz"))).read()
        b = gzip.open(str(tmp_path / "float" / (t + ".full_score.gz"))).read()
        assert a == b, t
    df = pd.read_csv(str(tmp_path / "text" / "trait_00.full_score.gz"), sep="\t", index_col=0)
    assert df.shape == (1100, 6 + n_ctrl) and list(df.columns[-2:]) == ["ctrl_norm_score_128", "ctrl_norm_score_129"]
    # the context is back in float mode afterwards
    one = scdrs_b200.score_cell(adata, *traits["trait_00"][:1], gene_weight=traits["trait_00"][1], n_ctrl=n_ctrl,
                                return_ctrl_norm_score=True)
    assert np.allclose(one.values, df.values.astype(np.float32), rtol=0, atol=0, equal_nan=True)
    scdrs_b200.clear_device_cache()


@pytest.mark.gpu
def test_score_cell_save_intermediate_against_the_reference_dumps(sc, tmp_path):
    """score_cell(save_intermediate=prefix) writes the reference's ten stage files (scdrs/method.py:507-596) from the
    device's raw scores; the returned frame is the one without the option."""
    import os

    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(here, "golden", "ref_intermediate_sparse_cov_vs.npz"))
    adata, cov, case = helpers.load_case(os.path.join(here, "golden", "ref_case_sparse_cov_vs.npz"))
    kw = dict(gene_weight=case["gene_weight"], ctrl_match_key=case["ctrl_match_key"], n_ctrl=case["n_ctrl"],
              weight_opt=case["weight_opt"])
    prefix = str(tmp_path / "dump")
    df = sc.score_cell(adata, case["gene_list"], save_intermediate=prefix, **kw)
    assert df.equals(sc.score_cell(adata, case["gene_list"], **kw)) and df.shape[1] == 6
    for key in gold.files:
        got = np.loadtxt("%s.%s.tsv.gz" % (prefix, key), delimiter="\t")
        assert got.shape == gold[key].shape, key
        assert np.allclose(got, gold[key], rtol=RTOL, atol=ATOL), key      # control raw scores pass through float32


@pytest.mark.gpu
def test_overlapped_trait_preparation_equals_the_serial_order(native_lib, monkeypatch):
    """The list building of trait i + 1 runs on a second stream under the tail of trait i (scdrs_ctx::prep_stream).
    Gene sets of very different sizes, binary and weighted, enough cells for the tails to take a while: every frame
    must equal, bit for bit, the frame of a context that does everything on one stream, run after run."""
    import scdrs_b200
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(20000, 5000, density=0.1, seed=77)
    scdrs_b200.preprocess(adata, cov=cov)
    traits = {}
    for i, (n_g, weighted) in enumerate([(40, False), (900, False), (150, True), (1200, False), (60, True), (400, False),
                                         (30, False), (700, True), (250, False), (1000, True), (80, False), (500, False)]):
        g, w = synth.make_traits(adata.var_names, 1, n_g, weighted=weighted, seed=300 + i)["trait_00"]
        traits["t%02d" % i] = (g, w)
    runs = [scdrs_b200.score_traits(adata, traits, n_ctrl=400, return_ctrl_norm_score=True) for _ in range(2)]
    monkeypatch.setenv("SCDRS_B200_PREP_OVERLAP", "0")
    scdrs_b200.clear_device_cache()                       # the next context reads the switch
    serial = scdrs_b200.score_traits(adata, traits, n_ctrl=400, return_ctrl_norm_score=True)
    for t in traits:
        for r in runs:
            assert np.array_equal(r[t].values, serial[t].values, equal_nan=True), t
    scdrs_b200.clear_device_cache()
