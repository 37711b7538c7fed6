"""CPU tests of the host logic and of the C-ABI library surface (no kernel launches)."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

import helpers
from oracle import scdrs_oracle as orc
from scipy import sparse

from scdrs_b200 import _native, method, util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native_lib):
    header = open(os.path.join(ROOT, "include", "scdrs_b200.h")).read()
    declared = set(re.findall(r"\b(scdrs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    for name in declared:
        assert hasattr(native_lib, name), name
    assert native_lib.scdrs_abi_version() == 1


def test_context_creation_fails_loudly_without_gpu(native_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback|failed"):
        _native.Context(0)


@pytest.mark.parametrize("seed", [0, 1, 12345, 2 ** 32 - 1])
def test_select_ctrl_draws_equal_numpy_legacy_generator(native_lib, seed):
    rng = np.random.default_rng(3)
    sizes = rng.integers(1, 90, size=120)
    bin_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    bin_genes = rng.permutation(bin_ptr[-1]).astype(np.int32)
    ntrait = np.minimum(rng.integers(0, 4, size=120), sizes).astype(np.int32)
    n_ctrl, n_s = 50, int(ntrait.sum())
    got = _native.select_ctrl(seed, bin_ptr, bin_genes, ntrait, n_ctrl, n_s)
    rs = np.random.RandomState(seed)
    want = [[] for _ in range(n_ctrl)]
    for b in range(120):
        if ntrait[b] == 0:
            continue
        g = bin_genes[bin_ptr[b]:bin_ptr[b + 1]]
        for i in range(n_ctrl):
            want[i].extend(g[rs.permutation(len(g))[: ntrait[b]]])
    assert np.array_equal(got, np.array(want))


def test_select_ctrl_full_permutations_at_mask_boundaries(native_lib):
    """Bin sizes around every power of two (where the rejection mask of the index draw changes), whole
    permutations (k = n), many draws per bin so that generator refills fall inside a shuffle."""
    sizes = np.array([1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1000, 1024, 1025])
    bin_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    bin_genes = np.arange(bin_ptr[-1], dtype=np.int32)
    ntrait = sizes.astype(np.int32)
    n_ctrl, n_s = 9, int(ntrait.sum())
    got = _native.select_ctrl(2024, bin_ptr, bin_genes, ntrait, n_ctrl, n_s)
    rs = np.random.RandomState(2024)
    want = np.empty((n_ctrl, n_s), np.int64)
    for b, n in enumerate(sizes):
        for i in range(n_ctrl):
            want[i, bin_ptr[b]:bin_ptr[b + 1]] = bin_ptr[b] + rs.permutation(int(n))
    assert np.array_equal(got, want)


def test_select_ctrl_equals_np_random_choice_on_strings(native_lib):
    names = sorted("g%03d" % i for i in range(57))
    np.random.seed(7)
    want = [list(np.random.choice(names, size=5, replace=False)) for _ in range(3)]
    got = _native.select_ctrl(7, [0, 57], np.arange(57), [5], 3, 5)
    assert [[names[j] for j in row] for row in got] == want


def test_select_ctrl_leaves_numpy_global_generator_where_the_reference_does(native_lib):
    """The reference draws on numpy's GLOBAL generator (scdrs/method.py:96, 276, 301): whatever a caller draws from
    np.random after score_cell / _select_ctrl_geneset must be the same numbers as after the reference."""
    names = sorted("g%03d" % i for i in range(300))
    df_gene = pd.DataFrame(index=names, data={"mean": np.linspace(0, 1, 300)})
    genes = names[::7]
    np.random.seed(11)                                     # what the reference does: seed, then choice() per bin x set
    bins = pd.qcut(df_gene["mean"], q=10, labels=False, duplicates="drop")
    for b in sorted(set(bins)):
        members = sorted(set(df_gene.index[bins == b]))
        hit = sorted(set(members) & set(genes))
        for _ in range(6):
            np.random.choice(members, size=len(hit), replace=False)
    want_next = np.random.random(5)
    np.random.seed(999)                                    # some other state: the call must re-seed like the reference
    method._select_ctrl_geneset(df_gene, genes, [1.0] * len(genes), "mean", 6, 10, 11)
    assert np.array_equal(np.random.random(5), want_next)
    # and the state tuple itself against a RandomState walked through the same draws
    got, state = _native.select_ctrl(5, [0, 40, 100], np.arange(100), [3, 4], 8, 7, want_state=True)
    rs = np.random.RandomState(5)
    for n, k in ((40, 3), (60, 4)):
        for _ in range(8):
            rs.permutation(n)
    ref_state = rs.get_state()
    assert np.array_equal(state[1], ref_state[1]) and state[2] == ref_state[2]


def test_resident_matrix_stamp_detects_in_place_edits():
    """adata.X edited in place (X.data *= 2, np.log1p(out=...)) must not be scored from the stale resident copy:
    the residency key is a content stamp, not object identity (and a new-but-equal .X object keeps the copy)."""
    from scipy import sparse as sp_

    from scdrs_b200 import _device

    rng = np.random.default_rng(0)
    X = sp_.random(500, 300, density=0.1, format="csr", random_state=1, dtype=np.float32)
    fp0 = _device._x_fingerprint(X)
    assert _device._x_fingerprint(X.copy()) == fp0                    # new object, same content: still resident
    X.data *= 2
    fp1 = _device._x_fingerprint(X)
    assert fp1 != fp0
    np.log1p(X.data, out=X.data)
    assert _device._x_fingerprint(X) not in (fp0, fp1)
    X.indices[-1] = (X.indices[-1] + 1) % 300
    assert _device._x_fingerprint(X) != fp1
    D = rng.random((50, 40))
    fd = _device._x_fingerprint(D)
    D[7, 3] += 1.0
    assert _device._x_fingerprint(D) != fd


def test_select_ctrl_rejects_bad_input(native_lib):
    with pytest.raises(RuntimeError, match="trait genes"):
        _native.select_ctrl(0, [0, 3], [0, 1, 2], [5], 2, 5)


@pytest.mark.parametrize("path", helpers.golden_cases())
def test_select_ctrl_geneset_matches_reference_golden(native_lib, path):
    adata, cov, case = helpers.load_case(path)
    df_gene = adata.uns["SCDRS_PARAM"]["GENE_STATS"].copy()
    df_gene["gene"] = df_gene.index
    genes = sorted(set(case["gene_list"]) & set(df_gene.index))
    w_of = dict(zip(case["gene_list"], case["gene_weight"]))
    d_list, d_w = method._select_ctrl_geneset(df_gene, genes, [w_of[g] for g in genes],
                                              case["ctrl_match_key"], case["n_ctrl"], 200, 0)
    col = {g: i for i, g in enumerate(adata.var_names)}
    got = np.array([[col[g] for g in d_list[i]] for i in range(case["n_ctrl"])])
    assert np.array_equal(got, case["ref_ctrl_idx"])
    assert np.allclose(np.array([d_w[i] for i in range(case["n_ctrl"])]), case["ref_ctrl_weight"])


def test_select_ctrl_geneset_reference_unit_test(native_lib):
    # restates tests/test_method_score_cell_other.py:10-61 of the reference
    np.random.seed(0)
    df_gene = pd.DataFrame(index=["g%d" % i for i in range(1000)],
                           data={"mean": np.random.rand(1000), "cat": np.random.choice(["a", "b", "c"], 1000)})
    genes = ["g%d" % i for i in range(0, 100, 2)]
    weights = list(np.arange(len(genes)) + 1.0)
    for key, tol in [("mean", 0.1), ("cat", None)]:
        d_list, d_w = method._select_ctrl_geneset(df_gene, genes, weights, key, 5, 10, 0)
        o_pos, o_w = orc.select_ctrl_geneset(df_gene.assign(gene=df_gene.index),
                                             df_gene.index.get_indexer(sorted(genes)),
                                             [dict(zip(genes, weights))[g] for g in sorted(genes)], key, 5, 10, 0)
        for i in range(5):
            assert d_list[i] == list(df_gene.index[o_pos[i]])
            assert np.allclose(d_w[i], o_w[i])
            assert len(set(d_list[i])) == len(genes)
            if tol is not None:
                # every control gene sits in the same quantile bin as the trait gene it replaces
                srt = sorted(genes)
                order = np.argsort(np.searchsorted(np.sort(df_gene["mean"].values), df_gene.loc[srt, "mean"].values) // 100, kind="stable")
                m_trait = df_gene.loc[[srt[j] for j in order], "mean"].values
                assert np.all(np.abs(df_gene.loc[d_list[i], "mean"].values - m_trait) < tol + 1e-9)


def test_load_gs_formats_and_duplicates(tmp_path):
    p = tmp_path / "a.gs"
    p.write_text("TRAIT\tGENESET\nt1\tA,B,C,A\nt2\tA:1.5,B:2,A:3\n")
    d = util.load_gs(str(p))
    assert list(d) == ["t1", "t2"]
    assert d["t1"] == (["A", "B", "C"], [1.0, 1.0, 1.0])
    assert d["t2"] == (["A", "B"], [3.0, 2.0])           # later duplicate wins
    d = util.load_gs(str(p), to_intersect=["B", "C"])
    assert d["t1"][0] == ["B", "C"]
    p.write_text("TRAIT\tGENESET\nt1\tA:1,B\n")
    with pytest.raises(ValueError):
        util.load_gs(str(p))
    util.save_gs(str(tmp_path / "b.gs"), {"t": (["A", "B"], [1.0, 2.5])})
    assert util.load_gs(str(tmp_path / "b.gs"))["t"] == (["A", "B"], [1.0, 2.5])


def test_load_gs_species_mapping():
    toy = helpers.TOY
    mouse = util.load_gs(os.path.join(toy, "toydata_mouse.gs"))
    human = util.load_gs(os.path.join(toy, "toydata_human.gs"), src_species="hsapiens", dst_species="mmusculus")
    assert set(human["toydata_gs_human"][0]) <= set(util.load_homolog_mapping("human", "mouse").values())
    assert len(set(mouse["toydata_gs_mouse"][0]) & set(human["toydata_gs_human"][0])) > 50
    with pytest.raises(ValueError):
        util.convert_species_name("zebrafish")


def test_read_h5ad_toy_file():
    adata = helpers.load_toy()[0]
    assert adata.shape == (30, 2500) and adata.X.nnz == 58179 and adata.X.dtype == np.float32
    assert list(adata.obs.columns) == ["label", "cell_type", "causal_variable", "non_causal_variable", "covariate"]
    assert adata.obs["cell_type"].iloc[0] == "causal_cell" and adata.var_names[0] == "Pip4k2a"


def test_count_preprocessing_restatement_on_toy_data():
    """The host restatement of load_h5ad's filters / normalisation (oracle/h5ad_oracle.py), the checker of
    the device path (tests/test_gpu_parity.py::test_preprocess_counts_*)."""
    from oracle import h5ad_oracle
    from scdrs_b200.anndata_lite import read_h5ad

    adata = read_h5ad(os.path.join(helpers.TOY, "toydata_mouse.h5ad"))
    Y, keep_c, keep_g, n_genes, n_cells, n_counts = h5ad_oracle.preprocess_counts(
        adata.X, flag_filter_data=True, flag_raw_count=True, min_genes=500, min_cells=5)
    assert Y.shape == (keep_c.sum(), keep_g.sum()) and Y.dtype == np.float32
    assert (n_genes[keep_c] >= 500).all() and (n_cells[keep_g] >= 5).all()
    assert np.allclose(np.expm1(Y.toarray()).sum(axis=1), 1e4, rtol=1e-4)
    assert np.allclose(n_counts[keep_c], np.asarray(adata.X[keep_c][:, keep_g].sum(axis=1)).reshape(-1))
    bad = adata.X.copy().astype(np.float32)
    bad.data[3] = -1.0
    with pytest.raises(ValueError):
        h5ad_oracle.preprocess_counts(bad)


def test_fdr_bh_matches_textbook():
    p = np.array([0.01, 0.04, 0.03, 0.2, 0.5])
    want = np.array([0.05, 0.0666666667, 0.0666666667, 0.25, 0.5])
    assert np.allclose(util.fdr_bh(p), want)


def test_score_cell_argument_checks():
    import scdrs_b200
    from scdrs_b200.anndata_lite import AnnData

    adata = AnnData(np.zeros((3, 4), dtype=np.float32))
    with pytest.raises(AssertionError, match="SCDRS_PARAM"):
        scdrs_b200.score_cell(adata, ["0"])
    adata.uns["SCDRS_PARAM"] = {}
    with pytest.raises(AssertionError, match="GENE_STATS"):
        scdrs_b200.score_cell(adata, ["0"])
    adata.uns["SCDRS_PARAM"]["GENE_STATS"] = pd.DataFrame({"mean": [0.0] * 4, "var": [0.0] * 4}, index=adata.var_names)
    with pytest.raises(AssertionError, match="var_tech"):
        scdrs_b200.score_cell(adata, ["0"])
    adata.uns["SCDRS_PARAM"]["GENE_STATS"]["var_tech"] = 0.0
    with pytest.raises(AssertionError, match="ctrl_match_key"):
        scdrs_b200.score_cell(adata, ["0"])
    with pytest.raises(AssertionError, match="weight_opt"):
        scdrs_b200.score_cell(adata, ["0"], ctrl_match_key="mean", weight_opt="bogus")


def test_cli_flag_parsing():
    from scdrs_b200 import cli

    cmd, kw = cli.parse_cli(["compute-score", "--h5ad-file", "a.h5ad", "--h5ad_species=mouse", "--gs-file", "g.gs",
                             "--gs-species", "mouse", "--out-folder", "o", "--n-ctrl", "20",
                             "--flag-filter-data", "False", "--flag_raw_count=False", "--cov-file", "None"])
    assert cmd == "compute_score" and kw["n_ctrl"] == 20 and kw["flag_filter_data"] is False
    assert kw["flag_raw_count"] is False and kw["cov_file"] is None and kw["flag_return_ctrl_norm_score"] is True
    assert kw["ctrl_match_opt"] == "mean_var" and kw["weight_opt"] == "vs" and kw["min_genes"] == 250
    with pytest.raises(SystemExit):
        cli.parse_cli(["compute-score", "--bogus", "1"])
    with pytest.raises(SystemExit):
        cli.main(["plot-something", "--out-file", "x"])
    with pytest.raises(ValueError):
        cli.compute_score("a", "mouse", "g", "mouse", "o", ctrl_match_opt="bad")


def test_reg_out_matches_lstsq():
    # tests/test_pp.py:344-381 of the reference
    from scdrs_b200 import pp

    rng = np.random.default_rng(0)
    X = np.column_stack([np.ones(200), rng.normal(size=(200, 3))])
    Y = rng.normal(size=(200, 7))
    want = Y - X @ np.linalg.lstsq(X, Y, rcond=None)[0]
    assert np.allclose(pp.reg_out(Y, X), want)
    assert np.allclose(pp.reg_out(Y[:, 0], X), want[:, 0])
    from scipy import sparse
    assert np.allclose(pp.reg_out(sparse.csr_matrix(Y), sparse.csr_matrix(X)), want)


def test_category2dummy():
    from scdrs_b200 import pp

    df = pd.DataFrame({"a": [1.0, 2.0, 3.0, 4.0], "sex": ["m", "f", None, "m"], "b": ["x", "y", "z", "x"]})
    out = pp.category2dummy(df)
    assert list(out.columns) == ["a", "sex_m", "b_y", "b_z"]
    assert np.isnan(out.loc[2, "sex_m"]) and out.loc[0, "sex_m"] == 1 and out.loc[1, "sex_m"] == 0
    with pytest.raises(AssertionError):
        pp.category2dummy(df, cols=["nope"])


def test_score_file_writer_matches_pandas_bytes(native_lib, tmp_path):
    import gzip
    import io

    rng = np.random.default_rng(0)
    vals = rng.normal(size=(2500, 37)) * 10.0 ** rng.integers(-8, 9, size=(2500, 37))
    vals[0, :12] = [0.0, -0.0, np.nan, np.inf, -np.inf, 1e-4, 1e6, 999999.94, 100.0, 1e-45, 3.4028235e38, 16777216.0]
    df = pd.DataFrame(vals.astype(np.float32), index=["cell %d" % i for i in range(2500)],
                      columns=["raw_score"] + ["ctrl_norm_score_%d" % i for i in range(36)])
    df.index.name = "index"
    for nt in (1, 5):
        path = str(tmp_path / ("s%d.gz" % nt))
        util.write_score_file(path, df, n_thread=nt)
        want = io.StringIO()
        df.to_csv(want, sep="\t", index=True)
        assert gzip.open(path, "rb").read().decode() == want.getvalue()
    back = pd.read_csv(path, sep="\t", index_col=0)
    assert np.array_equal(back.values.astype(np.float32), df.values, equal_nan=True)
    # unnamed index and empty frame edge cases
    df2 = df.iloc[:0].copy(); df2.index.name = None
    util.write_score_file(str(tmp_path / "e.gz"), df2)
    want = io.StringIO(); df2.to_csv(want, sep="\t", index=True)
    assert gzip.open(str(tmp_path / "e.gz"), "rb").read().decode() == want.getvalue()
    for v in [4.741197, 6.3260064, 0.04761905, 0.0016638935, 1e-5, 123456789.0]:
        assert _native.format_f32(v) == str(np.float32(v))


def test_munge_gs_cli_known_answers(tmp_path):
    """tests/test_CLI.py:74-134 of the reference: p-value and z-score inputs, three selection rules."""
    from scdrs_b200 import cli

    pd.DataFrame({"GENE": ["OR4F5", "DAZ1", "BPY2B"], "HEIGHT": [0.02, np.nan, 0.4], "BMI": [0.8, 0.02, np.nan]}).to_csv(
        tmp_path / "pval_file.tsv", sep="\t", index=False)
    pd.DataFrame({"GENE": ["OR4F5", "DAZ1", "BPY2B"], "HEIGHT": [2.0537, np.nan, 0.25335],
                  "BMI": [-0.84162, 2.0537, np.nan]}).to_csv(tmp_path / "zscore_file.tsv", sep="\t", index=False)
    for input_file in ["pval_file", "zscore_file"]:
        for selection in [["--n-max", "1"], ["--n-min", "1", "--n-max", "3", "--fdr", "0.05"],
                          ["--n-min", "1", "--n-max", "3", "--fwer", "0.05"]]:
            out = str(tmp_path / "outfile.gs")
            cli.main(["munge-gs", "--%s" % input_file, str(tmp_path / (input_file + ".tsv")), "--out-file", out,
                      "--weight", "zscore"] + selection)
            df = pd.read_csv(out, sep="\t", index_col=0)
            assert list(df.index) == ["BMI", "HEIGHT"]
            assert df.loc["BMI", "GENESET"] == "DAZ1:2.0537"
            assert df.loc["HEIGHT", "GENESET"] == "OR4F5:2.0537"
    cli.main(["munge-gs", "--pval-file", str(tmp_path / "pval_file.tsv"), "--out-file", out, "--weight", "uniform", "--n-max", "2"])
    df = pd.read_csv(out, sep="\t", index_col=0)
    assert df.loc["HEIGHT", "GENESET"] == "OR4F5:1,BPY2B:1"
    with pytest.raises(AssertionError, match="One of --pval-file"):
        cli.main(["munge-gs", "--out-file", out])


# ---- host side of the downstream analyses (scdrs_b200/downstream.py) ------------------------------
def test_quantile_positions_reproduce_numpy():
    """The (lower index, weight) handed to the device must be numpy's own for every group size."""
    from scdrs_b200 import downstream as ds

    rng = np.random.default_rng(0)
    sizes = np.arange(1, 260)
    for q in (0.95, 0.5, 0.0, 1.0, 0.013):
        lo, gamma = ds._quantile_positions(sizes, q)
        for n, a, t in zip(sizes, lo, gamma):
            x = np.sort(rng.normal(size=n))
            hi = min(a + 1, n - 1)
            d = x[hi] - x[a]
            got = x[hi] - d * (1 - t) if t >= 0.5 else x[a] + d * t
            assert got == np.quantile(x, q), (n, q)


def test_group_layout_and_graph_equal_scipy_slicing():
    """order / grp_ptr and the within-group sub-graph against ``graph[idx][:, idx]`` per group."""
    from scipy import sparse

    from scdrs_b200 import downstream as ds

    rng = np.random.default_rng(1)
    n, n_group = 300, 7
    codes = rng.integers(0, n_group, n)
    codes[codes == 3] = 2                                    # an empty group
    G = sparse.random(n, n, density=0.05, random_state=2, format="csr", dtype=np.float32)
    order, grp_ptr = ds._group_layout(codes, n_group)
    indptr, nbr, w = ds._group_graph(G, codes, order)
    w_sum = ds._group_weight_sums((indptr, nbr, w), grp_ptr)
    for g in range(n_group):
        idx = np.nonzero(codes == g)[0]
        assert np.array_equal(order[grp_ptr[g]:grp_ptr[g + 1]], idx)
        want = G[idx][:, idx].toarray().astype(np.float64)
        a, b = grp_ptr[g], grp_ptr[g + 1]
        got = np.zeros((b - a, b - a))
        for p in range(a, b):
            for e in range(indptr[p], indptr[p + 1]):
                assert a <= nbr[e] < b
                got[p - a, nbr[e] - a] = w[e]
        assert np.array_equal(got, want)
        assert np.isclose(w_sum[g], want.sum())


def test_align_equals_sorted_intersection():
    from scdrs_b200 import downstream as ds
    from scdrs_b200.anndata_lite import AnnData
    from scipy import sparse

    names = ["c%03d" % i for i in np.random.default_rng(3).permutation(50)]
    adata = AnnData(sparse.csr_matrix((50, 2), dtype=np.float32), pd.DataFrame(index=names))
    df = pd.DataFrame({"norm_score": np.arange(40.0)}, index=names[5:35][::-1] + ["x%d" % i for i in range(10)])
    df_rows, rows = ds._align(adata, df)
    cell_list = sorted(set(adata.obs_names) & set(df.index))
    assert list(df.index[df_rows]) == cell_list and list(adata.obs_names[rows]) == cell_list


def test_load_scdrs_score_reads_the_toy_folder():
    d = util.load_scdrs_score(os.path.join(helpers.TOY, "@.full_score.gz"))
    assert list(d) == ["toydata_gs_mouse.ref_Ctrl20_CovConstCovariate"]
    df = d["toydata_gs_mouse.ref_Ctrl20_CovConstCovariate"]
    assert df.shape == (30, 26) and "ctrl_norm_score_19" in df.columns
    assert util.load_scdrs_score(os.path.join(helpers.TOY, "@.full_score.gz"), obs_names=["nobody"] * 100) == {}
    with pytest.raises(AssertionError, match="full_score.gz"):
        util.load_scdrs_score(os.path.join(helpers.TOY, "@.score.gz"))


def test_downstream_host_logic_against_committed_reference_outputs(monkeypatch):
    """The host side of the downstream analyses (everything but the three device calls, which a numpy stand-in
    answers here) reproduces what the unmodified reference returned for tests/golden/ref_downstream_case.npz and
    the reference's golden files for the toy data."""
    from fake_device import NumpyDownstreamContext

    from scdrs_b200 import downstream as ds

    fake = NumpyDownstreamContext()
    monkeypatch.setattr(ds, "_ctx", lambda: fake)
    adata, df, ref = helpers.load_downstream_case()
    grp = ds.downstream_group_analysis(adata, df, ["grp"])["grp"]
    assert list(grp.columns) == list(ref["group"].columns) and all(dt == np.float32 for dt in grp.dtypes)
    np.testing.assert_allclose(grp.loc[ref["group"].index].values.astype(float), ref["group"].values, rtol=1e-5, equal_nan=True)
    corr = ds.downstream_corr_analysis(adata, df, ["v1", "v2"])
    np.testing.assert_allclose(corr.values.astype(float), ref["corr"], rtol=1e-5)
    order = sorted(adata.obs_names)
    sub, dfo = adata[order], df.loc[order]
    for opt, key in (("control_distribution_match", "rls"), ("control", "rls_ctrl")):
        rls = ds.test_gearysc(sub, dfo, "grp", opt=opt)
        assert list(rls.index) == list(ref[key].index) and list(rls.columns) == list(ref[key].columns)
        np.testing.assert_allclose(rls.values, ref[key].values, rtol=1e-9, equal_nan=True)
    assert abs(ds.gearys_c(adata, df["norm_score"].values) - ref["gearys_c"]) < 1e-12

    toy, _, _, toy_ref, _ = helpers.load_toy()
    full = toy_ref["REF_COV_FULL"]
    gold = pd.read_csv(os.path.join(helpers.TOY, "toydata_gs_mouse.ref_Ctrl20_CovConstCovariate.scdrs_cell_corr"), sep="\t", index_col=0)
    got = ds.downstream_corr_analysis(toy, full, list(gold.index))
    np.testing.assert_allclose(got.values, gold.values, rtol=2e-6)


def test_downstream_group_labels_categorical_or_integer(monkeypatch):
    """Real AnnData annotations are usually categorical (with unused categories) or integers."""
    from fake_device import NumpyDownstreamContext

    from scdrs_b200 import downstream as ds

    fake = NumpyDownstreamContext()
    monkeypatch.setattr(ds, "_ctx", lambda: fake)
    adata, df, ref = helpers.load_downstream_case()
    want = ds.downstream_group_analysis(adata, df, ["grp"])["grp"]
    adata.obs["grp_cat"] = pd.Categorical(adata.obs["grp"], categories=["g4", "g3", "g2", "g1", "g0", "unused"])
    adata.obs["grp_int"] = np.asarray([int(g[1:]) for g in adata.obs["grp"]])
    res = ds.downstream_group_analysis(adata, df, ["grp_cat", "grp_int"])
    assert list(res["grp_cat"].index) == list(want.index) and list(res["grp_int"].index) == [0, 1, 2, 3, 4]
    for key in ("grp_cat", "grp_int"):
        np.testing.assert_allclose(res[key].values.astype(float), want.values.astype(float), equal_nan=True)
    order = sorted(adata.obs_names)
    rls = ds.test_gearysc(adata[order], df.loc[order], "grp_cat")
    np.testing.assert_allclose(rls.loc[ref["rls"].index].values, ref["rls"].values, rtol=1e-9, equal_nan=True)


def test_downstream_score_frame_layouts(monkeypatch):
    """The three ways a score frame reaches the device (whole frame + column picks, a column sub-frame when
    dtypes are mixed, gathered rows when the frame holds cells the AnnData does not) give the same statistics."""
    from fake_device import NumpyDownstreamContext

    from scdrs_b200 import downstream as ds

    fake = NumpyDownstreamContext()
    monkeypatch.setattr(ds, "_ctx", lambda: fake)
    adata, df, _ = helpers.load_downstream_case()
    want = ds.downstream_group_analysis(adata, df, ["grp"])["grp"].values.astype(float)
    want_corr = ds.downstream_corr_analysis(adata, df, ["v1", "v2"]).values.astype(float)

    mixed = df.copy()
    mixed["norm_score"] = mixed["norm_score"].astype(np.float32).astype(np.float64)      # same values ...
    mixed["pval"] = mixed["pval"].astype(np.float32)                                      # ... mixed dtypes
    mixed["norm_score"] = df["norm_score"]
    extra = pd.DataFrame(np.random.default_rng(0).normal(size=(7, df.shape[1])), columns=df.columns,
                         index=["stranger%d" % i for i in range(7)])
    bigger = pd.concat([df.iloc[:100], extra, df.iloc[100:]])
    fortran = pd.DataFrame(np.asfortranarray(df.values), index=df.index, columns=df.columns)
    for frame in (mixed, bigger, fortran):
        got = ds.downstream_group_analysis(adata, frame, ["grp"])["grp"].values.astype(float)
        tol = dict(rtol=1e-6) if frame is mixed else dict(rtol=1e-12)
        if frame is mixed:       # float32 p-values can move an FDR count at a threshold; compare the rest
            np.testing.assert_allclose(got[:, :6], want[:, :6], equal_nan=True, **tol)
        else:
            np.testing.assert_allclose(got, want, equal_nan=True, **tol)
        np.testing.assert_allclose(ds.downstream_corr_analysis(adata, frame, ["v1", "v2"]).values.astype(float),
                                   want_corr, rtol=1e-9)


def test_h5ad_writer_round_trips_through_the_reader(tmp_path):
    """scdrs_b200.h5lite.write_h5ad (synthetic benchmark inputs for the CLI) against the reader: CSR / CSC / dense X,
    numeric, boolean and string annotation columns, index names."""
    from scdrs_b200 import h5lite, synth
    from scdrs_b200.anndata_lite import read_h5ad

    adata, _ = synth.make_adata(120, 90, density=0.15, seed=2, n_group=4)
    adata.obs["n_genes"] = np.diff(adata.X.indptr)
    adata.obs["score"] = np.linspace(0, 1, 120)
    adata.var["flag"] = np.arange(90) % 3 == 0
    for kind in ("csr", "csc", "dense"):
        a = adata.copy()
        if kind == "csc":
            a.X = a.X.tocsc()
        if kind == "dense":
            a.X = a.X.toarray()
        path = str(tmp_path / ("t_%s.h5ad" % kind))
        h5lite.write_h5ad(path, a)
        b = read_h5ad(path)
        assert b.shape == a.shape and list(b.obs_names) == list(a.obs_names) and list(b.var_names) == list(a.var_names)
        X0 = a.X.toarray() if sparse.issparse(a.X) else a.X
        X1 = b.X.toarray() if sparse.issparse(b.X) else b.X
        assert np.array_equal(X0, X1) and X1.dtype == np.float32
        assert list(b.obs.columns) == ["cell_type", "n_genes", "score"]
        assert (b.obs["cell_type"].values == a.obs["cell_type"].values).all()
        assert np.array_equal(b.obs["n_genes"].values, a.obs["n_genes"].values)
        assert np.array_equal(b.obs["score"].values, a.obs["score"].values)
        assert np.array_equal(b.var["flag"].values.astype(bool), a.var["flag"].values)


def test_device_formatter_function_against_to_chars(native_lib):
    """fmt_f32_shortest -- the __host__ __device__ function the GPU formats score files with -- against the
    std::to_chars writer over a stride of float32 bit patterns plus the boundaries of the positional range, and
    the slot writer against the float writer on the same values.  (All 2^32 patterns: scdrs_selftest_format(0,
    2**32, ...) -- 40 s on 8 threads, 0 mismatches; run by hand.)"""
    import gzip

    bad = ctypes.c_uint32()
    L = _native.lib()
    for first in range(0, 2 ** 32, 2 ** 27):
        assert L.scdrs_selftest_format(first, 2 ** 17, 4, ctypes.byref(bad)) == 0, hex(bad.value)
    for v in [1e-4, 9.9999997e-05, 1e6, 999999.94, 0.0, -0.0, 1e-45, 3.4028235e38, -1.5, 123456.0, np.inf, -np.inf]:
        buf = ctypes.create_string_buffer(64)
        n = L.scdrs_format_f32_shortest(ctypes.c_float(v), buf)
        assert buf.raw[:n].decode() == _native.format_f32(v), v
    buf = ctypes.create_string_buffer(64)
    assert L.scdrs_format_f32_shortest(ctypes.c_float(np.nan), buf) == 0


def test_slot_writer_equals_float_writer(native_lib, tmp_path):
    import gzip

    rng = np.random.default_rng(2)
    n, n_lead, n_slot = 700, 6, 37
    lead = rng.normal(size=(n, n_lead)).astype(np.float32)
    tail = (rng.normal(size=(n, n_slot)) * 10.0 ** rng.integers(-7, 8, size=(n, n_slot))).astype(np.float32)
    tail[3, 5] = np.nan
    tail[4, 0] = 0.0
    tail[5, 1], tail[5, 2] = -0.000123456784, -1.23456789e-10        # the longest texts: 15 characters
    slots = np.zeros((n, n_slot, 16), dtype=np.uint8)
    for i in range(n):
        for j in range(n_slot):
            txt = _native.format_f32(tail[i, j]).encode()
            slots[i, j, :len(txt)] = np.frombuffer(txt, dtype=np.uint8)
            slots[i, j, 15] = len(txt)
    idx = ["c%d" % i for i in range(n)]
    cols = ["l%d" % j for j in range(n_lead)]
    extra = ["t%d" % j for j in range(n_slot)]
    df_lead = pd.DataFrame(lead, index=idx, columns=cols)
    df_full = pd.DataFrame(np.concatenate([lead, tail], axis=1), index=idx, columns=cols + extra)
    _native.write_score_gz_slots(str(tmp_path / "a.gz"), df_lead, extra, slots, n_thread=3)
    _native.write_score_gz(str(tmp_path / "b.gz"), df_full, n_thread=3)
    assert gzip.open(str(tmp_path / "a.gz")).read() == gzip.open(str(tmp_path / "b.gz")).read()


def test_save_intermediate_dumps_match_the_reference_files(tmp_path):
    """`save_intermediate` (scdrs/method.py:507-596): the ten stage dumps, written from the reference's raw scores and
    the host restatement of the variance ratio, against the files the unmodified reference wrote for the same case
    (tests/golden/make_golden_intermediate.py)."""
    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(here, "golden", "ref_intermediate_sparse_cov_vs.npz"))
    adata, cov, case = helpers.load_case(os.path.join(here, "golden", "ref_case_sparse_cov_vs.npz"))
    prep = method._prepare_trait(adata, None, case["gene_list"], case["gene_weight"], case["ctrl_match_key"],
                                 case["n_ctrl"], 200, 0)
    ratio = method._var_ratio_c2t(prep, case["weight_opt"], case["n_ctrl"])
    _, internals = orc.score_cell(adata, case["gene_list"], gene_weight=case["gene_weight"],
                                  ctrl_match_key=case["ctrl_match_key"], n_ctrl=case["n_ctrl"],
                                  weight_opt=case["weight_opt"], return_internals=True)
    assert np.allclose(ratio, internals["ratio"], rtol=1e-12)
    prefix = str(tmp_path / "dump")
    method._dump_intermediate(prefix, gold["raw_score"], gold["ctrl_raw_score"], ratio)
    assert len(gold.files) == 10
    for key in gold.files:
        got = np.loadtxt("%s.%s.tsv.gz" % (prefix, key), delimiter="\t")
        assert got.shape == gold[key].shape, key
        assert np.allclose(got, gold[key], rtol=1e-6, atol=1e-7), key


def test_huffman_gzip_members_round_trip_and_beat_zlib_level_1(native_lib):
    """The score-file writer's level-1 encoder (one literal-only dynamic-Huffman deflate block per gzip member,
    csrc/score_writer.cu) against the gzip module: empty input, one byte, one repeated byte, every byte value with a
    Fibonacci-skewed histogram (forces the length limit of 15 bits), random bytes, and score-file text -- where it
    must not be larger than zlib at level 1."""
    import gzip
    import zlib

    rng = np.random.default_rng(5)
    fib = [1, 1]
    while len(fib) < 40:
        fib.append(fib[-1] + fib[-2])
    skew = b"".join(bytes([i]) * min(fib[i % 40], 200000) for i in range(256))
    vals = rng.standard_normal(200000).astype(np.float32)
    text = ("\t".join(_native.format_f32(v) for v in vals[:50000]) + "\n").encode()
    cases = [b"", b"x", b"a" * 100000, skew, rng.integers(0, 256, 300000, dtype=np.uint8).tobytes(), text,
             bytes(rng.integers(48, 58, 1 << 20, dtype=np.uint8))]
    for data in cases:
        gz = _native.gzip_member(data, level=1)
        assert gzip.decompress(gz) == data
        assert gzip.decompress(gz + gz) == data + data                 # members concatenate
        assert gzip.decompress(_native.gzip_member(data, level=6)) == data
    mine = len(_native.gzip_member(text, level=1))
    ref = zlib.compressobj(1, zlib.DEFLATED, 31)
    assert mine <= len(ref.compress(text) + ref.flush())
