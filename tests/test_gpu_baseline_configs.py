"""CUDA path against the CPU oracle ON the configurations BASELINE.json names:

  config 1  bundled toy data through the `scdrs compute-score` CLI at n_ctrl = 1000
  config 2  110,096 x 20,000, n_ctrl = 1000, one binary and one weighted trait, full size
  config 3  MAGMA-z-like weights + preprocess(adj_prop=...) over 120 Zipf groups
  config 4  1,000,000 x 20,000 on one GPU: a 50,000-cell slice scored by the oracle with the full run's column
            statistics injected (raw / normalised scores, Monte-Carlo counts), pooled ranks of 1,000 cells recounted
            against all 1e9 null values
  config 5  one GPU's share (500,000 x 30,000, n_ctrl = 5000): same slice check on 10,000 cells

Integer outputs exact given the scores; floats within 1e-5 (north_star); the cells whose pooled rank differs from
the float64 oracle are counted and printed, and the size of every difference is bounded by the number of null values
that lie within the float32 rounding of the score (see _compare).
"""
import os
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

import helpers
from oracle import scdrs_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = ATOL = 1e-5       # north_star: raw, normalised and z-scores within 1e-5 in fp32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compare(df, ref, n_ctrl, n_null, what):
    for c in ["raw_score", "norm_score"]:
        assert np.allclose(df[c].values, ref[c].values, rtol=RTOL, atol=ATOL), (what, c)
    p_got, p_ref = df["pval"].values.astype(np.float64), ref["pval"].values.astype(np.float64)
    r_got, r_ref = np.round(p_got * (n_null + 1)), np.round(p_ref * (n_null + 1))
    d_rank = np.abs(r_got - r_ref)
    # Counts are exact GIVEN the scores (_counts_consistent).  Against the float64 oracle a pooled rank moves by the
    # null values that lie within the rounding of the trait score: the scores are float32 (6e-8 relative) and agree
    # with the oracle to ~1e-6, and N * pdf(t) * 1e-6 null values lie that close -- 40 of 1.1e8 around t = 0, none
    # in the tails.  So at n_ctrl = 1000 MOST cells differ by a few counts; what is bounded is the size of the move
    # (a few null densities of 2e-6) and, through it, the p-value (rtol 1e-5 like every other float).
    nrm = np.abs(ref["norm_score"].values.astype(np.float64))
    dens = n_null * np.exp(-0.5 * np.minimum(nrm, 8.0) ** 2) / np.sqrt(2 * np.pi)
    assert (d_rank <= 2 + 4e-6 * dens * (1 + nrm) + 1e-5 * r_ref).all(), (what, d_rank.max())
    assert np.allclose(p_got, p_ref, rtol=2e-5, atol=2.0 / (n_null + 1)), what
    n_diff = int((d_rank > 0).sum())
    d_mc = np.abs(df["mc_pval"].values.astype(np.float64) - ref["mc_pval"].values) * (1 + n_ctrl)
    assert d_mc.max() <= 1.01 and (d_mc > 0.5).mean() <= 1e-3, (what, (d_mc > 0.5).sum())
    same = d_rank == 0
    assert np.allclose(df["zscore"].values[same], ref["zscore"].values[same], rtol=RTOL, atol=ATOL), what
    assert np.allclose(df["nlog10_pval"].values[same], ref["nlog10_pval"].values[same], rtol=RTOL, atol=ATOL), what
    print("%s: cells whose pooled rank differs from the float64 oracle: %d / %d (max |d rank| %d of %d null values, "
          "max relative p-value difference %.2g); Monte-Carlo counts that differ: %d" % (
              what, n_diff, len(df), int(d_rank.max()), n_null, float(np.abs(p_got / p_ref - 1).max()),
              int((d_mc > 0.5).sum())))


def _counts_consistent(df, n_ctrl):
    norm = df["norm_score"].values
    ctrl = df[["ctrl_norm_score_%d" % i for i in range(n_ctrl)]].values
    mc = (ctrl >= norm[:, None]).sum(axis=1)
    assert np.array_equal(df["mc_pval"].values, ((1 + mc) / (1 + n_ctrl)).astype(np.float32))
    ge = orc.null_ge_count(norm, ctrl.reshape(-1))
    assert np.array_equal(df["pval"].values, ((ge + 1) / (ctrl.size + 1)).astype(np.float32))


_FORK_DATA = {}


def _oracle_job(job):
    genes, w, n_ctrl = job          # the data set is inherited through fork(), not pickled
    return orc.score_cell(_FORK_DATA["adata"], genes, gene_weight=w, n_ctrl=n_ctrl)


# ---- config 1 -------------------------------------------------------------------------------------
def test_config1_toy_data_through_the_cli_at_n_ctrl_1000(native_lib, tmp_path):
    """`scdrs compute-score` on the bundled toy data with the default n_ctrl = 1000 (BASELINE configs[0]) against
    the oracle scoring the same SCDRS_PARAM.  The files carry float32 in shortest round-trip text: the values
    read back are the frame's own."""
    import scdrs_b200
    from scdrs_b200 import util

    toy = helpers.TOY
    cmd = [sys.executable, os.path.join(ROOT, "bin", "scdrs"), "compute-score",
           "--h5ad-file", os.path.join(toy, "toydata_mouse.h5ad"), "--h5ad-species", "mouse",
           "--gs-file", os.path.join(toy, "toydata_mouse.gs"), "--gs-species", "mouse",
           "--cov-file", os.path.join(toy, "toydata_mouse.cov"), "--n-ctrl", "1000",
           "--flag-filter-data", "False", "--flag-raw-count", "False", "--out-folder", str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = pd.read_csv(tmp_path / "toydata_gs_mouse.score.gz", sep="\t", index_col=0)
    full = pd.read_csv(tmp_path / "toydata_gs_mouse.full_score.gz", sep="\t", index_col=0)
    assert full.shape == (30, 6 + 1000)                       # default flags: the 1000 control columns are kept
    # the same preprocessing in this process, then the oracle on that SCDRS_PARAM
    adata = util.load_h5ad(os.path.join(toy, "toydata_mouse.h5ad"), flag_filter_data=False, flag_raw_count=False)
    df_cov = pd.read_csv(os.path.join(toy, "toydata_mouse.cov"), sep="\t", index_col=0)
    scdrs_b200.preprocess(adata, cov=df_cov)
    gs = util.load_gs(os.path.join(toy, "toydata_mouse.gs"), src_species="mouse", dst_species="mouse",
                      to_intersect=adata.var_names)
    genes, w = gs["toydata_gs_mouse"]
    ref = orc.score_cell(adata, genes, gene_weight=w, n_ctrl=1000, return_ctrl_norm_score=True)
    assert list(got.index) == list(ref.index)
    _compare(got, ref, 1000, 30 * 1000, "config 1 (toy data, CLI, n_ctrl=1000)")
    cn = ["ctrl_norm_score_%d" % i for i in range(1000)]
    assert np.allclose(full[cn].values, ref[cn].values, rtol=RTOL, atol=ATOL)
    _counts_consistent(full.astype(np.float32), 1000)


# ---- config 2 -------------------------------------------------------------------------------------
def test_config2_full_size_binary_and_weighted_against_oracle(native_lib):
    """BASELINE configs[1] at full size: 110,096 cells x 20,000 genes, n_ctrl = 1000.  One binary and one weighted
    trait, the oracle scoring the same SCDRS_PARAM in two worker processes while the GPU path runs."""
    import multiprocessing as mp

    import torch

    import bench
    import scdrs_b200
    from scdrs_b200 import synth

    args = bench.parse_args(["--config", "2", "--n-trait", "1"])
    adata, traits, _ = bench.build_workload(args, 0, torch.device("cuda", 0))
    g_bin, w_bin = traits["trait_00"]
    g_wt, w_wt = synth.make_traits(adata.var_names, 1, 1000, weighted=True, seed=2000)["trait_00"]
    n_cell, n_ctrl = adata.shape[0], 1000
    _FORK_DATA["adata"] = adata
    with mp.get_context("fork").Pool(2) as pool:
        pending = pool.map_async(_oracle_job, [(g_bin, w_bin, n_ctrl), (g_wt, w_wt, n_ctrl)])
        dfs = [scdrs_b200.score_cell(adata, g, gene_weight=w, n_ctrl=n_ctrl, return_ctrl_norm_score=True)
               for g, w in ((g_bin, w_bin), (g_wt, w_wt))]
        for df in dfs:
            _counts_consistent(df, n_ctrl)
        refs = pending.get()
    _FORK_DATA.clear()
    for what, df, ref in zip(("config 2 binary", "config 2 weighted"), dfs, refs):
        _compare(df, ref, n_ctrl, n_cell * n_ctrl, what)
    scdrs_b200.clear_device_cache()


# ---- config 3 -------------------------------------------------------------------------------------
def test_config3_weighted_with_adj_prop_against_oracle(native_lib):
    """BASELINE configs[2] at 20,000 cells x 20,000 genes: preprocess(cov, adj_prop) with 120 Zipf-sized groups
    (weighted gene statistics) against the oracle's dense restatement, then a MAGMA-z-weighted trait."""
    import scdrs_b200
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(20000, 20000, density=0.1, seed=33, n_group=120)
    ref_adata = adata.copy()
    scdrs_b200.preprocess(adata, cov=cov, adj_prop="cell_type")
    orc.preprocess(ref_adata, cov=cov, adj_prop="cell_type", loess_impl=loess)
    gs, gs_ref = adata.uns["SCDRS_PARAM"]["GENE_STATS"], ref_adata.uns["SCDRS_PARAM"]["GENE_STATS"]
    assert adata.uns["SCDRS_PARAM"]["FLAG_ADJ_PROP"] is True
    # COV_GENE_MEAN of the reference (and of the oracle, which calls the same scipy code) is ACCUMULATED in float32
    # (scipy keeps the dtype of X, scdrs/pp.py:206): 3e-6 off the exact mean at 20,000 cells, growing with the
    # number of cells.  The device sums in float64 and rounds once; the tolerance below is the reference's error.
    for c in ["mean", "var", "ct_mean", "ct_var"]:
        assert np.allclose(gs[c].values.astype(float), gs_ref[c].values.astype(float), rtol=2e-5, atol=1e-10), c
    for c in ["var_tech", "ct_var_tech"]:
        assert np.allclose(gs[c].values.astype(float), gs_ref[c].values.astype(float), rtol=5e-5, atol=1e-10), c
    n_bin_diff = int((gs["mean_var"].astype(str).values != gs_ref["mean_var"].astype(str).values).sum())
    assert n_bin_diff <= 20, n_bin_diff         # 1e-6 shifts of the means move genes that sit on a bin boundary
    genes, w = synth.make_traits(adata.var_names, 1, 1000, weighted=True, seed=77)["trait_00"]
    n_ctrl = 1000
    ref = orc.score_cell(adata, genes, gene_weight=w, n_ctrl=n_ctrl)        # same SCDRS_PARAM as the CUDA path
    df = scdrs_b200.score_cell(adata, genes, gene_weight=w, n_ctrl=n_ctrl, return_ctrl_norm_score=True)
    _counts_consistent(df, n_ctrl)
    _compare(df, ref, n_ctrl, adata.shape[0] * n_ctrl, "config 3 (weighted, adj_prop)")
    scdrs_b200.clear_device_cache()


# ---- configs 4 and 5: a slice of the full run against the oracle ------------------------------------
def _slice_check(args, lo, n_slice, n_rank_cells, what):
    import torch

    import bench
    import scdrs_b200
    from scdrs_b200 import _device
    from scdrs_b200 import method as M
    from scdrs_b200.anndata_lite import AnnData
    from scdrs_b200.distributed import CudaPhases, ShardedScorer

    dev = torch.device("cuda", 0)
    adata, traits, info = bench.build_workload(args, 0, dev)
    n_cell, n_ctrl = adata.shape[0], args.n_ctrl
    genes, w = traits["trait_00"]
    state = _device.get_state(adata)
    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, True)
    prep = M._prepare_trait(adata, state, genes, w, "mean_var", n_ctrl, 200, 0)
    ctx.set_gene_stats(prep["var"], prep["var_tech"])
    scorer = ShardedScorer(CudaPhases(ctx, dev), dev, group=False)
    out = scorer.score(prep["trait_cols"], prep["trait_gw"], prep["ctrl_cols"], prep["ctrl_gw"], "vs", True,
                       want_ctrl_norm=n_rank_cells > 0)
    torch.cuda.synchronize()
    colsum = scorer._buf["colsum"].cpu().numpy()
    col2 = scorer._buf["col2"].cpu().numpy()
    colmin = scorer._buf["colmin"].cpu().numpy()
    inject = dict(raw_mean=colsum / n_cell, norm_mean=col2 / n_cell, lo=float((colmin - col2 / n_cell).min()))
    # the oracle on the slice, with the full run's column statistics
    hi = lo + n_slice
    sub = AnnData(adata.X[lo:hi].tocsr().copy(), pd.DataFrame(index=adata.obs_names[lo:hi]),
                  pd.DataFrame(index=adata.var_names))
    param = dict(adata.uns["SCDRS_PARAM"])
    param["COV_MAT"] = param["COV_MAT"].iloc[lo:hi]
    sub.uns["SCDRS_PARAM"] = param
    ref, internals = orc.score_cell(sub, genes, gene_weight=w, n_ctrl=n_ctrl, inject=inject, return_internals=True)
    assert np.array_equal(internals["ctrl_cols"], prep["ctrl_cols"])             # control sets: bit-identical
    raw, nrm = out["raw_score"][lo:hi], out["norm_score"][lo:hi]
    assert np.allclose(raw, internals["v_raw"], rtol=RTOL, atol=0), what
    assert np.allclose(nrm, internals["v_norm"], rtol=RTOL, atol=ATOL), what
    d_mc = np.abs(out["mc_count"][lo:hi].astype(np.int64) - internals["mc_count"])
    assert d_mc.max() <= 1 and (d_mc > 0).mean() <= 1e-3, (what, int((d_mc > 0).sum()))
    ctrl = out["ctrl_norm"]
    if ctrl is not None:
        assert np.allclose(ctrl[lo:hi], internals["mat_norm"], rtol=RTOL, atol=ATOL), what
        # Monte-Carlo counts: exact against the returned float32 scores on the slice
        assert np.array_equal(out["mc_count"][lo:hi], (ctrl[lo:hi] >= nrm[:, None]).sum(axis=1))
    # pooled ranks of n_rank_cells cells recounted against ALL null values, in chunks:
    # pos(v) = #{queries <= v};  v >= qs[s]  <=>  pos(v) >= right[s] = #{queries <= qs[s]}  (ties count, method.py:630)
    n_null_seen = 0
    if n_rank_cells:
        q_idx = np.arange(lo, lo + n_rank_cells)
        q = out["norm_score"][q_idx]
        order = np.argsort(q, kind="stable")
        qs = q[order]
        hist = np.zeros(len(qs) + 1, dtype=np.int64)
        flat = ctrl.reshape(-1)
        n_null_seen = flat.shape[0]
        step = 1 << 26
        for s0 in range(0, flat.shape[0], step):
            hist += np.bincount(np.searchsorted(qs, flat[s0:s0 + step], side="right"), minlength=len(qs) + 1)
        lower = np.concatenate([[0], np.cumsum(hist)])          # lower[p] = #nulls with pos < p
        ge_want = np.empty(len(qs), dtype=np.int64)
        ge_want[order] = flat.shape[0] - lower[np.searchsorted(qs, qs, side="right")]
        assert np.array_equal(out["ge_count"][q_idx], ge_want), what
    assert out["n_null_total"] == n_cell * n_ctrl
    print("%s: %d-cell slice against the oracle with injected column statistics: max |norm - ref| = %.3g, "
          "Monte-Carlo counts differing: %d; pooled ranks of %d cells recounted exactly against %d null values" % (
              what, n_slice, float(np.abs(nrm - internals["v_norm"]).max()), int((d_mc > 0).sum()), n_rank_cells,
              n_null_seen))
    scdrs_b200.clear_device_cache()


def test_config4_slice_of_the_million_cell_atlas_against_oracle(native_lib):
    import torch

    import bench

    free, _total = torch.cuda.mem_get_info(0)
    if free < 60 * 2 ** 30:
        pytest.skip("needs 60 GB of free device memory")
    args = bench.parse_args(["--config", "4", "--n-trait", "1"])
    _slice_check(args, lo=400000, n_slice=50000, n_rank_cells=1000, what="config 4 (1M cells)")


def test_config5_slice_of_one_gpus_share_against_oracle(native_lib):
    import torch

    import bench

    free, _total = torch.cuda.mem_get_info(0)
    if free < 100 * 2 ** 30:
        pytest.skip("needs 100 GB of free device memory")
    args = bench.parse_args(["--config", "5", "--n-cell", "500000", "--n-trait", "1"])
    # the 2.5e9 normalised control scores of the shard are not downloaded (10 GB): no recount of pooled ranks here
    _slice_check(args, lo=250000, n_slice=10000, n_rank_cells=0, what="config 5 shard (500k x 30k, n_ctrl=5000)")
