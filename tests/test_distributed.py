"""The cell-sharded exchange logic (scdrs_b200/distributed.py) on CPU: world_size 2, gloo backend,
with a numpy stand-in for the five device phases (test infrastructure built on the oracle's maths).
Sharded results must equal the unsharded ones: integer counts exactly, scores to rounding."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class NumpyPhases:
    """Same phase contract as scdrs_b200.distributed.CudaPhases, computed with numpy on a shard."""

    def __init__(self, adata_shard, param):
        self.adata, self.param = adata_shard, param
        self.n_cell = adata_shard.shape[0]

    def raw(self, trait_genes, trait_gw, ctrl_genes, ctrl_gw, weight_opt, use_ratio, colsum):
        from oracle import scdrs_oracle as orc

        gs = self.param["GENE_STATS"]
        X = self.adata.X.tocsc()
        sets = [np.asarray(trait_genes)] + [np.asarray(c) for c in ctrl_genes]
        gws = [np.asarray(trait_gw)] + [np.asarray(ctrl_gw)] * len(ctrl_genes)
        kw = {}
        if self.param["FLAG_COV"]:
            kw = dict(cov_mat=self.param["COV_MAT"].loc[list(self.adata.obs_names)].values,
                      cov_beta=self.param["COV_BETA"].values, gene_mean=self.param["COV_GENE_MEAN"].values)
        ws = [orc.score_weights(gs, s, g, weight_opt) for s, g in zip(sets, gws)]
        self.rawm = np.stack([orc.compute_raw_score(X, s, w, **kw) for s, w in zip(sets, ws)], axis=1)
        var = gs["var"].values.astype(float)
        ratio = np.ones(len(sets))
        if use_ratio:
            num = np.array([(var[s] * w ** 2).sum() for s, w in zip(sets, ws)])
            ratio[1:] = num[1:] / num[0]
        self.a = 1 / np.sqrt(ratio)
        colsum.copy_(torch.from_numpy(self.rawm.sum(axis=0)))

    def rowstats(self, colsum, n_total, col2, colmin):
        self.n_total = n_total
        z = (self.rawm - colsum.numpy() / n_total) * self.a
        mu, sd = z[:, 1:].mean(axis=1), z[:, 1:].std(axis=1)
        self.n = (z - mu[:, None]) / sd[:, None]
        col2.copy_(torch.from_numpy(self.n.sum(axis=0)))
        colmin.copy_(torch.from_numpy(self.n.min(axis=0)))

    def queries(self, col2, colmin, q_local):
        self.m2 = col2.numpy() / self.n_total
        self.gmin = float((colmin.numpy() - self.m2).min())
        fin = self.n - self.m2
        t = fin[:, 0].copy()
        t[self.rawm[:, 0] == 0] = self.gmin - 1e-3
        c = fin[:, 1:].copy()
        c[self.rawm[:, 1:] == 0] = self.gmin
        self.t32, self.c32 = t.astype(np.float32), c.astype(np.float32)
        q_local.copy_(torch.from_numpy(self.t32))

    def count(self, q_all, hist):
        qs = np.sort(q_all.numpy())
        pos = np.searchsorted(qs, self.c32.reshape(-1), side="right")
        hist.copy_(torch.from_numpy(np.bincount(pos, minlength=len(qs) + 1).astype(np.int64)))
        self.q_all = q_all.numpy().copy()
        self.mc = (self.c32 >= self.t32[:, None]).sum(axis=1).astype(np.int32)

    def finish(self, hist, n_query_all, offset, n_null_total, n_ctrl, want_ctrl_raw, want_ctrl_norm):
        order = np.argsort(self.q_all, kind="stable")
        prefix = np.cumsum(hist.numpy())           # prefix[p] = sum_{p' <= p} hist
        ge_sorted = n_null_total - prefix[:-1]      # for sorted query s: nulls with pos > s
        ge = np.empty(n_query_all, dtype=np.int64)
        ge[order] = ge_sorted
        return dict(raw_score=self.rawm[:, 0].copy(), norm_score=self.t32, mc_count=self.mc,
                    ge_count=ge[offset: offset + self.n_cell], ctrl_raw=None,
                    ctrl_norm=self.c32 if want_ctrl_norm else None)


    # the two halves of `finish` (the CUDA phases keep results in pinned slots; here they are simply held)
    def finish_submit(self, hist, n_query_all, offset, n_null_total, n_ctrl, want_ctrl_raw, want_ctrl_norm):
        self._held = self.finish(hist, n_query_all, offset, n_null_total, n_ctrl, want_ctrl_raw, want_ctrl_norm)
        return (0, n_ctrl, want_ctrl_raw, want_ctrl_norm)

    def collect(self, ticket):
        out, self._held = self._held, None
        return out


def _make_problem():
    from oracle import scdrs_oracle as orc
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    adata, cov = synth.make_adata(301, 900, density=0.15, seed=3)
    orc.preprocess(adata, cov=cov, loess_impl=loess)
    genes, w = synth.make_traits(adata.var_names, 1, 80, weighted=True, seed=9)["trait_00"]
    df, info = orc.score_cell(adata, genes, gene_weight=w, n_ctrl=30, return_internals=True, return_ctrl_norm_score=True)
    idx = adata.var_names.get_indexer(sorted(set(genes)))
    w_of = dict(zip(genes, w))
    tgw = np.array([w_of[g] for g in sorted(set(genes))])
    return adata, info, idx, tgw, df


def _run_rank(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from scdrs_b200.distributed import ShardedScorer

        adata, info, idx, tgw, _ = _make_problem()
        n = adata.shape[0]
        cuts = [0, 130, n] if world == 2 else [0, n]          # unequal shards on purpose
        lo, hi = cuts[rank], cuts[rank + 1]
        shard = adata[np.arange(lo, hi), :]
        scorer = ShardedScorer(NumpyPhases(shard, adata.uns["SCDRS_PARAM"]), torch.device("cpu"))
        assert scorer.n_total == n and scorer.offset == lo
        # control weights: every control position inherits the trait weight of its bin position
        out = scorer.score(idx, tgw, info["ctrl_cols"], _ctrl_gw(info), "vs", True, want_ctrl_norm=True)
        # the pipelined entry points run the same phases and collectives
        ticket = scorer.submit(idx, tgw, info["ctrl_cols"], _ctrl_gw(info), "vs", True, want_ctrl_norm=True)
        again = scorer.collect(ticket)
        assert again["n_null_total"] == out["n_null_total"]
        for k, v in out.items():
            if isinstance(v, np.ndarray):
                assert np.array_equal(v, again[k], equal_nan=True), k
        np.savez(os.path.join(out_dir, "rank%d_of%d.npz" % (rank, world)), lo=lo, hi=hi,
                 **{k: v for k, v in out.items() if isinstance(v, np.ndarray)})
    finally:
        dist.destroy_process_group()


def _ctrl_gw(info):
    # oracle keeps per-set normalised weights; recover the gene weights up to the common scale of
    # set 0 by undoing the base weight is unnecessary here: NumpyPhases re-normalises, so any vector
    # proportional to the true control gene weights gives the same scores.  Use set 0's weights
    # divided by its base weights' stand-in (weights of a control set already include the base).
    return info["ctrl_gw_raw"]


@pytest.mark.parametrize("world", [2])
def test_sharded_scoring_equals_unsharded(tmp_path, world):
    import socket

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_run_rank, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port1 = s.getsockname()[1]; s.close()
    mp.spawn(_run_rank, args=(1, port1, str(tmp_path)), nprocs=1, join=True)
    full = np.load(os.path.join(str(tmp_path), "rank0_of1.npz"))
    parts = [np.load(os.path.join(str(tmp_path), "rank%d_of%d.npz" % (r, world))) for r in range(world)]
    for key in ["ge_count", "mc_count"]:
        got = np.concatenate([p[key] for p in parts])
        assert np.array_equal(got, full[key]), key                      # integer counts: exact
    for key in ["raw_score", "norm_score", "ctrl_norm"]:
        got = np.concatenate([p[key] for p in parts])
        assert np.allclose(got, full[key], rtol=1e-6, atol=1e-6), key
    # and the unsharded stand-in agrees with the oracle's score_cell
    _, info, _, _, df = _make_problem()
    assert np.allclose(full["norm_score"], df["norm_score"].values, rtol=1e-5, atol=1e-5)
    assert np.array_equal(full["mc_count"], info["mc_count"])
    d = np.abs(full["ge_count"] - info["ge_count"])
    assert d.max() <= 2 and (d > 0).mean() < 0.02


def _run_bcast(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_ctrl, n_s_max = 7, 40
        # the product path: only the drawn sets travel (everything else is computed by every rank), straight
        # into a device buffer; two buffers alternate, so a view stays valid until the call after next
        from scdrs_b200.distributed import DrawBroadcaster

        shipper = DrawBroadcaster(torch.device("cpu"), n_ctrl * n_s_max)
        views, truths = [], []
        for i in range(5):
            src = i % world
            rng = np.random.default_rng(200 + i)
            n_sc = 25 + i
            mine = rng.integers(0, 5000, (n_ctrl, n_sc)).astype(np.int32)
            views.append(shipper.ship(mine if rank == src else None, src, n_ctrl, n_sc))
            truths.append(mine)
            assert np.array_equal(views[-1].numpy(), mine), (rank, i)
            if i >= 1:
                assert np.array_equal(views[-2].numpy(), truths[-2]), (rank, i, "previous view")
        open(os.path.join(out_dir, "bcast_ok_%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_control_sets_drawn_by_one_rank_reach_all_ranks(tmp_path):
    """score_traits(sharded=True) deals the control-set draws over the ranks and broadcasts them."""
    import socket

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_run_bcast, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "bcast_ok_%d" % r)) for r in range(2))


# ---- distributed preprocess: gene-level parameters from the cells of all ranks ------------------------------
def _pp_problem():
    from scdrs_b200 import synth

    adata, cov = synth.make_adata(420, 700, density=0.15, seed=17, n_group=6)
    cov["batch"] = np.where(np.arange(420) % 3 == 0, "a", np.where(np.arange(420) % 3 == 1, "b", "c"))   # categorical
    cov.loc[cov.index[5], "age"] = np.nan                                                                     # one missing value
    cov = cov.drop(columns=["const"])           # preprocess has to add SCDRS_CONST itself ...
    cov["n_genes"] -= cov["n_genes"].mean()     # ... because nothing left spans the constant
    return adata, cov


def _run_pp_rank(rank, world, port, out_dir, adj_prop):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pandas as pd

        from fake_device import NumpyStatsState
        from scdrs_b200 import pp
        from scdrs_b200.anndata_lite import AnnData

        pp.get_state = lambda adata, device=None: NumpyStatsState(adata)         # numpy stand-in for the device calls
        adata, cov = _pp_problem()
        bounds = [0, 150, 420] if world == 2 else [0, 420]                       # ragged shards; "c"-heavy second shard
        lo, hi = bounds[rank], bounds[rank + 1]
        shard = AnnData(adata.X[lo:hi].tocsr(), adata.obs.iloc[lo:hi].copy(), adata.var.copy())
        pp.preprocess(shard, cov=cov.iloc[lo:hi], adj_prop="cell_type" if adj_prop else None, sharded=world > 1)
        p = shard.uns["SCDRS_PARAM"]
        gs = p["GENE_STATS"]
        np.savez(os.path.join(out_dir, "pp_rank%d_of%d_%d.npz" % (rank, world, int(adj_prop))), lo=lo, hi=hi,
                 gs=gs[["mean", "var", "var_tech", "ct_mean", "ct_var", "ct_var_tech"]].values.astype(np.float64),
                 bins=np.array([str(x) for x in gs["mean_var"].values]), beta=p["COV_BETA"].values, mu=p["COV_GENE_MEAN"].values,
                 cov_cols=np.array([str(x) for x in p["COV_MAT"].columns]), cov_mat=p["COV_MAT"].values,
                 cell=p["CELL_STATS"].values.astype(np.float64))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("adj_prop", [False, True])
def test_sharded_preprocess_equals_unsharded(tmp_path, adj_prop):
    """preprocess(..., sharded=True) on two ragged shards: GENE_STATS (bins included), COV_BETA and COV_GENE_MEAN are
    those of the whole atlas and identical on both ranks; COV_MAT / CELL_STATS are the rank's rows of the unsharded
    result.  Categorical covariates whose levels differ between shards, a missing value and the cell-type weights of
    adj_prop go through the same sums over ranks.  (Device reductions answered by a numpy stand-in.)"""
    import socket

    for world in (2, 1):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        mp.spawn(_run_pp_rank, args=(world, port, str(tmp_path), adj_prop), nprocs=world, join=True)
    full = np.load(os.path.join(str(tmp_path), "pp_rank0_of1_%d.npz" % int(adj_prop)))
    parts = [np.load(os.path.join(str(tmp_path), "pp_rank%d_of2_%d.npz" % (r, int(adj_prop)))) for r in range(2)]
    for key in ("gs", "beta", "mu"):
        assert np.array_equal(parts[0][key], parts[1][key]), key                   # identical bits on every rank
        assert np.allclose(parts[0][key], full[key], rtol=1e-9, atol=1e-12), key   # and the unsharded values
    assert np.array_equal(parts[0]["bins"], parts[1]["bins"]) and np.array_equal(parts[0]["bins"], full["bins"])
    assert list(parts[0]["cov_cols"]) == list(full["cov_cols"]) and "SCDRS_CONST" in list(full["cov_cols"])
    assert np.allclose(np.concatenate([p["cov_mat"] for p in parts]), full["cov_mat"], rtol=1e-12, atol=1e-12)
    assert np.allclose(np.concatenate([p["cell"] for p in parts]), full["cell"], rtol=1e-9, atol=1e-12)


def test_preprocess_host_algebra_against_oracle(monkeypatch):
    """The identities that turn per-gene moments into the mean / variance of X + COV_MAT*COV_BETA + COV_GENE_MEAN
    (scdrs_b200/pp.py:_gene_mean_var_fast) against the oracle's dense restatement of scdrs/pp.py:353-373."""
    from fake_device import NumpyStatsState
    from oracle import scdrs_oracle as orc
    from scdrs_b200 import pp, synth
    from scdrs_b200.loess1d import loess

    monkeypatch.setattr(pp, "get_state", lambda adata, device=None: NumpyStatsState(adata))
    for adj in (None, "cell_type"):
        for use_cov in (True, False):
            adata, cov = synth.make_adata(300, 500, density=0.2, seed=23, n_group=5)
            ref = adata.copy()
            pp.preprocess(adata, cov=cov if use_cov else None, adj_prop=adj)
            orc.preprocess(ref, cov=cov if use_cov else None, adj_prop=adj, loess_impl=loess)
            g, r = adata.uns["SCDRS_PARAM"]["GENE_STATS"], ref.uns["SCDRS_PARAM"]["GENE_STATS"]
            for c in ["mean", "var", "ct_mean", "ct_var", "var_tech"]:
                assert np.allclose(g[c].values.astype(float), r[c].values.astype(float), rtol=2e-6, atol=1e-9), (adj, use_cov, c)
            assert (g["mean_var"].astype(str).values != r["mean_var"].astype(str).values).sum() <= 2
