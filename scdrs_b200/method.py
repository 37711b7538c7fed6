"""Host side of the scoring hot path: same functions, arguments, assertions and return types as
``scdrs/method.py:12-632`` of the reference, with the arithmetic done by the sm_100a kernels in
``libscdrs_b200.so`` (no CPU fallback).

What stays on the host (Python/numpy/pandas, as in the reference): argument checks, the gene-list
intersection, binning genes by ``ctrl_match_key`` and -- in compiled host code, bit-identical to
numpy's legacy generator -- the control-gene draws; turning integer rank counts into p-values and
z-scores; building the result DataFrame.
"""
import numpy as np
import pandas as pd
import scipy as sp
import scipy.stats  # noqa: F401

from . import _native
from ._device import get_state

_LEGAL_WEIGHT_OPT = {"uniform", "vs", "inv_std", "od"}


# ----------------------------------------------------------------------------------------------
# gene bins                                                              scdrs/method.py:283-297
# ----------------------------------------------------------------------------------------------
class _GeneBins:
    """Genes grouped by ``ctrl_match_key`` in the order the reference visits them.

    bin_ptr[n_bin + 1], bin_pos[...]: row positions of df_gene, bins in pandas-groupby order
    (sorted labels, NaN dropped), genes inside a bin ordered by gene name
    (``sorted(set_of_names)``, scdrs/method.py:296); code[n_gene]: bin of each gene or -1.
    """

    def __init__(self, df_gene, ctrl_match_key, n_genebin):
        key = df_gene[ctrl_match_key]
        names = (df_gene["gene"] if "gene" in df_gene else df_gene.index.to_series()).values
        if key.unique().shape[0] < df_gene.shape[0] / 10:          # :284 categorical
            labels = key.values
        else:                                                        # :287-289 continuous
            labels = pd.qcut(key, q=n_genebin, labels=False, duplicates="drop").values
        code, uniq = pd.factorize(labels, sort=True)                 # sorted like groupby; NaN -> -1
        name_rank = np.empty(len(names), dtype=np.int64)
        name_rank[np.argsort(np.asarray(names, dtype=object), kind="stable")] = np.arange(len(names))
        keep = np.flatnonzero(code >= 0)
        order = keep[np.lexsort((name_rank[keep], code[keep]))]
        self.code = code.astype(np.int64)
        self.bin_pos = order.astype(np.int64)
        counts = np.bincount(code[keep], minlength=len(uniq))
        self.bin_ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.n_bin = len(uniq)


def _bins_for(df_gene, ctrl_match_key, n_genebin, cache=None):
    """Bins for one match key, cached per data set under a content hash of the key column."""
    if cache is None:
        return _GeneBins(df_gene, ctrl_match_key, n_genebin)
    h = pd.util.hash_pandas_object(df_gene[ctrl_match_key], index=True).values.sum()
    ck = ("bins", ctrl_match_key, n_genebin, df_gene.shape[0], int(h))
    if ck not in cache:
        for old in [k_ for k_ in cache if isinstance(k_, tuple) and k_[0] == "bins"]:
            del cache[old]
        cache[ck] = _GeneBins(df_gene, ctrl_match_key, n_genebin)
    return cache[ck]


def _ctrl_layout(bins, trait_pos, trait_gw):
    """What is known about the control sets before any draw: trait genes per bin and the weight every control
    position inherits.  Returns (ntrait[n_bin] int32, ctrl_gw[n_sc])."""
    tcode = bins.code[trait_pos]
    in_bin = tcode >= 0
    ntrait = np.bincount(tcode[in_bin], minlength=bins.n_bin).astype(np.int32)
    # the p-th control gene of a bin inherits the weight of the p-th trait gene (by name) of that bin
    order = np.argsort(tcode[in_bin], kind="stable")
    ctrl_gw = np.asarray(trait_gw, dtype=np.float64)[in_bin][order]
    return ntrait, ctrl_gw


def _draw_ctrl(bins, trait_pos, trait_gw, n_ctrl, random_seed, col_of=None):
    """Control sets as integer positions; see include/scdrs_b200.h:scdrs_select_ctrl_state.

    trait_pos must be ordered by gene name.  Returns (ctrl_pos[n_ctrl, n_sc], ctrl_gw[n_sc], rng_state):
    rng_state is numpy's legacy generator after the draws (what the reference leaves behind in ``np.random``).
    """
    ntrait, ctrl_gw = _ctrl_layout(bins, trait_pos, trait_gw)
    genes = bins.bin_pos if col_of is None else col_of[bins.bin_pos]
    ctrl, rng_state = _native.select_ctrl(random_seed, bins.bin_ptr, genes, ntrait, n_ctrl, int(ntrait.sum()),
                                          want_state=True)
    return ctrl, ctrl_gw, rng_state


# ----------------------------------------------------------------------------------------------
def score_cell(
    data,
    gene_list,
    gene_weight=None,
    ctrl_match_key="mean_var",
    n_ctrl=1000,
    n_genebin=200,
    weight_opt="vs",
    copy=False,
    return_ctrl_raw_score=False,
    return_ctrl_norm_score=False,
    random_seed=0,
    verbose=False,
    save_intermediate=None,
):
    """Score cells based on the disease gene set (drop-in for ``scdrs.method.score_cell``,
    scdrs/method.py:12-226; same parameters, same float32 DataFrame with columns
    ``raw_score, norm_score, mc_pval, pval, nlog10_pval, zscore[, ctrl_raw_score_*][, ctrl_norm_score_*]``).

    Differences from the reference, all outside the numerical contract: ``adata.X`` is NOT
    converted to CSC in place (:98-99); ``weight_opt="od"`` (benchmarking only, :352-355) raises.  The global
    numpy generator is left exactly where the reference leaves it (re-seeded, :96 / :276, then
    advanced by the control draws, :301).  ``save_intermediate`` (a file-path prefix, :507-596) writes the reference's
    ten ``.tsv.gz`` dumps of the background-correction stages: a debugging aid, derived on the host from the raw
    scores the device produced (``_dump_intermediate``); the returned frame does not depend on it.
    """
    np.random.seed(random_seed)                                       # :96
    adata = data.copy() if copy else data
    n_cell, n_gene = adata.shape

    assert (
        "SCDRS_PARAM" in adata.uns
    ), "adata.uns['SCDRS_PARAM'] not found, run `scdrs.pp.preprocess` first"
    assert (
        "GENE_STATS" in adata.uns["SCDRS_PARAM"]
    ), "adata.uns['SCDRS_PARAM']['GENE_STATS'] not found, run `scdrs.pp.preprocess` first"
    gene_stats_set_expect = {"mean", "var", "var_tech"}
    gene_stats_set = set(adata.uns["SCDRS_PARAM"]["GENE_STATS"])
    assert (
        len(gene_stats_set_expect - gene_stats_set) == 0
    ), "One of 'mean', 'var', 'var_tech' not found in adata.uns['SCDRS_PARAM']['GENE_STATS'], run `scdrs.pp.preprocess` first"
    assert ctrl_match_key in adata.uns["SCDRS_PARAM"]["GENE_STATS"], (
        "ctrl_match_key=%s not found in adata.uns['SCDRS_PARAM']['GENE_STATS']" % ctrl_match_key
    )
    assert weight_opt in _LEGAL_WEIGHT_OPT, (
        "weight_opt=%s is not one of {'uniform', 'vs', 'inv_std', 'od}'" % weight_opt
    )
    if weight_opt == "od":
        raise NotImplementedError("weight_opt='od' (benchmarking-only overdispersion score) is outside the B200 hot path")

    if verbose:
        msg = "# scdrs.method.score_cell summary:"
        msg += "\n    n_cell=%d, n_gene=%d," % (n_cell, n_gene)
        msg += "\n    n_disease_gene=%d," % len(gene_list)
        msg += "\n    n_ctrl=%d, n_genebin=%d," % (n_ctrl, n_genebin)
        msg += "\n    ctrl_match_key='%s'," % ctrl_match_key
        msg += "\n    weight_opt='%s'," % weight_opt
        msg += "\n    return_ctrl_raw_score=%s," % return_ctrl_raw_score
        msg += "\n    return_ctrl_norm_score=%s," % return_ctrl_norm_score
        msg += "\n    random_seed=%d, verbose=%s," % (random_seed, verbose)
        msg += "\n    save_intermediate=%s," % save_intermediate
        print(msg)

    param = adata.uns["SCDRS_PARAM"]
    implicit = bool(param["FLAG_SPARSE"] and param["FLAG_COV"])
    state = get_state(adata)
    prep = _prepare_trait(adata, state, gene_list, gene_weight, ctrl_match_key, n_ctrl, n_genebin, random_seed)
    np.random.set_state(prep["rng_state"])                            # the draws of :301 were made on np.random
    if verbose:
        print("# scdrs.method.score_cell: use %d overlapping genes for scoring" % len(prep["genes"]))

    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, implicit)
    ctx.set_gene_stats(prep["var"], prep["var_tech"])
    use_ratio = (ctrl_match_key == "mean_var") and (weight_opt in ("uniform", "vs", "inv_std"))  # :187
    out = ctx.score_trait(
        prep["trait_cols"], prep["trait_gw"], prep["ctrl_cols"], prep["ctrl_gw"], weight_opt, use_ratio,
        want_ctrl_raw=return_ctrl_raw_score or save_intermediate is not None, want_ctrl_norm=return_ctrl_norm_score,
    )
    if save_intermediate is not None:
        ratio = _var_ratio_c2t(prep, weight_opt, n_ctrl) if use_ratio else np.ones(n_ctrl)
        _dump_intermediate(save_intermediate, out["raw_score"], out["ctrl_raw"], ratio)
    return _result_frame(adata, out, n_ctrl, return_ctrl_raw_score, return_ctrl_norm_score)


def _var_ratio_c2t(prep, weight_opt, n_ctrl):
    """Ratio of the independent variance of every control score to that of the disease score (:186-195), on the
    host, for the ``save_intermediate`` dumps only (the device computes its own in ``prep_sets_kernel``)."""
    var = prep["var"]

    def weights(cols, gw):
        base = {"uniform": np.ones(len(cols)), "vs": 1 / np.sqrt(prep["var_tech"][cols] + 1e-2),
                "inv_std": 1 / np.sqrt(var[cols] + 1e-2)}[weight_opt] * gw
        return base / base.sum()

    w_t = weights(prep["trait_cols"], prep["trait_gw"])
    num = np.array([(var[prep["ctrl_cols"][i]] * weights(prep["ctrl_cols"][i], prep["ctrl_gw"]) ** 2).sum()
                    for i in range(n_ctrl)])
    return num / (var[prep["trait_cols"]] * w_t ** 2).sum()


def _dump_intermediate(prefix, v_raw_score, mat_ctrl_raw_score, v_var_ratio_c2t):
    """The reference's ``save_intermediate`` files (:507-596): ``<prefix>.raw_score<stage>.tsv.gz`` and
    ``<prefix>.ctrl_raw_score<stage>.tsv.gz`` for the raw scores and the four stages of the background correction,
    ``%.9e``, tab separated.  Host numpy in float64 -- a debugging side output, not the scoring path."""
    v = np.asarray(v_raw_score, dtype=np.float64).copy()
    m = np.asarray(mat_ctrl_raw_score, dtype=np.float64).copy()
    zero_v, zero_m = v == 0, m == 0

    def dump(stage):
        for name, arr in (("raw_score", v), ("ctrl_raw_score", m)):
            np.savetxt("%s.%s%s.tsv.gz" % (prefix, name, stage), arr, fmt="%.9e", delimiter="\t")

    dump("")
    v = v - v.mean()                                                    # first gene-set alignment
    m = (m - m.mean(axis=0)) / np.sqrt(np.asarray(v_var_ratio_c2t, dtype=np.float64))
    dump(".1st_gs_alignment")
    mu, sd = m.mean(axis=1), m.std(axis=1)                              # cell-wise standardisation
    v = (v - mu) / sd
    m = ((m.T - mu) / sd).T
    dump(".cellwise_standardization")
    v = v - v.mean()                                                    # second gene-set alignment
    m = m - m.mean(axis=0)
    dump(".2nd_gs_alignment")
    floor = min(v.min(), m.min()) if m.size else v.min()                # zero raw scores get the smallest value
    v[zero_v] = floor - 1e-3
    m[zero_m] = floor
    dump(".final")


class _GeneTable:
    """Per-dataset lookup tables derived from GENE_STATS / var_names (scdrs/method.py:146-148)."""

    def __init__(self, adata):
        param = adata.uns["SCDRS_PARAM"]
        df_gene = param["GENE_STATS"].loc[adata.var_names].copy()       # :146
        df_gene["gene"] = df_gene.index
        df_gene.drop_duplicates(subset="gene", inplace=True)            # :148
        names = [str(x) for x in df_gene["gene"].values]
        self.df_gene = df_gene
        self.col_of = pd.Index(adata.var_names).get_indexer(df_gene["gene"].values).astype(np.int64)
        self.pos_of = {g: i for i, g in enumerate(names)}
        n_gene = adata.shape[1]
        self.var = np.zeros(n_gene)
        self.var_tech = np.zeros(n_gene)
        self.var[self.col_of] = df_gene["var"].values.astype(np.float64)
        self.var_tech[self.col_of] = df_gene["var_tech"].values.astype(np.float64)
        self.stamp = _gene_table_stamp(adata)


def _gene_table_stamp(adata):
    """Cheap content stamp: users do edit GENE_STATS in place (docs/faq.rst:84-87 of the reference)."""
    gs = adata.uns["SCDRS_PARAM"]["GENE_STATS"]
    return (id(gs), len(adata.var_names), gs.shape, tuple(gs.columns),
            int(pd.util.hash_pandas_object(gs, index=True).values.sum()))


def _gene_table(adata, state):
    cache = state.bins_cache if state is not None else None
    if cache is not None:
        tab = cache.get("gene_table")
        if tab is not None and tab.stamp == _gene_table_stamp(adata):
            return tab
    tab = _GeneTable(adata)
    if cache is not None:
        cache["gene_table"] = tab
    return tab


def _prepare_trait(adata, state, gene_list, gene_weight, ctrl_match_key, n_ctrl, n_genebin, random_seed,
                   tab=None, bins=None, draw=True):
    """Host work of one trait: gene intersection (:146-159) and control draws (:168).

    ``tab`` / ``bins`` let a multi-trait caller compute the per-dataset tables once.  ``draw=False`` leaves the
    draws out (``ctrl_cols`` is None): what a rank needs when another rank draws for it."""
    if tab is None:
        tab = _gene_table(adata, state)
    gene_list = list(gene_list)
    gene_weight = list(gene_weight) if gene_weight is not None else [1] * len(gene_list)
    dic_gene_weight = {x: y for x, y in zip(gene_list, gene_weight)}    # :157
    genes = sorted(g for g in dic_gene_weight if g in tab.pos_of)       # :158 sorted(set & set)
    trait_pos = np.fromiter((tab.pos_of[g] for g in genes), dtype=np.int64, count=len(genes))
    trait_gw = np.array([dic_gene_weight[x] for x in genes], dtype=np.float64)
    if bins is None:
        bins = _bins_for(tab.df_gene, ctrl_match_key, n_genebin, state.bins_cache if state is not None else None)
    if draw:
        ctrl_cols, ctrl_gw, rng_state = _draw_ctrl(bins, trait_pos, trait_gw, n_ctrl, random_seed, col_of=tab.col_of)
    else:
        ctrl_cols, ctrl_gw, rng_state = None, _ctrl_layout(bins, trait_pos, trait_gw)[1], None
    return dict(genes=genes, trait_cols=tab.col_of[trait_pos].astype(np.int32), trait_gw=trait_gw,
                ctrl_cols=ctrl_cols, ctrl_gw=ctrl_gw, var=tab.var, var_tech=tab.var_tech, rng_state=rng_state)


def _result_frame(adata, out, n_ctrl, return_ctrl_raw_score, return_ctrl_norm_score):
    """Integer counts -> p-values / z-scores -> float32 DataFrame (:205-226)."""
    n_cell = adata.shape[0]
    n_null = out.get("n_null_total", n_cell * n_ctrl)
    mc_p = (1 + out["mc_count"].astype(np.float64)) / (1 + n_ctrl)            # :205
    pooled_p = (out["ge_count"].astype(np.float64) + 1) / (n_null + 1)        # :631
    with np.errstate(divide="ignore"):
        nlog10_pooled_p = -np.log10(pooled_p)                                 # :207
    pooled_z = -sp.stats.norm.ppf(pooled_p).clip(min=-10, max=10)             # :208
    # one float32 block [n_cell x (6 [+ n_ctrl] [+ n_ctrl])]: the frame is a view of it, and so is what the score
    # file writer reads (pd.concat of separate frames would copy the control matrix twice more)
    cols = ["raw_score", "norm_score", "mc_pval", "pval", "nlog10_pval", "zscore"]
    n_extra = (n_ctrl if return_ctrl_raw_score else 0) + (n_ctrl if return_ctrl_norm_score else 0)
    full = np.empty((n_cell, 6 + n_extra), dtype=np.float32)
    for j, v in enumerate((out["raw_score"], out["norm_score"], mc_p, pooled_p, nlog10_pooled_p, pooled_z)):
        full[:, j] = v
    o = 6
    if return_ctrl_raw_score:
        full[:, o:o + n_ctrl] = out["ctrl_raw"]
        cols += ["ctrl_raw_score_%d" % i for i in range(n_ctrl)]
        o += n_ctrl
    if return_ctrl_norm_score:
        full[:, o:o + n_ctrl] = out["ctrl_norm"]
        cols += ["ctrl_norm_score_%d" % i for i in range(n_ctrl)]
    return pd.DataFrame(full, index=adata.obs.index, columns=cols, copy=False)


# ----------------------------------------------------------------------------------------------
def _select_ctrl_geneset(input_df_gene, gene_list, gene_weight, ctrl_match_key, n_ctrl, n_genebin, random_seed):
    """Matched control gene sets (drop-in for scdrs/method.py:229-309).

    Returns (dic_ctrl_list, dic_ctrl_weight): control set index -> list of gene names / weights.
    Bit-identical to the reference's numpy draws.
    """
    np.random.seed(random_seed)                                           # :276
    df_gene = input_df_gene.copy()
    if "gene" not in df_gene:
        df_gene["gene"] = df_gene.index
    names = df_gene["gene"].values
    dic_gene_weight = {x: y for x, y in zip(gene_list, gene_weight)}
    pos_index = pd.Index(names)
    genes = sorted(set(gene_list) & set(names))
    trait_pos = pos_index.get_indexer(genes).astype(np.int64)
    trait_gw = np.array([dic_gene_weight[x] for x in genes], dtype=np.float64)
    bins = _GeneBins(df_gene, ctrl_match_key, n_genebin)
    ctrl_pos, ctrl_gw, rng_state = _draw_ctrl(bins, trait_pos, trait_gw, n_ctrl, random_seed)
    np.random.set_state(rng_state)                                        # :301 draws on the global generator
    # weights come back as the caller's own objects, like the reference's lists
    w_list = [dic_gene_weight[genes[i]] for i in
              np.flatnonzero(bins.code[trait_pos] >= 0)[np.argsort(bins.code[trait_pos][bins.code[trait_pos] >= 0], kind="stable")]]
    dic_ctrl_list = {i: list(names[ctrl_pos[i]]) for i in range(n_ctrl)}
    dic_ctrl_weight = {i: list(w_list) for i in range(n_ctrl)}
    return dic_ctrl_list, dic_ctrl_weight


def _compute_raw_score(adata, gene_list, gene_weight, weight_opt):
    """Raw score of one gene set (drop-in for scdrs/method.py:312-402).

    Returns (v_raw_score[n_cell] float64, v_score_weight[n_gene_in_list] float64).
    """
    gene_list = list(gene_list)
    gene_weight = list(gene_weight)
    assert weight_opt in _LEGAL_WEIGHT_OPT, (
        "weight_opt=%s is not one of {'uniform', 'vs', 'inv_std', 'od}'" % weight_opt
    )
    if weight_opt == "od":
        raise NotImplementedError("weight_opt='od' (benchmarking-only overdispersion score) is outside the B200 hot path")
    assert (
        "SCDRS_PARAM" in adata.uns
    ), "adata.uns['SCDRS_PARAM'] not found, run `scdrs.pp.preprocess` first"
    param = adata.uns["SCDRS_PARAM"]
    df_gene = param["GENE_STATS"]
    if weight_opt == "uniform":
        v_score_weight = np.ones(len(gene_list))
    if weight_opt == "vs":
        v_score_weight = 1 / np.sqrt(df_gene.loc[gene_list, "var_tech"].values.astype(np.float64) + 1e-2)
    if weight_opt == "inv_std":
        v_score_weight = 1 / np.sqrt(df_gene.loc[gene_list, "var"].values.astype(np.float64) + 1e-2)
    if gene_weight is not None:
        v_score_weight = v_score_weight * np.array(gene_weight)
    v_score_weight = v_score_weight / v_score_weight.sum()

    gene_idx = adata.var_names.get_indexer(gene_list)
    keep = gene_idx >= 0
    state = get_state(adata)
    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, bool(param["FLAG_SPARSE"] and param["FLAG_COV"]))
    v_raw_score = ctx.compute_raw_score(gene_idx[keep], v_score_weight[keep])[:, 0]
    return v_raw_score, v_score_weight


def _correct_background(v_raw_score, mat_ctrl_raw_score, v_var_ratio_c2t, save_intermediate=None):
    """Cell-wise and gene-set-wise background correction (drop-in for scdrs/method.py:482-598).

    The device computes in float64 and stores float32, so results carry float32 resolution.
    """
    if save_intermediate is not None:
        _dump_intermediate(save_intermediate, v_raw_score, mat_ctrl_raw_score, v_var_ratio_c2t)
    return _scratch_context().correct_background(v_raw_score, mat_ctrl_raw_score, v_var_ratio_c2t)


def _get_p_from_empi_null(v_t, v_t_null):
    """p = (1 + #{null >= t}) / (1 + N) (drop-in for scdrs/method.py:601-632); exact counts."""
    v_t = np.array(v_t)
    v_t_null = np.array(v_t_null)
    ge = _scratch_context().empi_null_count(v_t, v_t_null)
    return (ge.astype(np.float64) + 1) / (v_t_null.shape[0] + 1)


_SCRATCH = {}


def _scratch_context():
    from ._device import _DEFAULT_DEVICE

    dev = _DEFAULT_DEVICE[0]
    if dev not in _SCRATCH:
        _SCRATCH[dev] = _native.Context(dev)
    return _SCRATCH[dev]


# ----------------------------------------------------------------------------------------------
# many traits against one data set (what `scdrs compute-score` loops over, bin/scdrs:255-274)
# ----------------------------------------------------------------------------------------------
class DeviceText:
    """The control columns of one trait as text slots in pinned memory (``score_traits(text_output=True)``).  pandas
    deep-copies ``DataFrame.attrs`` whenever a frame is sliced; this handle copies as itself, so the 1.8 GB block
    behind it is never duplicated.  ``array``: uint8 [n_cell, n_ctrl, 16]; ``columns``: the column names."""

    def __init__(self, array, columns):
        self.array, self.columns = array, columns

    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self


def _shard_layout(sharded, shard):
    """(mode, world, rank, n_trait_group, trait_group, cell_group) of this process.

    mode "single": one process, one GPU.  "cells": the ranks of `cell_group` hold the cell shards of one atlas.
    "traits": every rank holds all cells and scores traits[rank::world].  A grid (t, c) combines both: rank =
    trait_group * c + cell_shard; the c ranks of a trait group hold the c cell shards and score
    traits[trait_group::t]."""
    if shard is None:
        shard = "cells" if sharded else None
    if shard is None:
        return "single", 1, 0, 1, 0, None
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return "single", 1, 0, 1, 0, None
    world, rank = dist.get_world_size(), dist.get_rank()
    if shard == "cells":
        return "cells", world, rank, 1, 0, None
    if shard == "traits":
        return "traits", world, rank, world, rank, None
    n_tg, n_cs = (int(x) for x in shard)
    assert n_tg * n_cs == world, "shard grid %d x %d does not match the %d ranks" % (n_tg, n_cs, world)
    if n_cs == 1:
        return "traits", world, rank, world, rank, None
    if n_tg == 1:
        return "cells", world, rank, 1, 0, None
    mine = None
    for tg in range(n_tg):              # every rank creates every group, in the same order
        grp = dist.new_group(ranks=list(range(tg * n_cs, (tg + 1) * n_cs)))
        if tg == rank // n_cs:
            mine = grp
    return "cells", world, rank, n_tg, rank // n_cs, mine


def _check_ranks_agree(adata, tab, ctrl_match_key, group, device):
    """Cell-sharded scoring needs the SAME gene-level parameters on every rank (same bins -> same control sets,
    same weights); a rank that preprocessed its own shard alone would silently score against different sets."""
    import torch
    import torch.distributed as dist

    gs = tab.df_gene
    h = int(pd.util.hash_pandas_object(gs[["var", "var_tech", ctrl_match_key]], index=True).values.sum()) & ((1 << 62) - 1)
    param = adata.uns["SCDRS_PARAM"]
    if "COV_BETA" in param:
        h ^= int(pd.util.hash_pandas_object(param["COV_BETA"], index=True).values.sum()) & ((1 << 62) - 1)
    t = torch.tensor([h, -h], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    if int(t[0].item()) != -int(t[1].item()):
        raise RuntimeError(
            "score_traits(sharded): GENE_STATS / COV_BETA differ between ranks -- run preprocess(..., sharded=True) "
            "(statistics over the cells of all ranks) or give every rank the same SCDRS_PARAM gene tables")


def score_traits(
    data,
    dict_gs,
    ctrl_match_key="mean_var",
    n_ctrl=1000,
    n_genebin=200,
    weight_opt="vs",
    return_ctrl_raw_score=False,
    return_ctrl_norm_score=False,
    random_seed=0,
    min_genes=0,
    n_jobs=None,
    on_result=None,
    sharded=False,
    shard=None,
    max_prefetch=None,
    text_output=False,
):
    """Score every gene set of ``dict_gs`` = {trait: (gene_list, gene_weight_list)}.

    One process per GPU under torchrun (``torch.distributed`` initialised), three ways to split the work:

    * ``shard="cells"`` (= ``sharded=True``): ``data`` holds this rank's block of cells; the gene-set alignment
      statistics and the pooled empirical null span the cells of all ranks
      (``scdrs_b200.distributed.ShardedScorer``), the control-set draws are dealt over the ranks and broadcast, and
      the returned frames cover this rank's cells.  The gene-level parameters must be those of the whole atlas
      (``preprocess(..., sharded=True)``); this is checked.
    * ``shard="traits"``: every rank holds ALL cells and scores ``traits[rank::world]`` -- the reference's own
      parallelism (``.gs`` batches over array jobs, bin/scdrs:255-274), no collective at all.
    * ``shard=(t, c)``: a grid of t trait groups x c cell shards (rank = trait_group * c + cell_shard).

    Each trait gets exactly what ``score_cell`` would return for it (every call re-seeds, so traits
    are independent, scdrs/method.py:96, :276); the host work of upcoming traits (gene
    intersection and the control-gene draws, which release the GIL) runs in worker threads, at most
    ``max_prefetch`` traits ahead, while the GPU scores the current one; result frames are built and handed to
    ``on_result(trait, df)`` (in ``.gs`` order) off the critical path.  Returns {trait: DataFrame}, or the list of
    scored traits when a callback is given.

    ``text_output=True`` (needs ``on_result`` and ``return_ctrl_norm_score``; one GPU per process): the
    ``ctrl_norm_score_*`` columns do not come back as floats.  The GPU turns them into the text of the score file
    (``csrc/score_format.cu``) and the frame handed to ``on_result`` holds the six leading columns, with
    ``df.attrs["ctrl_norm_text"]`` = a ``DeviceText`` (``.array``: uint8 ``[n_cell, n_ctrl, 16]`` text slots in
    pinned memory, valid until ``on_result`` returns; ``.columns``) for ``_native.write_score_gz_slots``.
    """
    import os
    import time
    from concurrent.futures import ThreadPoolExecutor

    trace = [] if os.environ.get("SCDRS_B200_TRACE", "0") == "1" else None
    t_start = time.perf_counter()

    def mark(what):
        if trace is not None:
            trace.append((what, time.perf_counter() - t_start))

    adata = data
    assert "SCDRS_PARAM" in adata.uns, "adata.uns['SCDRS_PARAM'] not found, run `scdrs.pp.preprocess` first"
    assert weight_opt in _LEGAL_WEIGHT_OPT, (
        "weight_opt=%s is not one of {'uniform', 'vs', 'inv_std', 'od}'" % weight_opt
    )
    if weight_opt == "od":
        raise NotImplementedError("weight_opt='od' (benchmarking-only overdispersion score) is outside the B200 hot path")
    param = adata.uns["SCDRS_PARAM"]
    assert ctrl_match_key in param["GENE_STATS"], (
        "ctrl_match_key=%s not found in adata.uns['SCDRS_PARAM']['GENE_STATS']" % ctrl_match_key
    )
    implicit = bool(param["FLAG_SPARSE"] and param["FLAG_COV"])
    state = get_state(adata)
    use_ratio = (ctrl_match_key == "mean_var") and (weight_opt in ("uniform", "vs", "inv_std"))
    mode, world, rank, n_tg, my_tg, cell_group = _shard_layout(sharded, shard)
    text_output = bool(text_output and return_ctrl_norm_score and n_ctrl > 0)
    if text_output:
        assert on_result is not None, "text_output hands pinned text to on_result; it needs the callback"
        assert mode != "cells", "text_output is a single-GPU path (score files of sharded runs are written per rank from floats)"
    all_traits = [t for t in dict_gs if len(dict_gs[t][0]) >= min_genes]
    traits = all_traits[my_tg::n_tg] if n_tg > 1 else all_traits
    if not traits and mode != "cells":
        return {} if on_result is None else []
    # per-dataset tables once, outside the pool
    tab = _gene_table(adata, state)
    bins = _GeneBins(tab.df_gene, ctrl_match_key, n_genebin)
    n_jobs = n_jobs or max(1, min(len(traits) or 1, (os.cpu_count() or 2) - 1, 16))
    window = max(2, max_prefetch if max_prefetch is not None else 2 * n_jobs)
    results = {}
    last_rng_state = [None]
    cell_sharded = mode == "cells"
    with ThreadPoolExecutor(max_workers=n_jobs) as pool, ThreadPoolExecutor(max_workers=2) as frame_pool, \
            ThreadPoolExecutor(max_workers=1) as writer:
        if cell_sharded:
            import torch
            import torch.distributed as dist

            g_world, g_rank = dist.get_world_size(cell_group), dist.get_rank(cell_group)
        else:
            g_world, g_rank = 1, 0
        # Every rank of a cell group needs the same control sets; the draws (~0.1 s of one core per trait) are
        # dealt round-robin over the group's ranks and broadcast, instead of every rank drawing all of them.
        share_draws = g_world > 1 and os.environ.get("SCDRS_B200_LOCAL_DRAWS", "0") != "1"
        futs = {}
        n_submitted = [0]

        def prefetch(upto):
            # host work (intersection + control draws) of the traits [n_submitted, upto): a bounded window, so a
            # .gs file with thousands of sets does not pile up drawn control matrices (4 MB each at the defaults)
            while n_submitted[0] < min(upto, len(traits)):
                i_t = n_submitted[0]
                g, w = dict_gs[traits[i_t]]
                futs[i_t] = pool.submit(_prepare_trait, adata, state, g, w, ctrl_match_key, n_ctrl, n_genebin,
                                        random_seed, tab, bins, not (share_draws and i_t % g_world != g_rank))
                n_submitted[0] += 1

        prefetch(window)          # the draws start now; the upload of X overlaps with the first of them
        mark("tables+submit")
        ctx = state.ensure_matrix(adata)
        mark("matrix resident")
        state.ensure_cov(adata, implicit)
        ctx.set_gene_stats(tab.var, tab.var_tech)
        ctx.set_text_output(text_output)
        mark("cov+stats resident")
        text_writes = {}      # trait number -> future of the writer that still reads the trait's pinned text
        if cell_sharded:
            from .distributed import CudaPhases, DrawBroadcaster, ShardedScorer

            dev = torch.device("cuda", torch.cuda.current_device())
            _check_ranks_agree(adata, tab, ctrl_match_key, cell_group, dev)
            scorer = ShardedScorer(CudaPhases(ctx, dev), dev, group=cell_group)
            if share_draws and traits:
                shipper = DrawBroadcaster(dev, n_ctrl * max(len(dict_gs[t][0]) for t in traits), group=cell_group)
        pending = []          # (trait, future of the result frame): frames are built off the critical path
        written = []          # futures of on_result calls (one writer thread: .gs order is kept)
        in_flight = None      # (trait, ticket) submitted to the GPU and not collected yet

        def collect(item, i_prev=None):
            t_prev, ticket = item
            out_prev = scorer.collect(ticket) if cell_sharded else ctx.score_trait_collect(ticket)
            if text_output:
                # the six leading columns are cheap: build the frame here and hand it, with the pinned text of the
                # control columns, straight to the writer thread; the slot is taken again two submits from now
                df_prev = _result_frame(adata, out_prev, n_ctrl, return_ctrl_raw_score, False)
                df_prev.attrs["ctrl_norm_text"] = DeviceText(out_prev["ctrl_norm_text"],
                                                             ["ctrl_norm_score_%d" % i for i in range(n_ctrl)])
                fut = writer.submit(on_result, t_prev, df_prev)
                text_writes[i_prev] = fut
                written.append(fut)
                return
            pending.append((t_prev, frame_pool.submit(_result_frame, adata, out_prev, n_ctrl, return_ctrl_raw_score,
                                                      return_ctrl_norm_score)))
            drain(2)

        def drain(limit):
            while len(pending) > limit:
                t_done, fut = pending.pop(0)
                df_done = fut.result()
                if on_result is None:
                    results[t_done] = df_done
                else:
                    written.append(writer.submit(on_result, t_done, df_done))
                    while len(written) > 4:          # back-pressure: frames are large (n_cell x 1006 by default)
                        written.pop(0).result()

        def run_all():
            nonlocal in_flight
            for i_t, t in enumerate(traits):
                prefetch(i_t + 1 + window)
                prep = futs.pop(i_t).result()
                if prep["rng_state"] is not None:
                    last_rng_state[0] = prep["rng_state"]
                mark("prep %d" % i_t)
                if cell_sharded:
                    # shared draws travel owner -> pinned -> device -> NCCL broadcast and are consumed on the
                    # device; the other ranks computed everything else about the trait themselves
                    ctrl = shipper.ship(prep["ctrl_cols"], i_t % g_world, n_ctrl, prep["ctrl_gw"].shape[0]) \
                        if share_draws else prep["ctrl_cols"]
                    ticket = scorer.submit(prep["trait_cols"], prep["trait_gw"], ctrl, prep["ctrl_gw"], weight_opt,
                                           use_ratio, want_ctrl_raw=return_ctrl_raw_score,
                                           want_ctrl_norm=return_ctrl_norm_score)
                else:
                    # the next trait is queued behind the current one before the current one's results are
                    # fetched: the GPU never waits for the host between traits
                    if text_output and (i_t - 2) in text_writes:
                        text_writes.pop(i_t - 2).result()      # this submit overwrites that trait's pinned text
                    ticket = ctx.score_trait_submit(prep["trait_cols"], prep["trait_gw"], prep["ctrl_cols"],
                                                    prep["ctrl_gw"], weight_opt, use_ratio,
                                                    want_ctrl_raw=return_ctrl_raw_score,
                                                    want_ctrl_norm=return_ctrl_norm_score)
                mark("submitted %d" % i_t)
                if in_flight is not None:
                    collect(in_flight, i_t - 1)
                    mark("collected %d" % (i_t - 1))
                in_flight = (t, ticket)
            if in_flight is not None:
                last, in_flight = in_flight, None
                collect(last, len(traits) - 1)
            drain(0)
            for w_ in written:
                w_.result()

        try:
            run_all()
        finally:
            if in_flight is not None:       # an error above: free the slot, drop its results
                try:
                    ctx.score_trait_collect(in_flight[1])
                except Exception:
                    pass
            for w_ in written:              # nobody may still read pinned text when the context moves on
                try:
                    w_.result()
                except Exception:
                    pass
            ctx.set_text_output(False)
    mark("done")
    if trace is not None:
        import sys

        sys.stderr.write("score_traits trace: " + " | ".join("%s %.3f" % x for x in trace) + "\n")
    # numpy's global generator: where the reference's loop over traits would leave it (after the last trait's draws)
    np.random.seed(random_seed)
    if last_rng_state[0] is not None:
        np.random.set_state(last_rng_state[0])
    return results if on_result is None else traits


# ----------------------------------------------------------------------------------------------
# downstream analyses (scdrs/method.py:712-1166): implemented in scdrs_b200/downstream.py
# ----------------------------------------------------------------------------------------------
from .downstream import (  # noqa: E402,F401
    _pearson_corr,
    _pearson_corr_sparse,
    downstream_corr_analysis,
    downstream_gene_analysis,
    downstream_group_analysis,
    gearys_c,
    test_gearysc,
)
