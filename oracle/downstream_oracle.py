"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's downstream analyses
(scdrs/method.py:712-1166: downstream_group_analysis, downstream_corr_analysis, downstream_gene_analysis,
test_gearysc, gearys_c, _pearson_corr_sparse), used to check ``scdrs_b200/downstream.py`` (device path).
Only ``tests/`` may import this module.

Pinned: against the reference's own golden outputs for the toy data set
(tests/golden/toy/*.scdrs_cell_corr, *.scdrs_gene, *.scdrs_group.cell_type -- all columns except hetero_*,
which need scanpy's neighbour graph, absent here) and against the UNMODIFIED reference functions executed
through oracle/ref_shim.py on seeded inputs with a synthetic neighbour graph (tests/test_oracle.py,
container only; their outputs are committed as tests/golden/ref_downstream_case.npz by
tests/golden/make_golden_downstream.py).  ``multipletests(method="fdr_bh")`` (statsmodels, absent) is restated
as the Benjamini-Hochberg step-up.
"""
import numpy as np
import pandas as pd
from scipy import sparse
from scipy.stats import rankdata


def fdr_bh(p):
    p = np.asarray(p, dtype=np.float64)
    n = p.shape[0]
    o = np.argsort(p)
    adj = np.minimum.accumulate((p[o] * n / np.arange(1, n + 1))[::-1])[::-1]
    out = np.empty(n)
    out[o] = np.minimum(adj, 1.0)
    return out


def gearys_c(graph, vals):
    """scdrs/method.py:1044-1085, the per-cell loop written over the COO entries."""
    G = sparse.csr_matrix(graph)
    data = G.data.astype(np.float64)
    vals = np.asarray(vals)
    W = data.sum()
    N = G.shape[0]
    vals_bar = vals.mean()
    vals = vals.astype(np.float64)
    rows = np.repeat(np.arange(N), np.diff(G.indptr))
    total = float(np.sum(data * (vals[rows] - vals[G.indices]) ** 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64((N - 1) * total) / np.float64(2 * W * ((vals - vals_bar) ** 2).sum())


def distribution_match(v, ref):
    """scdrs/method.py:976-981"""
    return np.sort(ref)[rankdata(v, method="ordinal").astype(np.int64) - 1]


def test_gearysc(graph, labels, norm_score, ctrl_norm, opt="control_distribution_match"):
    """scdrs/method.py:915-1041 on arrays: labels[n], norm_score[n], ctrl_norm[n x n_null].
    Returns a frame indexed by group in order of first appearance."""
    G = sparse.csr_matrix(graph)
    labels = np.asarray(labels)
    index = list(pd.unique(labels))
    n_null = ctrl_norm.shape[1]
    stats = pd.DataFrame(index=index, columns=["trait"] + ["null_%d" % i for i in range(n_null)], data=np.nan)
    for g in sorted(set(labels)):
        idx = np.nonzero(labels == g)[0]
        Gg = G[idx][:, idx]
        s = np.asarray(norm_score)[idx]
        stats.loc[g, "trait"] = gearys_c(Gg, s)
        for i in range(n_null):
            v = np.asarray(ctrl_norm)[idx, i]
            if opt == "control_distribution_match":
                v = distribution_match(v, s)
            elif opt == "permutation":
                v = np.random.permutation(s)
            stats.loc[g, "null_%d" % i] = gearys_c(Gg, v)
    null_cols = [c for c in stats.columns if c.startswith("null_")]
    pval = ((stats["trait"].values > stats[null_cols].values.T).sum(axis=0) + 1) / (n_null + 1)
    pval[np.isnan(stats["trait"])] = np.nan
    out = pd.DataFrame({"pval": pval, "trait": stats["trait"].values, "ctrl_mean": stats[null_cols].mean(axis=1).values,
                        "ctrl_std": stats[null_cols].std(axis=1).values}, index=stats.index)
    out["zsc"] = -(out["trait"].values - out["ctrl_mean"]) / out["ctrl_std"]
    return out


test_gearysc.__test__ = False


def group_analysis(graph, labels, norm_score, ctrl_norm, pval, fdr_thresholds=(0.05, 0.1, 0.2)):
    """scdrs/method.py:712-819 for one annotation, cells already aligned."""
    labels = np.asarray(labels)
    groups = sorted(set(labels))
    cols = ["n_cell", "n_ctrl", "assoc_mcp", "assoc_mcz", "hetero_mcp", "hetero_mcz"] + ["n_fdr_%s" % t for t in fdr_thresholds]
    res = pd.DataFrame(index=groups, columns=cols, dtype=np.float32)
    fdr = fdr_bh(pval)
    n_ctrl = ctrl_norm.shape[1]
    for g in groups:
        m = labels == g
        res.loc[g, ["n_cell", "n_ctrl"]] = [m.sum(), n_ctrl]
        for t in fdr_thresholds:
            res.loc[g, "n_fdr_%s" % t] = (fdr[m] < t).sum()
        q = np.quantile(np.asarray(norm_score)[m], 0.95)
        vq = np.quantile(np.asarray(ctrl_norm)[m], 0.95, axis=0)
        res.loc[g, ["assoc_mcp", "assoc_mcz"]] = [((vq >= q).sum() + 1) / (vq.shape[0] + 1), (q - vq.mean()) / vq.std()]
    if graph is not None:
        rls = test_gearysc(graph, labels, norm_score, ctrl_norm)
        for g in groups:
            res.loc[g, ["hetero_mcp", "hetero_mcz"]] = [rls.loc[g, "pval"], rls.loc[g, "zsc"]]
    return res


def corr_analysis(variables, norm_score, ctrl_norm):
    """scdrs/method.py:822-865: rows (n_ctrl, corr_mcp, corr_mcz) per variable."""
    out = []
    for v in variables:
        c = np.corrcoef(v, norm_score)[0, 1]
        vc = np.array([np.corrcoef(v, ctrl_norm[:, i])[0, 1] for i in range(ctrl_norm.shape[1])])
        out.append([ctrl_norm.shape[1], ((vc >= c).sum() + 1) / (vc.shape[0] + 1), (c - vc.mean()) / vc.std()])
    return np.array(out)


def pearson_corr_sparse(mat_X, mat_Y):
    """scdrs/method.py:1127-1166 with _get_mean_var (scdrs/pp.py:467-517) written out."""
    X = sparse.csr_matrix(mat_X)
    Y = np.asarray(mat_Y, dtype=np.float64).reshape(X.shape[0], -1)
    n = X.shape[0]
    mx = np.asarray(X.mean(axis=0)).reshape(-1).astype(np.float64)
    vx = np.asarray(X.power(2).mean(axis=0)).reshape(-1).astype(np.float64) - mx ** 2
    my, vy = Y.mean(axis=0), (Y ** 2).mean(axis=0) - Y.mean(axis=0) ** 2
    c = np.asarray(X.T.dot(Y)) / n - mx[:, None] * my[None, :]
    c = c / (np.sqrt(vx).clip(min=1e-8)[:, None] * np.sqrt(vy).clip(min=1e-8)[None, :])
    return np.array(c, dtype=np.float32)


def gene_analysis(X, var_names, norm_score):
    """scdrs/method.py:868-909"""
    res = pd.DataFrame(index=var_names, columns=["CORR", "RANK"], dtype=np.float32)
    res["CORR"] = pearson_corr_sparse(X, norm_score).reshape(-1)
    res.sort_values("CORR", ascending=False, inplace=True)
    res["RANK"] = np.arange(res.shape[0])
    return res
