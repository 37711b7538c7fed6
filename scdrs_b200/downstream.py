"""The analyses run on a trait's ``.full_score`` right after scoring: same
functions, arguments, assertions and return frames as ``scdrs/method.py:712-1166`` of the reference
(``downstream_group_analysis``, ``downstream_corr_analysis``, ``downstream_gene_analysis``, ``test_gearysc``,
``gearys_c``, ``_pearson_corr``, ``_pearson_corr_sparse``), re-exported from ``scdrs_b200.method``.

On the device (``csrc/downstream.cu``, no CPU fallback): everything that touches the n_cell x (1 + n_ctrl)
score matrix, the neighbour graph or the expression matrix -- correlations, per-group sorts / quantiles /
distribution matching, Geary's C of every column of every group.  On the host, in numpy / pandas as in the
reference: aligning cells, building the group order and the within-group graph, FDR, and the per-group
Monte-Carlo p-values and z-scores computed from n_group x (1 + n_ctrl) numbers.
"""
import warnings

import numpy as np
import pandas as pd
from scipy import sparse

from . import _device
from .util import fdr_bh


def _ctx():
    from .method import _scratch_context

    return _scratch_context()


def _align(adata, df_full_score):
    """``cell_list = sorted(set(adata.obs_names) & set(df_full_score.index))`` (scdrs/method.py:757) as row
    numbers: (rows of df_full_score, rows of adata), both in cell_list order."""
    in_adata = adata.obs_names.get_indexer(df_full_score.index)
    if df_full_score.index.is_unique and adata.obs_names.is_unique:
        df_rows = np.nonzero(in_adata >= 0)[0]
        by_name = np.asarray(df_full_score.index[df_rows].argsort())       # unique names: any sort is stable
        return df_rows[by_name], in_adata[df_rows][by_name]
    cell_list = sorted(set(adata.obs_names) & set(df_full_score.index))
    return df_full_score.index.get_indexer(cell_list), adata.obs_names.get_indexer(cell_list)


def _upload_scores(ctx, df, cols, rows=None):
    """Put df[cols] on the device.  When every row takes part the frame's memory is uploaded as it is (no
    host copy, the device picks the columns) and the rows keep the frame's order; otherwise the selected
    rows are gathered first.  Returns the device row of every selected cell."""
    n = df.shape[0]
    whole = rows is None or (len(rows) == n)
    same = len(set(df.dtypes)) == 1 and df.dtypes.iloc[0] in (np.float32, np.float64)
    if whole and same:
        ctx.ds_set_scores(df.to_numpy(copy=False), df.columns.get_indexer(cols))
        return np.arange(n, dtype=np.int64) if rows is None else np.asarray(rows, dtype=np.int64)
    sub = df[cols]
    f32 = all(dt == np.float32 for dt in sub.dtypes)
    vals = sub.to_numpy(dtype=np.float32 if f32 else np.float64)
    if whole:
        ctx.ds_set_scores(vals)
        return np.arange(n, dtype=np.int64) if rows is None else np.asarray(rows, dtype=np.int64)
    ctx.ds_set_scores(np.ascontiguousarray(vals[rows]))
    return np.arange(len(rows), dtype=np.int64)


def _quantile_positions(sizes, q):
    """numpy's index arithmetic for np.quantile(x, q) (method "linear") on a sample of each size: the lower
    order statistic and the interpolation weight (numpy/lib/_function_base_impl.py: virtual index
    (n - 1) * q of the "linear" method, _get_indexes, _get_gamma)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    quantiles = np.float64(q)
    lo = np.zeros(sizes.shape[0], np.int32)
    gamma = np.zeros(sizes.shape[0], np.float64)
    for i, n in enumerate(sizes):
        if n <= 0:
            continue
        virtual = (int(n) - 1) * quantiles
        prev = np.floor(virtual)
        if virtual >= n - 1:
            lo[i], gamma[i] = n - 1, 0.0
        elif virtual < 0:
            lo[i], gamma[i] = 0, 0.0
        else:
            lo[i], gamma[i] = int(prev), virtual - prev
    return lo, gamma


def _group_layout(codes, n_group):
    """cells sorted by group code (order inside a group unchanged) and the group boundaries."""
    codes = np.asarray(codes, dtype=np.int64)
    order = np.argsort(codes, kind="stable").astype(np.int64)
    grp_ptr = np.zeros(n_group + 1, np.int64)
    np.cumsum(np.bincount(codes, minlength=n_group), out=grp_ptr[1:])
    return order, grp_ptr


def _group_graph(graph, codes, order):
    """The neighbour graph with cells in group order, edges between groups dropped: what
    ``adata[group_index].obsp["connectivities"]`` holds for every group at once.  Returns the CSR parts
    (indptr int64, neighbour positions int32, weights float64); the entries of a row keep their stored order."""
    G = sparse.csr_matrix(graph)
    assert G.shape[0] == G.shape[1]
    n = G.shape[0]
    codes = np.asarray(codes, dtype=np.int64)
    Gp = G[order]                                            # rows in group order (a C-level row gather)
    row_p = np.repeat(np.arange(n, dtype=np.int64), np.diff(Gp.indptr))
    keep = codes[order][row_p] == codes[Gp.indices]
    pos = np.empty(n, np.int64)
    pos[order] = np.arange(n, dtype=np.int64)
    indptr = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(row_p[keep], minlength=n), out=indptr[1:])
    return indptr, pos[Gp.indices[keep]].astype(np.int32), Gp.data[keep].astype(np.float64)


def _geary_from_totals(total, denom, n_in_group, w_in_group):
    """C = (N - 1) total / (2 W denom)                                         scdrs/method.py:1080-1083"""
    with np.errstate(divide="ignore", invalid="ignore"):
        numer = (np.asarray(n_in_group, dtype=np.float64)[:, None] - 1) * total
        den = 2 * np.asarray(w_in_group, dtype=np.float64)[:, None] * denom
        return numer / den


def _group_weight_sums(graph_parts, grp_ptr):
    """W of every group: graph.data.sum() of the group's own block                scdrs/method.py:1063"""
    indptr, _, w = graph_parts
    return np.array([w[indptr[grp_ptr[g]]:indptr[grp_ptr[g + 1]]].sum() for g in range(len(grp_ptr) - 1)])


# ----------------------------------------------------------------------------------------------
# Geary's C                                                              scdrs/method.py:1044-1085
# ----------------------------------------------------------------------------------------------
def gearys_c(adata, vals):
    """Geary's C of ``vals`` on ``adata.obsp["connectivities"]`` (drop-in for scdrs/method.py:1044-1085)."""
    graph = adata.obsp["connectivities"]
    assert graph.shape[0] == graph.shape[1]
    vals = np.asarray(vals)
    assert graph.shape[0] == vals.shape[0]
    assert np.ndim(vals) == 1
    G = sparse.csr_matrix(graph)
    G.sort_indices()
    W = G.data.astype(np.float64, copy=False).sum()
    N = G.shape[0]
    ctx = _ctx()
    ctx.ds_set_scores(vals.astype(np.float64).reshape(-1, 1))
    _, total, denom = ctx.ds_group_stats(
        np.arange(N, dtype=np.int64), np.array([0, N], np.int64), geary_mode=2,
        graph=(G.indptr.astype(np.int64), G.indices.astype(np.int32), G.data.astype(np.float64)))
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((N - 1) * total[0, 0]) / (2 * W * denom[0, 0])


def _gearys_by_group(ctx, codes, n_group, graph, geary_mode, q=None, dev_rows=None):
    """Device pass for one grouping of the uploaded score matrix: (quantile or None, C[n_group x ncol]).
    codes / graph follow the caller's cell order; dev_rows[i] = row of cell i in the uploaded matrix."""
    order, grp_ptr = _group_layout(codes, n_group)
    sizes = np.diff(grp_ptr)
    q_lo = q_gamma = None
    if q is not None:
        q_lo, q_gamma = _quantile_positions(sizes, q)
    parts = None
    if geary_mode:
        parts = _group_graph(graph, codes, order)
    quant, total, denom = ctx.ds_group_stats(order if dev_rows is None else dev_rows[order], grp_ptr, q_lo, q_gamma,
                                             geary_mode, parts)
    C = None
    if geary_mode:
        w_sum = _group_weight_sums(parts, grp_ptr)
        C = _geary_from_totals(total, denom, sizes, w_sum)
    return quant, C


def test_gearysc(adata, df_full_score, groupby, opt="control_distribution_match"):
    """Significance of Geary's C per cell group (drop-in for scdrs/method.py:915-1041)."""
    assert np.all(df_full_score.index == adata.obs_names), "adata.obs_names must match df_full_score.index"
    ctrl_cols = [col for col in df_full_score.columns if col.startswith("ctrl_norm_score_")]
    n_null = len(ctrl_cols)
    df_meta = adata.obs
    labels = df_meta[groupby]
    index_groups = list(labels.unique())
    codes, groups = pd.factorize(np.asarray(labels.values, dtype=object), sort=True)   # groupby order: sorted, non-empty
    groups = list(groups)
    code_of = {g: i for i, g in enumerate(groups)}

    ctx = _ctx()
    if opt == "control_distribution_match":
        mode = 1
    elif opt == "control":
        mode = 2
    elif opt == "permutation":
        mode = 2
    else:
        raise NotImplementedError
    if opt == "permutation":
        # the reference permutes the trait scores of a group once per control, groups in sorted order
        mat = np.empty((df_full_score.shape[0], 1 + n_null), np.float64)
        mat[:, 0] = df_full_score["norm_score"].values
        for g in groups:
            rows = np.nonzero(codes == code_of[g])[0]
            for i_null in range(n_null):
                mat[rows, 1 + i_null] = np.random.permutation(mat[rows, 0])
        ctx.ds_set_scores(mat)
    else:
        _upload_scores(ctx, df_full_score, ["norm_score"] + ctrl_cols)
    _, C = _gearys_by_group(ctx, codes, len(groups), adata.obsp["connectivities"], mode)

    # per group: Geary's C of the trait against its controls                     scdrs/method.py:1022-1040
    trait, null = C[:, 0], C[:, 1:]
    pval = ((trait[:, None] > null).sum(axis=1) + 1) / (n_null + 1)
    pval[np.isnan(trait)] = np.nan
    with np.errstate(divide="ignore", invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)                # groups without edges: all NaN
        ctrl_mean = np.nanmean(null, axis=1)
        ctrl_std = np.nanstd(null, axis=1, ddof=1) if n_null > 1 else np.full(len(groups), np.nan)
        zsc = -(trait - ctrl_mean) / ctrl_std
    df_rls = pd.DataFrame({"pval": pval, "trait": trait, "ctrl_mean": ctrl_mean, "ctrl_std": ctrl_std, "zsc": zsc},
                          index=groups)
    return df_rls.reindex(index_groups)             # rows in order of first appearance, like the reference


test_gearysc.__test__ = False     # a library function named like the reference's, not a pytest test


# ----------------------------------------------------------------------------------------------
# group analysis                                                         scdrs/method.py:712-819
# ----------------------------------------------------------------------------------------------
def downstream_group_analysis(adata, df_full_score, group_cols, fdr_thresholds=[0.05, 0.1, 0.2]):
    """Group-level statistics keyed by annotation (drop-in for scdrs/method.py:712-819)."""
    assert "connectivities" in adata.obsp, "Expect `connectivities` in `adata.obsp`; run `sc.pp.neighbors` first"
    assert len(set(group_cols) - set(adata.obs)) == 0, "Missing `group_cols` variables from `adata.obs.columns`."

    df_rows, rows = _align(adata, df_full_score)            # cell_list order
    control_list = [x for x in df_full_score.columns if x.startswith("ctrl_norm_score")]
    n_ctrl = len(control_list)
    df_reg = adata.obs.iloc[rows][list(group_cols)]
    pval = df_full_score["pval"].values[df_rows]
    graph = sparse.csr_matrix(adata.obsp["connectivities"])[rows][:, rows]

    ctx = _ctx()
    dev_rows = _upload_scores(ctx, df_full_score, ["norm_score"] + control_list, df_rows)
    v_fdr = fdr_bh(pval)

    dict_df_res = {}
    for group_col in group_cols:
        group_list = sorted(pd.unique(adata.obs[group_col]))              # = sorted(set(...)), :771
        res_cols = ["n_cell", "n_ctrl", "assoc_mcp", "assoc_mcz", "hetero_mcp", "hetero_mcz"]
        for fdr_threshold in fdr_thresholds:
            res_cols.append(f"n_fdr_{fdr_threshold}")
        out = np.full((len(group_list), len(res_cols)), np.nan, np.float32)

        codes, present = pd.factorize(np.asarray(df_reg[group_col].values, dtype=object), sort=True)
        present = list(present)
        q95, C = _gearys_by_group(ctx, codes, len(present), graph, 1, q=0.95, dev_rows=dev_rows)

        where = np.array([group_list.index(g) for g in present], dtype=np.int64)

        def put(col, values):
            out[where, res_cols.index(col)] = values

        put("n_cell", np.bincount(codes, minlength=len(present)))
        put("n_ctrl", n_ctrl)
        for fdr_threshold in fdr_thresholds:
            put(f"n_fdr_{fdr_threshold}", np.bincount(codes, weights=(v_fdr < fdr_threshold), minlength=len(present)))
        # association: the group's 0.95 quantile against the controls'              :797-806
        score_q95, v_ctrl_score_q95 = q95[:, 0], q95[:, 1:]
        put("assoc_mcp", ((v_ctrl_score_q95 >= score_q95[:, None]).sum(axis=1) + 1) / (n_ctrl + 1))
        put("assoc_mcz", (score_q95 - v_ctrl_score_q95.mean(axis=1)) / v_ctrl_score_q95.std(axis=1))
        # heterogeneity: Geary's C against the distribution-matched controls        :809-814, 1022-1040
        trait, null = C[:, 0], C[:, 1:]
        h_p = ((trait[:, None] > null).sum(axis=1) + 1) / (n_ctrl + 1)
        h_p[np.isnan(trait)] = np.nan
        with np.errstate(divide="ignore", invalid="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)            # groups without edges: all NaN
            h_z = -(trait - np.nanmean(null, axis=1)) / np.nanstd(null, axis=1, ddof=1)
        put("hetero_mcp", h_p)
        put("hetero_mcz", h_z)
        df_res = pd.DataFrame(out, index=group_list, columns=res_cols)
        df_res.index.name = "group"
        dict_df_res[group_col] = df_res
    return dict_df_res


# ----------------------------------------------------------------------------------------------
# cell-level variable correlation                                         scdrs/method.py:822-865
# ----------------------------------------------------------------------------------------------
def downstream_corr_analysis(adata, df_full_score, var_cols):
    """Monte-Carlo test of the correlation between cell-level variables and the trait score (drop-in for
    scdrs/method.py:822-865)."""
    assert len(set(var_cols) - set(adata.obs)) == 0, "Missing `var_cols` variables from `adata.obs.columns`."
    df_rows, rows = _align(adata, df_full_score)            # cell_list order
    control_list = [x for x in df_full_score.columns if x.startswith("ctrl_norm_score")]
    n_ctrl = len(control_list)

    ctx = _ctx()
    dev_rows = _upload_scores(ctx, df_full_score, ["norm_score"] + control_list, df_rows)
    variables = np.empty((len(var_cols), len(rows)), np.float64)
    for i, v in enumerate(var_cols):                        # each variable in the order of the uploaded rows
        variables[i, dev_rows] = np.asarray(adata.obs[v].values[rows], dtype=np.float64)
    corr = ctx.ds_corr(variables)

    df_res = pd.DataFrame(index=var_cols, columns=["n_ctrl", "corr_mcp", "corr_mcz"], dtype=np.float32)
    for i, var_col in enumerate(var_cols):
        corr_, v_corr_ = corr[i, 0], corr[i, 1:]
        mc_p = ((v_corr_ >= corr_).sum() + 1) / (v_corr_.shape[0] + 1)
        mc_z = (corr_ - v_corr_.mean()) / v_corr_.std()
        df_res.loc[var_col] = [n_ctrl, mc_p, mc_z]
    return df_res


# ----------------------------------------------------------------------------------------------
# gene-level correlation                                          scdrs/method.py:868-909, 1088-1166
# ----------------------------------------------------------------------------------------------
def _corr_on_device(adata_like, mat_Y, centered_var):
    """Pearson correlation of every column of the resident matrix with every column of mat_Y [n x m]."""
    state = _device.get_state(adata_like)
    ctx = state.ensure_matrix(adata_like)
    state.ensure_cov(adata_like, False)
    n = adata_like.shape[0]
    gs, gq, _, _ = ctx.gene_cell_stats(0, want_gene=True, want_cell=False, center=centered_var, denom=float(n))
    v_X_mean = gs / n
    v_X_var = gq / n if centered_var else gq / n - v_X_mean ** 2
    v_X_sd = np.sqrt(v_X_var).clip(min=1e-8)
    v_Y_mean = mat_Y.mean(axis=0)
    v_Y_sd = (mat_Y.std(axis=0) if centered_var else np.sqrt((mat_Y ** 2).mean(axis=0) - v_Y_mean ** 2)).clip(min=1e-8)
    xty = np.concatenate(
        [ctx.gene_mean_xtc(np.ascontiguousarray(mat_Y[:, j:j + 16]))[1] for j in range(0, mat_Y.shape[1], 16)], axis=1)
    mat_corr = xty / n
    mat_corr = mat_corr - v_X_mean.reshape([-1, 1]).dot(v_Y_mean.reshape([1, -1]))
    mat_corr = mat_corr / v_X_sd.reshape([-1, 1]).dot(v_Y_sd.reshape([1, -1]))
    return np.array(mat_corr, dtype=np.float32)


def _pearson_corr(mat_X, mat_Y):
    """Pearson's correlation between every column of mat_X and of mat_Y (drop-in for
    scdrs/method.py:1088-1124; sparse or dense input, the products run on the device)."""
    from .anndata_lite import AnnData

    sparse_in = sparse.issparse(mat_X) or sparse.issparse(mat_Y)
    if not sparse.issparse(mat_X):
        mat_X = np.asarray(mat_X)
        if mat_X.ndim == 1:
            mat_X = mat_X.reshape([-1, 1])
    if sparse.issparse(mat_Y):
        mat_Y = mat_Y.toarray()
    mat_Y = np.asarray(mat_Y, dtype=np.float64)
    if mat_Y.ndim == 1:
        mat_Y = mat_Y.reshape([-1, 1])
    tmp = AnnData(mat_X if sparse.issparse(mat_X) else np.asarray(mat_X, dtype=np.float32))
    try:
        mat_corr = _corr_on_device(tmp, mat_Y, centered_var=not sparse_in)
    finally:
        _device._drop(id(tmp))
    if (mat_X.shape[1] == 1) | (mat_Y.shape[1] == 1):
        return mat_corr.reshape([-1])
    return mat_corr


def _pearson_corr_sparse(mat_X, mat_Y):
    """scdrs/method.py:1127-1166: the same through the sparse formulas (E[xy] - E[x]E[y]) / (sd sd)."""
    if not sparse.issparse(mat_X):
        mat_X = np.asarray(mat_X)
        mat_X = sparse.csr_matrix(mat_X.reshape([-1, 1]) if mat_X.ndim == 1 else mat_X)
    return _pearson_corr(mat_X, mat_Y)


def downstream_gene_analysis(adata, df_full_score):
    """Correlation between every gene and the trait score (drop-in for scdrs/method.py:868-909)."""
    cell_list = sorted(set(adata.obs_names) & set(df_full_score.index))
    y = df_full_score["norm_score"].loc[cell_list]
    if len(cell_list) == adata.shape[0]:
        # every cell takes part: use the matrix already resident for this AnnData, scores in its cell order
        v_y = np.asarray(y.reindex(adata.obs_names).values, dtype=np.float64).reshape(-1, 1)
        v_corr = _corr_on_device(adata, v_y, centered_var=not sparse.issparse(adata.X)).reshape([-1])
    else:
        v_corr = _pearson_corr(adata[adata.obs_names.get_indexer(cell_list)].X, y.values)
    df_res = pd.DataFrame(index=adata.var_names, columns=["CORR", "RANK"], dtype=np.float32)
    df_res["CORR"] = v_corr
    df_res.sort_values("CORR", ascending=False, inplace=True)
    df_res["RANK"] = np.arange(df_res.shape[0])
    return df_res
