"""Host side of preprocessing: same functions and ``SCDRS_PARAM`` schema as ``scdrs/pp.py`` of the
reference.  The per-gene / per-cell reductions of ``compute_stats`` run on the GPU
(``csrc/stats.cu``); covariate regression (a k x k solve), the LOESS mean-variance trend and the
``pd.qcut`` binning stay on the host, as BASELINE.json's north_star prescribes.
"""
import warnings
from typing import List

import numpy as np
import pandas as pd
from scipy import sparse

from ._device import get_state
from .loess1d import loess


def category2dummy(df: pd.DataFrame, cols: List[str] = None, verbose: bool = False) -> pd.DataFrame:
    """Categorical columns -> ``drop_first`` dummy columns (scdrs/pp.py:10-60)."""
    df = df.copy()
    if cols is None:
        cols = [c for c in df.columns if c not in set(df._get_numeric_data().columns)]
    assert set(cols).issubset(set(df.columns)), "'cols' must be a subset of df.columns"
    added, pieces = [], []
    for col in cols:
        dummy = pd.get_dummies(df[col], drop_first=True)
        dummy.columns = [f"{col}_{s}" for s in dummy.columns]
        dummy = dummy.astype(float)
        dummy.loc[df[col].isnull(), dummy.columns] = np.nan
        added.extend(dummy.columns)
        pieces.append(dummy)
    df = pd.concat([df, *pieces], axis=1).drop(columns=list(cols))
    if len(added) > 0 and verbose:
        print(
            "scdrs.pp.category2dummy: "
            f"Detected categorical columns: {','.join(cols)}. "
            f"Added dummy columns: {','.join(added)}. "
            f"Dropped columns: {','.join(cols)}."
        )
    return df


def reg_out(mat_Y, mat_X):
    """Residual of mat_Y after OLS on mat_X (scdrs/pp.py:425-464)."""
    mat_X = mat_X.toarray() if sparse.issparse(mat_X) else np.array(mat_X)
    if mat_X.ndim == 1:
        mat_X = mat_X.reshape([-1, 1])
    mat_Y = mat_Y.toarray() if sparse.issparse(mat_Y) else np.array(mat_Y)
    if mat_Y.ndim == 1:
        mat_Y = mat_Y.reshape([-1, 1])
    n_sample = mat_Y.shape[0]
    coef = np.linalg.solve(mat_X.T @ mat_X / n_sample, mat_X.T @ mat_Y / n_sample)
    resid = mat_Y - mat_X @ coef
    if resid.shape[1] == 1:
        resid = resid.reshape([-1])
    return resid


def preprocess(data, cov=None, adj_prop=None, n_mean_bin=20, n_var_bin=20, n_chunk=None, copy=False):
    """Covariate correction + gene/cell statistics (drop-in for scdrs/pp.py:63-273).

    Writes ``data.uns["SCDRS_PARAM"]`` = {FLAG_SPARSE, FLAG_COV, FLAG_ADJ_PROP[, COV_MAT, COV_BETA,
    COV_GENE_MEAN], GENE_STATS, CELL_STATS}.  ``n_chunk`` is accepted for compatibility; the device
    reductions never densify X, so it has no effect.
    """
    adata = data.copy() if copy else data
    n_cell, n_gene = adata.shape

    flag_sparse = sparse.issparse(adata.X)
    flag_cov = cov is not None
    flag_adj_prop = adj_prop is not None
    adata.uns["SCDRS_PARAM"] = {
        "FLAG_SPARSE": flag_sparse,
        "FLAG_COV": flag_cov,
        "FLAG_ADJ_PROP": flag_adj_prop,
    }
    if flag_sparse:
        if not isinstance(adata.X, sparse.csr_matrix):
            adata.X = sparse.csr_matrix(adata.X)                       # :181-182
    elif not isinstance(adata.X, np.ndarray):
        adata.X = np.array(adata.X)                                     # :185-186

    if flag_cov:
        assert (
            len(pd.Index(cov.index).intersection(pd.Index(adata.obs_names))) > 0.75 * n_cell
        ), "cov does not match the cells in data"
        df_cov = pd.DataFrame(index=adata.obs_names)
        df_cov = df_cov.join(cov)
        df_cov = category2dummy(df=df_cov, verbose=True)
        df_cov = df_cov.fillna(df_cov.mean())
        v_resid = reg_out(np.ones(n_cell), df_cov.values.astype(np.float64))
        if (v_resid ** 2).mean() > 0.01:                                # :201-203
            df_cov["SCDRS_CONST"] = 1
        cov_values = df_cov.values.astype(np.float64)
        if flag_sparse and adata.X.dtype == np.float32 and cov_values.shape[1] <= 16:
            # gene means and C^T X in one device pass over X, in scipy's own operation order
            ctx = get_state(adata).ensure_matrix(adata)
            v_gene_mean, xtc = ctx.gene_mean_xtc(cov_values)
            mat_xtc = xtc.T
        else:
            v_gene_mean = np.array(adata.X.mean(axis=0)).flatten()      # :206
            mat_xtc = np.asarray((adata.X.T @ cov_values).T) if flag_sparse else None
        if flag_sparse:
            mat_beta = np.linalg.solve(                                 # :209-212
                cov_values.T @ cov_values / n_cell,
                mat_xtc / n_cell,
            )
            adata.uns["SCDRS_PARAM"]["COV_MAT"] = df_cov
            adata.uns["SCDRS_PARAM"]["COV_BETA"] = pd.DataFrame(
                -mat_beta.T, index=adata.var_names, columns=df_cov.columns
            )
            adata.uns["SCDRS_PARAM"]["COV_GENE_MEAN"] = pd.Series(v_gene_mean, index=adata.var_names)
        else:
            adata.X = reg_out(adata.X, cov_values)                      # :223-224
            adata.X += v_gene_mean

    implicit_cov_corr = bool(flag_sparse and flag_cov)

    if flag_adj_prop:
        err_msg = "'adj_prop'=%s not in 'adata.obs.columns'" % adj_prop
        assert adj_prop in adata.obs, err_msg
        err_msg = "On average <10 cells per group, maybe `%s` is not categorical?" % adj_prop
        assert adata.obs[adj_prop].unique().shape[0] < 0.1 * n_cell, err_msg
        size = adata.obs[adj_prop].value_counts()
        # smallest group size is at least 1% of the largest (docstring scdrs/pp.py:93-94; the
        # reference's in-place clip at :253 is a silent no-op under pandas >= 3)
        size = size.clip(lower=int(0.01 * size.max()))
        cell_weight = (n_cell / size.reindex(adata.obs[adj_prop].values).values).astype(np.float64)
        cell_weight = cell_weight / cell_weight.mean()
    else:
        cell_weight = None

    df_gene, df_cell = compute_stats(
        adata,
        implicit_cov_corr=implicit_cov_corr,
        cell_weight=cell_weight,
        n_mean_bin=n_mean_bin,
        n_var_bin=n_var_bin,
        n_chunk=n_chunk,
    )
    adata.uns["SCDRS_PARAM"]["GENE_STATS"] = df_gene
    adata.uns["SCDRS_PARAM"]["CELL_STATS"] = df_cell
    return adata if copy else None


def _device_mean_var(adata, implicit_cov_corr, axis, weights=None, transform=0):
    """mean / variance along ``axis`` of T(corrected X) from the device sums."""
    state = get_state(adata)
    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, implicit_cov_corr)
    n_cell, n_gene = adata.shape
    if axis == 0:
        # np.var (centred, two passes) for unweighted dense blocks -- implicit mode and dense X
        # (scdrs/pp.py:498-501); E[x^2] - E[x]^2 otherwise (:494-496, :508-515)
        center = weights is None and (implicit_cov_corr or not sparse.issparse(adata.X))
        denom = float(n_cell) if weights is None else float(np.sum(weights))
        gs, gq, _, _ = ctx.gene_cell_stats(transform, cell_weight=weights, want_gene=True, want_cell=False,
                                           center=center, denom=denom)
        v_mean = gs / denom
        v_var = gq / denom if center else gq / denom - v_mean ** 2
    else:
        assert weights is None or len(weights) == n_gene
        if weights is not None:
            raise NotImplementedError("gene-weighted per-cell statistics are not on the scDRS path")
        _, _, cs, cq = ctx.gene_cell_stats(transform, want_gene=False, want_cell=True)
        v_mean = cs / n_gene
        v_var = cq / n_gene - v_mean ** 2
    return v_mean, v_var


def compute_stats(adata, implicit_cov_corr=False, cell_weight=None, n_mean_bin=20, n_var_bin=20, n_chunk=20):
    """Gene- and cell-level statistics (drop-in for scdrs/pp.py:276-419).

    Returns (df_gene [mean, var, var_tech, ct_mean, ct_var, ct_var_tech, mean_var], df_cell [mean, var]).
    """
    if implicit_cov_corr:
        assert (
            "SCDRS_PARAM" in adata.uns
        ), "adata.uns['SCDRS_PARAM'] is not found, run `scdrs.pp.preprocess` before calling this function"
    df_gene = pd.DataFrame(
        index=adata.var_names,
        columns=["mean", "var", "var_tech", "ct_mean", "ct_var", "ct_var_tech", "mean_var"],
    )
    df_cell = pd.DataFrame(index=adata.obs_names, columns=["mean", "var"])
    w = None if cell_weight is None else np.asarray(cell_weight, dtype=np.float64)

    df_gene["mean"], df_gene["var"] = _device_mean_var(adata, implicit_cov_corr, 0, w, transform=0)       # :353-355/:368-370
    df_gene["ct_mean"], df_gene["ct_var"] = _device_mean_var(adata, implicit_cov_corr, 0, w, transform=1)  # :359-364/:371-373

    # technical variance from the LOESS mean-variance trend (:376-390)
    ct_mean = df_gene["ct_mean"].values.astype(np.float64)
    ct_var = df_gene["ct_var"].values.astype(np.float64)
    not_const = ct_var > 0
    if (ct_mean <= 0).sum() > 0:
        warnings.warn("%d genes with ct_mean<0" % (ct_mean < 0).sum())
        not_const = not_const & (ct_mean > 0)
    estimat_var = np.zeros(adata.shape[1], dtype=np.float64)
    model = loess(np.log10(ct_mean[not_const]), np.log10(ct_var[not_const]), span=0.3, degree=2)
    model.fit()
    estimat_var[not_const] = model.outputs.fitted_values
    df_gene["ct_var_tech"] = 10 ** estimat_var
    with np.errstate(divide="ignore", invalid="ignore"):
        var_tech = df_gene["var"].values * df_gene["ct_var_tech"].values / ct_var
    var_tech[np.isnan(var_tech)] = 0
    df_gene["var_tech"] = var_tech

    # n_mean_bin * n_var_bin mean_var bins (:392-407)
    n_bin_max = np.floor(np.sqrt(adata.shape[1] / 10)).astype(int)
    if (n_mean_bin > n_bin_max) | (n_var_bin > n_bin_max):
        n_mean_bin, n_var_bin = n_bin_max, n_bin_max
        print("Too few genes for 20*20 bins, setting n_mean_bin=n_var_bin=%d" % (n_bin_max))
    v_mean_bin = pd.qcut(df_gene["mean"], n_mean_bin, labels=False, duplicates="drop")
    mean_var = np.full(adata.shape[1], "", dtype=object)
    for bin_ in sorted(set(v_mean_bin)):
        ind_select = (v_mean_bin == bin_).values
        v_var_bin = pd.qcut(df_gene.loc[ind_select, "var"], n_var_bin, labels=False, duplicates="drop")
        mean_var[ind_select] = ["%d.%d" % (bin_, x) for x in v_var_bin]
    df_gene["mean_var"] = mean_var

    df_cell["mean"], df_cell["var"] = _device_mean_var(adata, implicit_cov_corr, 1)                        # :409-417
    return df_gene, df_cell


def _get_mean_var(sparse_X, axis=0, weights=None):
    """Mean and variance of a (sparse or dense) matrix (drop-in for scdrs/pp.py:467-517), computed
    on the device from one pass over its non-zeros."""
    from .anndata_lite import AnnData

    X = sparse_X if sparse.issparse(sparse_X) else np.asarray(sparse_X)
    tmp = AnnData(X)
    tmp.uns["SCDRS_PARAM"] = {"FLAG_SPARSE": True, "FLAG_COV": False}
    if axis == 0:
        return _device_mean_var(tmp, False, 0, None if weights is None else np.asarray(weights, dtype=np.float64))
    if weights is None:
        return _device_mean_var(tmp, False, 1)
    # weights over genes: transpose so that they become cell weights
    tmpT = AnnData(sparse.csr_matrix(X.T) if sparse.issparse(X) else np.ascontiguousarray(X.T))
    tmpT.uns["SCDRS_PARAM"] = {"FLAG_SPARSE": True, "FLAG_COV": False}
    return _device_mean_var(tmpT, False, 0, np.asarray(weights, dtype=np.float64))


def _get_mean_var_implicit_cov_corr(adata, axis=0, weights=None, transform_func=None, n_chunk=20):
    """Mean and variance of ``transform(adata.X + COV_MAT * COV_BETA + COV_GENE_MEAN)`` (drop-in for
    scdrs/pp.py:550-641).  ``transform_func`` must be None or ``np.expm1`` (the two the reference uses)."""
    assert axis in [0, 1], "axis must be one of [0, 1]"
    assert (
        "SCDRS_PARAM" in adata.uns
    ), "adata.uns['SCDRS_PARAM'] not found, run `preprocess` before calling this function"
    if transform_func is None:
        transform = 0
    elif transform_func is np.expm1:
        transform = 1
    else:
        raise NotImplementedError("only identity and np.expm1 transforms run on the device")
    w = None if weights is None else np.asarray(weights, dtype=np.float64)
    if axis == 1 and w is not None:
        raise NotImplementedError("gene-weighted per-cell statistics are not on the scDRS path")
    return _device_mean_var(adata, True, axis, w, transform=transform)
