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


LAST_TIMING = {}      # seconds per phase of the last preprocess / compute_stats call (bench.py reports them)
_FLAG_REL = 1e-6      # var below this fraction of the second moment: numerically constant, recomputed in numpy's order


def _tick(name, t0):
    import time

    LAST_TIMING[name] = LAST_TIMING.get(name, 0.0) + (time.perf_counter() - t0)
    return time.perf_counter()


def _exact_order():
    """SCDRS_B200_EXACT_ORDER=1: every gene through the sequential kernels that add cells in numpy's own order
    (round 1's path, 20 - 50 ms per pass at 110k cells); default: parallel moments, and only genes whose variance
    is rounding noise of that order are recomputed sequentially."""
    import os

    return os.environ.get("SCDRS_B200_EXACT_ORDER", "0") == "1"


class _Ranks:
    """The process group a sharded preprocess sums over (None: this process alone)."""

    def __init__(self, sharded, group=None):
        self.on = False
        self.group = group
        if sharded:
            import torch.distributed as dist

            self.on = dist.is_initialized() and dist.get_world_size(group) > 1

    def sum(self, *arrays):
        """Element-wise sums over the ranks of float64 arrays (scalars come back as 0-d arrays); identity when alone.
        Every rank receives the same bits (one all-reduce), so gene-level results are identical on all ranks."""
        arrays = [np.asarray(a, dtype=np.float64) for a in arrays]
        if self.on:
            import torch
            import torch.distributed as dist

            flat = torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrays]))
            if dist.get_backend(self.group) == "nccl":
                flat = flat.cuda()
            dist.all_reduce(flat, group=self.group)
            flat = flat.cpu().numpy()
            out, o = [], 0
            for a in arrays:
                out.append(flat[o:o + a.size].reshape(a.shape))
                o += a.size
            arrays = out
        return arrays[0] if len(arrays) == 1 else arrays

    def gather_objects(self, obj):
        if not self.on:
            return [obj]
        import torch.distributed as dist

        box = [None] * dist.get_world_size(self.group)
        dist.all_gather_object(box, obj, group=self.group)
        return box


def preprocess(data, cov=None, adj_prop=None, n_mean_bin=20, n_var_bin=20, n_chunk=None, copy=False, sharded=False,
               group=None):
    """Covariate correction + gene/cell statistics (drop-in for scdrs/pp.py:63-273).

    Writes ``data.uns["SCDRS_PARAM"]`` = {FLAG_SPARSE, FLAG_COV, FLAG_ADJ_PROP[, COV_MAT, COV_BETA,
    COV_GENE_MEAN], GENE_STATS, CELL_STATS}.  ``n_chunk`` is accepted for compatibility; the device
    reductions never densify X, so it has no effect.

    ``sharded=True`` (one process per GPU, ``torch.distributed`` initialised; sparse X): ``data`` holds this rank's
    block of cells of ONE atlas.  Everything gene-level -- gene means, C^T X and C^T C of the covariate regression
    (:206-212), the mean / variance moments of :353-373, the cell-group sizes of ``adj_prop`` -- is summed over the
    ranks of ``group``, so GENE_STATS, COV_BETA and COV_GENE_MEAN are those of the whole atlas and identical on
    every rank; COV_MAT and CELL_STATS cover this rank's cells.
    """
    import time

    LAST_TIMING.clear()
    t0 = time.perf_counter()
    adata = data.copy() if copy else data
    n_cell, n_gene = adata.shape
    ranks = _Ranks(sharded, group)

    flag_sparse = sparse.issparse(adata.X)
    flag_cov = cov is not None
    flag_adj_prop = adj_prop is not None
    assert flag_sparse or not ranks.on, "a sharded preprocess needs a sparse X (the dense mode residualises X in place)"
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
    n_total = float(ranks.sum(float(n_cell)))

    pre = None
    if flag_cov:
        assert (
            len(pd.Index(cov.index).intersection(pd.Index(adata.obs_names))) > 0.75 * n_cell
        ), "cov does not match the cells in data"
        df_cov = pd.DataFrame(index=adata.obs_names)
        df_cov = df_cov.join(cov)
        if ranks.on:
            # the dummy columns must not depend on which categories happen to live on this rank
            for col in [c for c in df_cov.columns if c not in set(df_cov._get_numeric_data().columns)]:
                cats = sorted(set().union(*ranks.gather_objects(set(df_cov[col].dropna().unique()))))
                df_cov[col] = pd.Categorical(df_cov[col], categories=cats)
        df_cov = category2dummy(df=df_cov, verbose=True)
        if ranks.on:
            col_sum, col_cnt = ranks.sum(df_cov.sum().values.astype(np.float64), df_cov.notna().sum().values.astype(np.float64))
            df_cov = df_cov.fillna(pd.Series(col_sum / col_cnt, index=df_cov.columns))
        else:
            df_cov = df_cov.fillna(df_cov.mean())
        # does the covariate matrix span the constant?  residual of 1 on the covariates (:199-203)
        c0 = df_cov.values.astype(np.float64)
        ctc0, ct1 = ranks.sum(c0.T @ c0, c0.sum(axis=0))
        resid = 1.0 - c0 @ np.linalg.solve(ctc0 / n_total, ct1 / n_total)
        if float(ranks.sum(float((resid ** 2).sum()))) / n_total > 0.01:   # :201-203
            df_cov["SCDRS_CONST"] = 1
        cov_values = df_cov.values.astype(np.float64)
        t0 = _tick("covariate_frame", t0)
        on_device = flag_sparse and adata.X.dtype == np.float32 and (cov_values.shape[1] <= 16 or not _exact_order())
        if on_device and _exact_order() and not ranks.on:
            # gene means and C^T X in one device pass over X, in scipy's own operation order
            ctx = get_state(adata).ensure_matrix(adata)
            t0 = _tick("upload_x", t0)
            v_gene_mean, xtc = ctx.gene_mean_xtc(cov_values)
            mat_xtc = xtc.T
        elif on_device:
            # sum x, sum x^2 and C^T X per gene in one parallel pass; the moments of all ranks add up
            ctx = get_state(adata).ensure_matrix(adata)
            t0 = _tick("upload_x", t0)
            pre = ctx.gene_moments(0, cov_values)
            pre = ranks.sum(pre)
            v_gene_mean = (pre[:, 0] / n_total).astype(np.float32)      # X.mean(axis=0) of float32 data is float32
            mat_xtc = pre[:, 2:].T
        else:
            v_gene_mean = np.array(adata.X.mean(axis=0)).flatten()      # :206
            mat_xtc = np.asarray((adata.X.T @ cov_values).T) if flag_sparse else None
        t0 = _tick("gene_mean_xtc", t0)
        if flag_sparse:
            ctc = ranks.sum(cov_values.T @ cov_values)
            mat_beta = np.linalg.solve(ctc / n_total, mat_xtc / n_total)  # :209-212
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
        size = adata.obs[adj_prop].value_counts()
        if ranks.on:
            total = {}
            for part in ranks.gather_objects(size.to_dict()):
                for key, val in part.items():
                    total[key] = total.get(key, 0) + int(val)
            size = pd.Series(total)
        assert size.shape[0] < 0.1 * n_total, err_msg
        # smallest group size is at least 1% of the largest (docstring scdrs/pp.py:93-94; the
        # reference's in-place clip at :253 is a silent no-op under pandas >= 3)
        size = size.clip(lower=int(0.01 * size.max()))
        cell_weight = (n_total / size.reindex(adata.obs[adj_prop].values).values).astype(np.float64)
        cell_weight = cell_weight / (float(ranks.sum(float(cell_weight.sum()))) / n_total)
    else:
        cell_weight = None
    t0 = _tick("covariate_frame", t0)

    df_gene, df_cell = compute_stats(
        adata,
        implicit_cov_corr=implicit_cov_corr,
        cell_weight=cell_weight,
        n_mean_bin=n_mean_bin,
        n_var_bin=n_var_bin,
        n_chunk=n_chunk,
        _ranks=ranks,
        _pre=pre if cell_weight is None else None,
        _keep_timing=True,
    )
    adata.uns["SCDRS_PARAM"]["GENE_STATS"] = df_gene
    adata.uns["SCDRS_PARAM"]["CELL_STATS"] = df_cell
    return adata if copy else None


def _device_mean_var(adata, implicit_cov_corr, axis, weights=None, transform=0):
    """mean / variance along ``axis`` of T(corrected X) from the device sums (sequential kernels: numpy's order)."""
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


def _cov_arrays(adata):
    """(C [n_cell x k], B [n_gene x k], mu [n_gene]) aligned to obs_names / var_names, as the device holds them."""
    param = adata.uns["SCDRS_PARAM"]
    cols = list(param["COV_MAT"])
    C = np.ascontiguousarray(param["COV_MAT"].reindex(index=adata.obs_names, columns=cols).values, dtype=np.float64)
    B = np.ascontiguousarray(param["COV_BETA"].reindex(index=adata.var_names, columns=cols).values, dtype=np.float64)
    mu = np.ascontiguousarray(param["COV_GENE_MEAN"].reindex(adata.var_names).values, dtype=np.float64)
    return C, B, mu


def _gene_mean_var_fast(adata, implicit_cov_corr, weights, ranks, pre=None):
    """Per-gene (mean, var, ct_mean, ct_var) of the (implicitly corrected) matrix from parallel device moments
    (csrc/stats_par.cu; scdrs/pp.py:353-373), summed over the ranks of a sharded atlas.  Genes whose variance is
    numerically zero against their second moment are recomputed by the sequential kernels in numpy's own order."""
    import time

    t0 = time.perf_counter()
    state = get_state(adata)
    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, implicit_cov_corr)
    t0 = _tick("upload_x", t0)
    n_cell, n_gene = adata.shape
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    wv = np.ones(n_cell) if w is None else w
    sparse_x = sparse.issparse(adata.X)
    if implicit_cov_corr:
        C, B, mu = _cov_arrays(adata)
        k = C.shape[1]
        if pre is None:
            pre = ranks.sum(ctx.gene_moments(0, C, w))
        W, S_C, G = ranks.sum(float(wv.sum()), C.T @ wv, (C * wv[:, None]).T @ C)
        W = float(W)
        M0, M1, Mk = pre[:, 0], pre[:, 1], pre[:, 2:2 + k]
        bs = B @ S_C
        bgb = np.einsum("gi,ij,gj->g", B, G, B)
        v_mean = (M0 + bs + W * mu) / W
        d = mu - v_mean
        v_var = (bgb + 2.0 * d * bs + W * d * d + 2.0 * (Mk * B).sum(axis=1) + 2.0 * d * M0 + M1) / W
        scale = mu * mu + v_mean * v_mean + (M1 + np.abs(bgb)) / W
    else:
        if pre is None:
            pre = ranks.sum(ctx.gene_moments(0, None, w))
        W = float(ranks.sum(float(wv.sum())))
        v_mean = pre[:, 0] / W
        v_var = pre[:, 1] / W - v_mean ** 2
        scale = pre[:, 1] / W
    t0 = _tick("gene_moments", t0)
    if implicit_cov_corr:
        shift = np.expm1(mu + B @ (S_C / W))
        D = ranks.sum(ctx.gene_expm1_implicit(shift, w))
        ct_mean = shift + D[:, 0] / W
        ct_var = D[:, 1] / W - (D[:, 0] / W) ** 2
        ct_scale = D[:, 1] / W
    else:
        E = ranks.sum(ctx.gene_moments(1, None, w))
        ct_mean = E[:, 0] / W
        ct_var = E[:, 1] / W - ct_mean ** 2
        ct_scale = E[:, 1] / W
    t0 = _tick("gene_expm1_moments", t0)

    # numerically constant genes: the reference's variance there is rounding noise of numpy's summation order
    center = w is None and (implicit_cov_corr or not sparse_x)
    for transform, vm, vv, sc in ((0, v_mean, v_var, scale), (1, ct_mean, ct_var, ct_scale)):
        with np.errstate(invalid="ignore"):
            flag = ~(vv > np.maximum(_FLAG_REL * sc, 1e-10 * vm * vm)) & (sc > 0)
        genes = np.flatnonzero(flag)
        if genes.size == 0:
            continue
        gs, gq = np.zeros(n_gene), np.zeros(n_gene)
        if not ranks.on:
            idx = ctx.gene_stats_exact(transform, genes, gs, gq, cell_weight=w, center=center, denom=W)
            m = gs[idx] / W
        elif center:
            idx = ctx.gene_stats_exact(transform, genes, gs, gq, cell_weight=w, center=False, denom=W)
            m_all = np.zeros(n_gene)
            m_all[idx] = ranks.sum(gs[idx]) / W
            ctx.gene_stats_exact(transform, genes, gs, gq, cell_weight=w, center=True, denom=W, mean_in=m_all)
            gq[idx] = ranks.sum(gq[idx])
            m = m_all[idx]
        else:
            idx = ctx.gene_stats_exact(transform, genes, gs, gq, cell_weight=w, center=False, denom=W)
            gs[idx], gq[idx] = ranks.sum(gs[idx], gq[idx])
            m = gs[idx] / W
        vm[idx] = m
        vv[idx] = gq[idx] / W if center else gq[idx] / W - m ** 2
    _tick("exact_order_genes", t0)
    return v_mean, v_var, ct_mean, ct_var


def compute_stats(adata, implicit_cov_corr=False, cell_weight=None, n_mean_bin=20, n_var_bin=20, n_chunk=20,
                  _ranks=None, _pre=None, _keep_timing=False):
    """Gene- and cell-level statistics (drop-in for scdrs/pp.py:276-419).

    Returns (df_gene [mean, var, var_tech, ct_mean, ct_var, ct_var_tech, mean_var], df_cell [mean, var]).
    """
    import time

    if not _keep_timing:
        LAST_TIMING.clear()
    t0 = time.perf_counter()
    ranks = _ranks if _ranks is not None else _Ranks(False)
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

    if _exact_order() and not ranks.on:
        df_gene["mean"], df_gene["var"] = _device_mean_var(adata, implicit_cov_corr, 0, w, transform=0)       # :353-355/:368-370
        df_gene["ct_mean"], df_gene["ct_var"] = _device_mean_var(adata, implicit_cov_corr, 0, w, transform=1)  # :359-364/:371-373
        t0 = _tick("gene_moments", t0)
    else:
        df_gene["mean"], df_gene["var"], df_gene["ct_mean"], df_gene["ct_var"] = _gene_mean_var_fast(
            adata, implicit_cov_corr, w, ranks, _pre)
        t0 = time.perf_counter()

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
    t0 = _tick("loess", t0)

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
    t0 = _tick("mean_var_bins", t0)

    df_cell["mean"], df_cell["var"] = _device_mean_var(adata, implicit_cov_corr, 1)                        # :409-417
    _tick("cell_stats", t0)
    return df_gene, df_cell


def _get_mean_var(sparse_X, axis=0, weights=None):
    """Mean and variance of a (sparse or dense) matrix (drop-in for scdrs/pp.py:467-517), computed
    on the device from one pass over its non-zeros."""
    from .anndata_lite import AnnData

    X = sparse_X if sparse.issparse(sparse_X) else np.asarray(sparse_X)
    tmp = AnnData(X)
    tmp.uns["SCDRS_PARAM"] = {"FLAG_SPARSE": True, "FLAG_COV": False}
    if axis == 0:
        w0 = None if weights is None else np.asarray(weights, dtype=np.float64)
        if _exact_order():
            return _device_mean_var(tmp, False, 0, w0)
        return _gene_mean_var_fast(tmp, False, w0, _Ranks(False))[:2]
    if weights is None:
        return _device_mean_var(tmp, False, 1)
    # weights over genes: transpose so that they become cell weights
    tmpT = AnnData(sparse.csr_matrix(X.T) if sparse.issparse(X) else np.ascontiguousarray(X.T))
    tmpT.uns["SCDRS_PARAM"] = {"FLAG_SPARSE": True, "FLAG_COV": False}
    if _exact_order():
        return _device_mean_var(tmpT, False, 0, np.asarray(weights, dtype=np.float64))
    return _gene_mean_var_fast(tmpT, False, np.asarray(weights, dtype=np.float64), _Ranks(False))[:2]


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
    if axis == 1 or _exact_order():
        return _device_mean_var(adata, True, axis, w, transform=transform)
    res = _gene_mean_var_fast(adata, True, w, _Ranks(False))
    return (res[0], res[1]) if transform == 0 else (res[2], res[3])
