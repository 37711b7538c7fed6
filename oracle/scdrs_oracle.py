"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy/pandas) of the scDRS scoring hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module, and only as the checker / the CPU arm.  The product
(``scdrs_b200``) never imports it and has no CPU fallback.

Parity pinning (not "unpinned"):
  * against the reference's golden score files for the bundled toy data
    (``tests/golden/toy/*.score.gz`` -- byte copies of /root/reference/scdrs/data fixtures), and
  * against outputs of the UNMODIFIED reference sources (loaded through ``oracle/ref_shim.py``)
    on seeded synthetic inputs, committed under ``tests/golden/`` by
    ``tests/golden/make_golden.py``.

Every function cites the reference lines (relative to /root/reference) it restates.  The
restatement works on integer gene indices instead of gene-name labels; the random stream is
numpy's own legacy ``RandomState`` so that control sets are bit-identical.
"""
import numpy as np
import pandas as pd
import scipy as sp
from scipy import sparse
import scipy.stats  # noqa: F401

LEGAL_WEIGHT_OPT = ("uniform", "vs", "inv_std")


# ----------------------------------------------------------------------------------------------
# control gene sets                                                     scdrs/method.py:229-309
# ----------------------------------------------------------------------------------------------
def gene_bins(df_gene, ctrl_match_key, n_genebin):
    """Bins of row positions of ``df_gene`` in the reference's iteration order.

    Restates scdrs/method.py:284-290: a key with fewer unique values than n_gene/10 is
    categorical (group by value), otherwise ``pd.qcut`` into ``n_genebin`` quantile bins.
    Bins come back in pandas-groupby order (sorted keys, NaN dropped); inside a bin the gene
    positions are ordered by gene NAME, which is what ``sorted(set_of_names)`` yields
    (scdrs/method.py:296).
    """
    names = df_gene["gene"].values if "gene" in df_gene else df_gene.index.values
    key = df_gene[ctrl_match_key]
    if key.unique().shape[0] < df_gene.shape[0] / 10:
        labels = key
    else:
        labels = pd.qcut(key, q=n_genebin, labels=False, duplicates="drop")
    pos = pd.Series(np.arange(df_gene.shape[0]), index=df_gene.index)
    bins = []
    for _, grp in pos.groupby(labels.values, sort=True):
        p = grp.values
        order = sorted(range(len(p)), key=lambda i: names[p[i]])
        bins.append(p[order])
    return bins


def select_ctrl_geneset(df_gene, trait_pos, trait_weight, ctrl_match_key, n_ctrl, n_genebin, seed):
    """Matched control sets as integer positions.  Restates scdrs/method.py:276-309.

    trait_pos: positions (rows of df_gene) of the trait genes, ordered by gene name
    (scdrs/method.py:158).  Returns (ctrl_pos[n_ctrl, n_S], ctrl_weight[n_ctrl, n_S]).

    ``np.random.choice(names, size=k, replace=False)`` on the legacy global generator draws
    ``permutation(len(names))[:k]`` (numpy/random/mtrand.pyx, ``choice`` with p=None), so a
    ``RandomState(seed)`` walking the bins in the same order reproduces it draw for draw.
    """
    rng = np.random.RandomState(seed)
    w_of = dict(zip(trait_pos.tolist(), list(trait_weight)))
    names = df_gene["gene"].values if "gene" in df_gene else df_gene.index.values
    is_trait = np.zeros(df_gene.shape[0], dtype=bool)
    is_trait[trait_pos] = True
    ctrl_pos = [[] for _ in range(n_ctrl)]
    ctrl_w = [[] for _ in range(n_ctrl)]
    for members in gene_bins(df_gene, ctrl_match_key, n_genebin):
        hit = members[is_trait[members]]  # already in name order
        if len(hit) == 0:
            continue
        hit_w = [w_of[p] for p in hit.tolist()]
        for i in range(n_ctrl):
            pick = rng.permutation(len(members))[: len(hit)]
            ctrl_pos[i].extend(members[pick].tolist())
            ctrl_w[i].extend(hit_w)
    del names
    return np.array(ctrl_pos, dtype=np.int64), np.array(ctrl_w, dtype=np.float64)


# ----------------------------------------------------------------------------------------------
# raw score                                                             scdrs/method.py:312-402
# ----------------------------------------------------------------------------------------------
def score_weights(gene_stats, pos, gene_weight, weight_opt):
    """Normalised per-gene weights.  Restates scdrs/method.py:366-375."""
    assert weight_opt in {"uniform", "vs", "inv_std", "od"}, (
        "weight_opt=%s is not one of {'uniform', 'vs', 'inv_std', 'od}'" % weight_opt
    )
    if weight_opt == "od":
        raise NotImplementedError("weight_opt='od' is a benchmarking-only option outside the hot path")
    if weight_opt == "uniform":
        w = np.ones(len(pos))
    elif weight_opt == "vs":
        w = 1 / np.sqrt(gene_stats["var_tech"].values[pos].astype(np.float64) + 1e-2)
    else:
        w = 1 / np.sqrt(gene_stats["var"].values[pos].astype(np.float64) + 1e-2)
    w = w * np.asarray(gene_weight, dtype=np.float64)
    return w / w.sum()


def compute_raw_score(X, pos, w, cov_mat=None, cov_beta=None, gene_mean=None):
    """``X[:, pos] @ w`` (+ implicit covariate terms).  Restates scdrs/method.py:377-400.

    X: scipy CSC/CSR matrix or ndarray (n_cell, n_gene); cov_mat (n_cell, k);
    cov_beta (n_gene, k) -- the stored COV_BETA, i.e. minus the regression coefficients;
    gene_mean (n_gene,).
    """
    v = np.asarray(X[:, pos].dot(w)).reshape(-1)
    if cov_mat is not None:
        v = v + cov_mat @ (cov_beta[pos].T @ w) + gene_mean[pos] @ w
    return v


# ----------------------------------------------------------------------------------------------
# background correction                                                 scdrs/method.py:482-598
# ----------------------------------------------------------------------------------------------
def correct_background(v_raw, mat_ctrl, var_ratio, inject=None):
    """Restates scdrs/method.py:521-583 (the optional text dumps are not part of the path).

    ``inject`` (tests of a SLICE of a larger atlas only): the statistics the reference takes over all cells --
    ``raw_mean`` [1 + n_ctrl] (:526-527), ``norm_mean`` [1 + n_ctrl] (:564-565), ``lo`` (:581) -- supplied from
    the full run, column 0 being the trait; everything cell-wise is computed here as the reference does."""
    v_raw = np.asarray(v_raw, dtype=np.float64)
    mat_ctrl = np.asarray(mat_ctrl, dtype=np.float64)
    zero_t = v_raw == 0                                   # :522
    zero_c = mat_ctrl == 0                                # :523
    raw_mean_t = v_raw.mean() if inject is None else inject["raw_mean"][0]
    raw_mean_c = mat_ctrl.mean(axis=0) if inject is None else inject["raw_mean"][1:]
    t = v_raw - raw_mean_t                                # :526
    c = (mat_ctrl - raw_mean_c) / np.sqrt(var_ratio)      # :527-528
    mu = c.mean(axis=1)                                   # :544
    sd = c.std(axis=1)                                    # :545
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (t - mu) / sd                                 # :547
        c = ((c.T - mu) / sd).T                           # :548
    t = t - (t.mean() if inject is None else inject["norm_mean"][0])          # :564
    c = c - (c.mean(axis=0) if inject is None else inject["norm_mean"][1:])   # :565
    lo = min(t.min(), c.min()) if inject is None else inject["lo"]            # :581
    t[zero_t] = lo - 1e-3                                 # :582
    c[zero_c] = lo                                        # :583
    return t, c


# ----------------------------------------------------------------------------------------------
# empirical-null p-value                                                scdrs/method.py:601-632
# ----------------------------------------------------------------------------------------------
def null_ge_count(v_t, v_null):
    """#{null >= t_i}: the integer the p-value is made of (N - lower_bound)."""
    s = np.sort(np.asarray(v_null))
    return s.shape[0] - np.searchsorted(s, np.asarray(v_t), side="left")


def get_p_from_empi_null(v_t, v_null):
    """Restates scdrs/method.py:626-632: p = (1 + #{null >= t}) / (1 + N)."""
    n = np.asarray(v_null).shape[0]
    return (null_ge_count(v_t, v_null) + 1) / (n + 1)


# ----------------------------------------------------------------------------------------------
# score_cell                                                            scdrs/method.py:12-226
# ----------------------------------------------------------------------------------------------
def score_cell(
    adata,
    gene_list,
    gene_weight=None,
    ctrl_match_key="mean_var",
    n_ctrl=1000,
    n_genebin=200,
    weight_opt="vs",
    return_ctrl_raw_score=False,
    return_ctrl_norm_score=False,
    random_seed=0,
    return_internals=False,
    inject=None,
):
    """Restates scdrs/method.py:96-226 on top of the helpers above.  ``inject``: see ``correct_background`` (the
    pooled p-value is then taken against the slice's own null and is only meaningful to the caller as such)."""
    param = adata.uns["SCDRS_PARAM"]
    n_cell, n_gene = adata.shape
    df_gene = param["GENE_STATS"].loc[adata.var_names].copy()      # :146
    df_gene["gene"] = df_gene.index
    df_gene = df_gene.drop_duplicates(subset="gene")                # :148
    pos_of = {g: i for i, g in enumerate(df_gene["gene"].values)}
    col_of = pd.Index(adata.var_names).get_indexer(df_gene["gene"].values)

    gene_list = list(gene_list)
    gene_weight = [1] * len(gene_list) if gene_weight is None else list(gene_weight)
    w_of = dict(zip(gene_list, gene_weight))                        # :157 (last duplicate wins)
    genes = sorted(set(gene_list) & set(pos_of))                    # :158
    trait_pos = np.array([pos_of[g] for g in genes], dtype=np.int64)
    trait_gw = np.array([w_of[g] for g in genes], dtype=np.float64)

    ctrl_pos, ctrl_gw = select_ctrl_geneset(
        df_gene, trait_pos, trait_gw, ctrl_match_key, n_ctrl, n_genebin, random_seed
    )

    X = adata.X
    if sparse.issparse(X) and not isinstance(X, sparse.csc_matrix):
        X = sparse.csc_matrix(X)                                    # :98-99 (not written back)
    cov = {}
    if param["FLAG_SPARSE"] and param["FLAG_COV"]:                  # :377-385
        cov_cols = list(param["COV_MAT"])
        cov = dict(
            cov_mat=param["COV_MAT"].loc[list(adata.obs_names), cov_cols].values.astype(np.float64),
            cov_beta=param["COV_BETA"].loc[df_gene["gene"].values, cov_cols].values.astype(np.float64),
            gene_mean=param["COV_GENE_MEAN"].loc[df_gene["gene"].values].values.astype(np.float64),
        )
        # cov_beta / gene_mean above are indexed by df_gene position; X by column -> remap X
    def raw(pos, gw):
        w = score_weights(df_gene, pos, gw, weight_opt)
        v = np.asarray(X[:, col_of[pos]].dot(w)).reshape(-1)
        if cov:
            v = v + cov["cov_mat"] @ (cov["cov_beta"][pos].T @ w) + cov["gene_mean"][pos] @ w
        return v, w

    v_raw, v_w = raw(trait_pos, trait_gw)                           # :172
    mat_raw = np.zeros((n_cell, n_ctrl))
    mat_w = np.zeros((len(genes), n_ctrl))
    for i in range(n_ctrl):                                         # :178-183
        mat_raw[:, i], mat_w[:, i] = raw(ctrl_pos[i], ctrl_gw[i])

    ratio = np.ones(n_ctrl)                                         # :186-195
    if ctrl_match_key == "mean_var" and weight_opt in LEGAL_WEIGHT_OPT:
        var = df_gene["var"].values.astype(np.float64)
        for i in range(n_ctrl):
            ratio[i] = (var[ctrl_pos[i]] * mat_w[:, i] ** 2).sum()
        ratio /= (var[trait_pos] * v_w ** 2).sum()

    v_norm, mat_norm = correct_background(v_raw, mat_raw, ratio, inject)    # :197
    mc_count = (mat_norm.T >= v_norm).sum(axis=0)                   # :205
    mc_p = (1 + mc_count) / (1 + n_ctrl)
    null = mat_norm.flatten()
    ge = null_ge_count(v_norm, null)                                # :206
    pooled_p = (ge + 1) / (null.shape[0] + 1)
    with np.errstate(divide="ignore"):
        nlog10 = -np.log10(pooled_p)                                # :207
    z = -sp.stats.norm.ppf(pooled_p).clip(min=-10, max=10)          # :208

    res = {
        "raw_score": v_raw, "norm_score": v_norm, "mc_pval": mc_p,
        "pval": pooled_p, "nlog10_pval": nlog10, "zscore": z,
    }
    if return_ctrl_raw_score:
        for i in range(n_ctrl):
            res["ctrl_raw_score_%d" % i] = mat_raw[:, i]
    if return_ctrl_norm_score:
        for i in range(n_ctrl):
            res["ctrl_norm_score_%d" % i] = mat_norm[:, i]
    df = pd.DataFrame(index=adata.obs.index, data=res, dtype=np.float32)   # :225
    if return_internals:
        return df, dict(
            ctrl_cols=col_of[ctrl_pos], trait_cols=col_of[trait_pos], mc_count=mc_count,
            ge_count=ge, ratio=ratio, mat_raw=mat_raw, mat_norm=mat_norm, v_raw=v_raw,
            v_norm=v_norm, trait_w=v_w, ctrl_w=mat_w, ctrl_gw_raw=ctrl_gw[0] if n_ctrl > 0 else np.zeros(0),
        )
    return df


# ----------------------------------------------------------------------------------------------
# preprocessing statistics                                                 scdrs/pp.py:63-641
# ----------------------------------------------------------------------------------------------
def mean_var(X, axis=0, weights=None):
    """Restates _get_mean_var (scdrs/pp.py:467-517) in float64 on a dense copy."""
    A = np.asarray(X.toarray() if sparse.issparse(X) else X, dtype=np.float64)
    if weights is None:
        return A.mean(axis=axis), A.var(axis=axis)
    w = np.asarray(weights, dtype=np.float64)
    m = np.average(A, axis=axis, weights=w)
    return m, np.average(A * A, axis=axis, weights=w) - m ** 2


def corrected_matrix(adata):
    """Dense ``X + COV_MAT @ COV_BETA.T + COV_GENE_MEAN`` (scdrs/pp.py:556, :613-617) or X itself."""
    param = adata.uns.get("SCDRS_PARAM", {})
    X = adata.X
    A = np.asarray(X.toarray() if sparse.issparse(X) else X, dtype=np.float64)
    if param.get("FLAG_SPARSE") and param.get("FLAG_COV"):
        cols = list(param["COV_MAT"])
        C = param["COV_MAT"].loc[list(adata.obs_names), cols].values.astype(np.float64)
        B = param["COV_BETA"].loc[list(adata.var_names), cols].values.astype(np.float64)
        mu = param["COV_GENE_MEAN"].loc[list(adata.var_names)].values.astype(np.float64)
        A = A + C @ B.T + mu
    return A


def compute_stats(adata, implicit_cov_corr=False, cell_weight=None, n_mean_bin=20, n_var_bin=20, loess_impl=None):
    """Restates compute_stats (scdrs/pp.py:337-419) on a dense float64 copy of the corrected matrix
    (small inputs only).  ``loess_impl`` stands in for skmisc.loess.loess."""
    A = corrected_matrix(adata) if implicit_cov_corr else np.asarray(
        adata.X.toarray() if sparse.issparse(adata.X) else adata.X, dtype=np.float64)
    df_gene = pd.DataFrame(index=adata.var_names,
                           columns=["mean", "var", "var_tech", "ct_mean", "ct_var", "ct_var_tech", "mean_var"])
    df_cell = pd.DataFrame(index=adata.obs_names, columns=["mean", "var"])
    df_gene["mean"], df_gene["var"] = mean_var(A, 0, cell_weight)                       # :353/:368
    df_gene["ct_mean"], df_gene["ct_var"] = mean_var(np.expm1(A), 0, cell_weight)       # :359-364/:371
    ct_mean, ct_var = df_gene["ct_mean"].values.astype(float), df_gene["ct_var"].values.astype(float)
    not_const = ct_var > 0                                                               # :376-380
    if (ct_mean <= 0).sum() > 0:
        not_const = not_const & (ct_mean > 0)
    est = np.zeros(A.shape[1])
    model = loess_impl(np.log10(ct_mean[not_const]), np.log10(ct_var[not_const]), span=0.3, degree=2)
    model.fit()
    est[not_const] = model.outputs.fitted_values                                         # :384-386
    df_gene["ct_var_tech"] = 10 ** est
    with np.errstate(divide="ignore", invalid="ignore"):
        vt = df_gene["var"].values.astype(float) * df_gene["ct_var_tech"].values / ct_var   # :389
    vt[np.isnan(vt)] = 0                                                                 # :390
    df_gene["var_tech"] = vt
    n_bin_max = int(np.floor(np.sqrt(A.shape[1] / 10)))                                  # :393-399
    if n_mean_bin > n_bin_max or n_var_bin > n_bin_max:
        n_mean_bin = n_var_bin = n_bin_max
    mbin = pd.qcut(df_gene["mean"].astype(float), n_mean_bin, labels=False, duplicates="drop")   # :400
    labels = np.full(A.shape[1], "", dtype=object)
    for b in sorted(set(mbin)):                                                          # :402-407
        sel = (mbin == b).values
        vbin = pd.qcut(df_gene.loc[sel, "var"].astype(float), n_var_bin, labels=False, duplicates="drop")
        labels[sel] = ["%d.%d" % (b, x) for x in vbin]
    df_gene["mean_var"] = labels
    df_cell["mean"], df_cell["var"] = mean_var(A, 1)                                     # :409-417
    return df_gene, df_cell


def reg_out(Y, X):
    """Restates reg_out (scdrs/pp.py:425-464)."""
    X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
    Y2 = np.asarray(Y, dtype=np.float64).reshape(len(Y), -1)
    n = Y2.shape[0]
    coef = np.linalg.solve(X.T @ X / n, X.T @ Y2 / n)
    R = Y2 - X @ coef
    return R.reshape(-1) if R.shape[1] == 1 else R


def preprocess(adata, cov=None, adj_prop=None, n_mean_bin=20, n_var_bin=20, loess_impl=None):
    """Restates preprocess (scdrs/pp.py:165-271); numeric covariates only; writes SCDRS_PARAM."""
    n_cell, n_gene = adata.shape
    flag_sparse, flag_cov = sparse.issparse(adata.X), cov is not None
    adata.uns["SCDRS_PARAM"] = {"FLAG_SPARSE": flag_sparse, "FLAG_COV": flag_cov,
                                "FLAG_ADJ_PROP": adj_prop is not None}
    if flag_sparse:
        adata.X = sparse.csr_matrix(adata.X)
    if flag_cov:
        df_cov = pd.DataFrame(index=adata.obs_names).join(cov)                          # :195-198
        df_cov = df_cov.fillna(df_cov.mean())
        if (reg_out(np.ones(n_cell), df_cov.values) ** 2).mean() > 0.01:                # :201-203
            df_cov["SCDRS_CONST"] = 1
        C = df_cov.values.astype(np.float64)
        gmean = np.asarray(adata.X.mean(axis=0)).reshape(-1)                            # :206
        if flag_sparse:
            beta = np.linalg.solve(C.T @ C / n_cell, np.asarray((adata.X.T @ C).T) / n_cell)   # :209-212
            adata.uns["SCDRS_PARAM"]["COV_MAT"] = df_cov
            adata.uns["SCDRS_PARAM"]["COV_BETA"] = pd.DataFrame(-beta.T, index=adata.var_names, columns=df_cov.columns)
            adata.uns["SCDRS_PARAM"]["COV_GENE_MEAN"] = pd.Series(gmean, index=adata.var_names)
        else:
            adata.X = reg_out(np.asarray(adata.X), C) + gmean                           # :223-224
    w = None
    if adj_prop is not None:                                                             # :242-257
        size = adata.obs[adj_prop].value_counts()
        size = size.clip(lower=int(0.01 * size.max()))      # the docstring's floor (:93-94)
        w = (n_cell / size.reindex(adata.obs[adj_prop].values).values).astype(np.float64)
        w = w / w.mean()
    df_gene, df_cell = compute_stats(adata, bool(flag_sparse and flag_cov), w, n_mean_bin, n_var_bin, loess_impl)
    adata.uns["SCDRS_PARAM"]["GENE_STATS"] = df_gene
    adata.uns["SCDRS_PARAM"]["CELL_STATS"] = df_cell
