"""I/O glue on either side of the hot path: same behaviour as the loaders in ``scdrs/util.py`` of
the reference (``load_h5ad :46-92``, ``load_gs :194-274``, ``load_homolog_mapping :152-191``,
``save_gs :277-308``).  ``anndata`` / ``scanpy`` are absent from this image, so reading goes through
``anndata_lite`` and the scanpy preprocessing steps are restated on scipy.sparse."""
import os
from typing import Dict, List

import numpy as np
import pandas as pd
from scipy import sparse

from .anndata_lite import AnnData, read_h5ad


def convert_species_name(species):
    if species in ["Mouse", "mouse", "Mus_musculus", "mus_musculus", "mmusculus"]:
        return "mmusculus"
    if species in ["Human", "human", "Homo_sapiens", "homo_sapiens", "hsapiens"]:
        return "hsapiens"
    raise ValueError("species name '%s' is not supported" % species)


def str_or_list_like(x):
    """'str', 'list_like' or 'others' by duck typing (scdrs/util.py:26-43)."""
    if hasattr(x, "strip"):
        return "str"
    if hasattr(x, "__getitem__") or hasattr(x, "__iter__"):
        return "list_like"
    return "others"


def load_h5ad(h5ad_file, flag_filter_data=False, flag_raw_count=True, min_genes=250, min_cells=50):
    """Load an h5ad file, check it, optionally filter and normalise it (scdrs/util.py:46-92).

    The scanpy calls of the reference are restated: ``filter_cells(min_genes)`` keeps cells with
    at least ``min_genes`` expressed genes (adds ``obs["n_genes"]``); ``filter_genes(min_cells)``
    keeps genes expressed in at least ``min_cells`` cells (adds ``var["n_cells"]``);
    ``normalize_per_cell(counts_per_cell_after=1e4)`` drops cells with fewer than 1 count, stores
    ``obs["n_counts"]`` and scales each cell to 1e4 counts; ``log1p`` is the natural log.

    Sparse matrices take the device path (``scdrs_preprocess_counts``): the raw matrix is uploaded
    once, checked, filtered and normalised on the GPU, and stays resident for ``preprocess`` and
    ``score_cell`` (no second upload); the processed matrix is also downloaded into ``adata.X``.  The
    host restatement of the same steps lives in ``oracle/h5ad_oracle.py`` and is what the tests
    compare with.  Dense matrices (not a scDRS use case at scale) are handled with numpy.
    """
    adata = read_h5ad(h5ad_file)
    return preprocess_counts(adata, flag_filter_data, flag_raw_count, min_genes, min_cells)


_ERR_NAN = "h5ad expression matrix should not contain NaN. Please impute them beforehand."
_ERR_NEG = (
    "h5ad expression matrix should not contain negative values. "
    "This is because in the preprocessing step, "
    "scDRS models the gene-level log mean-variance relationship. "
    "See scdrs.pp.compute_stats for details."
)


def preprocess_counts(adata, flag_filter_data=False, flag_raw_count=True, min_genes=250, min_cells=50):
    """The checks / filters / normalisation of ``load_h5ad`` on an AnnData already in memory."""
    if not sparse.issparse(adata.X):
        return _preprocess_counts_dense(adata, flag_filter_data, flag_raw_count, min_genes, min_cells)
    from . import _device, _native

    indptr, indices, data = _device.to_csr_parts(adata.X)
    ctx = _native.Context(_device._DEFAULT_DEVICE[0])
    try:
        ctx.set_csr_host(indptr, indices, data, adata.shape[1])
        out = ctx.preprocess_counts(flag_filter_data, min_genes, min_cells, flag_raw_count)
        if out["has_nan"]:
            raise ValueError(_ERR_NAN)
        if out["has_negative"]:
            raise ValueError(_ERR_NEG)
        if flag_filter_data or flag_raw_count:
            keep_c, keep_g = out["keep_cell"].astype(bool), out["keep_gene"].astype(bool)
            ip, ix, da = ctx.get_csr(out["nnz"])
            X = sparse.csr_matrix((da, ix, ip), shape=(out["n_cell"], out["n_gene"]))
            obs, var = adata.obs.loc[keep_c].copy(), adata.var.loc[keep_g].copy()
            if flag_filter_data:
                obs["n_genes"] = out["n_genes"][keep_c]
                var["n_cells"] = out["n_cells"][keep_g]
            if flag_raw_count:
                obs["n_counts"] = out["n_counts"][keep_c]
            rows = np.nonzero(keep_c)[0]
            adata = AnnData(X, obs, var, adata.uns, {k: v[rows][:, rows] for k, v in getattr(adata, "obsp", {}).items()})
        elif not (isinstance(adata.X, sparse.csr_matrix) and adata.X.dtype == np.float32):
            adata = AnnData(sparse.csr_matrix((data, indices, indptr), shape=adata.shape), adata.obs, adata.var, adata.uns,
                            dict(getattr(adata, "obsp", {})))
    except BaseException:
        ctx.close()
        raise
    _device.adopt_context(adata, ctx)      # the matrix is already resident: no upload in preprocess / score_cell
    return adata


def _preprocess_counts_dense(adata, flag_filter_data, flag_raw_count, min_genes, min_cells):
    X = np.asarray(adata.X)
    if np.isnan(X.sum()):
        raise ValueError(_ERR_NAN)
    if (X < 0).sum() > 0:
        raise ValueError(_ERR_NEG)
    if flag_filter_data:
        n_genes = (X != 0).sum(axis=1)
        keep = n_genes >= min_genes
        adata = adata[keep, :]
        adata.obs["n_genes"] = n_genes[keep]
        n_cells = (np.asarray(adata.X) != 0).sum(axis=0)
        keepg = n_cells >= min_cells
        adata = adata[:, keepg]
        adata.var["n_cells"] = n_cells[keepg]
    if flag_raw_count:
        counts = np.asarray(adata.X).sum(axis=1)
        keep = counts >= 1
        if not keep.all():
            adata = adata[keep, :]
            counts = counts[keep]
        adata.obs["n_counts"] = counts
        scale = (1e4 / counts).astype(np.float32)
        adata = AnnData(np.log1p(np.asarray(adata.X, dtype=np.float32) * scale[:, None]), adata.obs, adata.var, adata.uns,
                        adata.obsp)
    return adata


def load_scdrs_score(score_file, obs_names=None):
    """Load ``.full_score.gz`` files keyed by trait; "@" in the file name is a wildcard (drop-in for
    scdrs/util.py:95-149).  Files sharing fewer than 10% of ``obs_names`` are skipped."""
    import fnmatch

    assert score_file.endswith("full_score.gz"), "Expect scDRS .full_score.gz files for score_file"
    score_file_pattern = score_file.split(os.path.sep)[-1]
    score_dir = score_file.replace(os.path.sep + score_file_pattern, "")
    score_file_list = [x for x in os.listdir(score_dir) if fnmatch.fnmatch(x, score_file_pattern.replace("@", "*"))]
    print("Score file folder: %s" % score_dir)
    print("Find %d score files: %s" % (len(score_file_list), ",".join(score_file_list)))
    dict_score = {}
    for name in score_file_list:
        temp_df = pd.read_csv(score_dir + os.path.sep + name, sep="\t", index_col=0)
        temp_df.index = [str(x) for x in temp_df.index]
        if obs_names is not None:
            n_cell_overlap = len(set(obs_names) & set(temp_df.index))
            if n_cell_overlap < 0.1 * len(obs_names):
                print("WARNING: %s skipped, containing %d/%d target cells" % (name, n_cell_overlap, len(obs_names)))
                continue
        dict_score[name.replace(".full_score.gz", "")] = temp_df
    return dict_score


def load_homolog_mapping(src_species: str, dst_species: str) -> dict:
    """Mouse <-> human gene symbol homologs (scdrs/util.py:152-191)."""
    src_species = convert_species_name(src_species)
    dst_species = convert_species_name(dst_species)
    assert src_species != dst_species, "src and dst cannot be the same"
    df_hom = pd.read_csv(
        os.path.join(os.path.dirname(__file__), "data/mouse_human_homologs.txt"), sep="\t"
    )
    if (src_species == "hsapiens") & (dst_species == "mmusculus"):
        return dict(zip(df_hom["HUMAN_GENE_SYM"], df_hom["MOUSE_GENE_SYM"]))
    if (src_species == "mmusculus") & (dst_species == "hsapiens"):
        return dict(zip(df_hom["MOUSE_GENE_SYM"], df_hom["HUMAN_GENE_SYM"]))
    raise NotImplementedError(f"gene conversion from {src_species} to {dst_species} is not supported")


def load_gs(gs_path: str, src_species: str = None, dst_species: str = None,
            to_intersect: List[str] = None) -> dict:
    """Parse a ``.gs`` file into ``{trait: (gene_list, gene_weight_list)}`` (scdrs/util.py:194-274):
    ``g1,g2`` (weights 1.0) or ``g1:w1,g2:w2``; later duplicates win; optional homolog mapping and
    intersection with ``to_intersect``; file order preserved."""
    assert (src_species is None) == (
        dst_species is None
    ), "src_species and dst_species must be both None or not None"
    dict_map = None
    if (src_species is not None) and (dst_species is not None) and (src_species != dst_species):
        dict_map = load_homolog_mapping(src_species, dst_species)
    keep = set(to_intersect) if to_intersect is not None else None
    dict_gs = {}
    df_gs = pd.read_csv(gs_path, sep="\t")
    for trait, gs in zip(df_gs.iloc[:, 0], df_gs.iloc[:, 1]):
        fields = [g.split(":") for g in gs.split(",")]
        if all(len(g) == 1 for g in fields):
            weights = {g[0]: 1.0 for g in fields}
        elif all(len(g) == 2 for g in fields):
            weights = {g[0]: float(g[1]) for g in fields}
        else:
            raise ValueError(f"gene set {trait} contains genes with invalid format")
        if dict_map is not None:
            weights = {dict_map[g]: w for g, w in weights.items() if g in dict_map}
        if keep is not None:
            weights = {g: w for g, w in weights.items() if g in keep}
        genes = list(weights.keys())
        dict_gs[trait] = (genes, [weights[g] for g in genes])
    return dict_gs


def save_gs(gs_path: str, dict_gs: dict) -> None:
    """Write a ``.gs`` file (scdrs/util.py:277-308)."""
    rows: Dict = {"TRAIT": [], "GENESET": []}
    for trait, val in dict_gs.items():
        rows["TRAIT"].append(trait)
        if isinstance(val, tuple):
            rows["GENESET"].append(",".join([g + ":" + str(w) for g, w in zip(*val)]))
        else:
            rows["GENESET"].append(",".join(val))
    pd.DataFrame(rows).to_csv(gs_path, sep="\t", index=False)


def write_score_file(path, df, n_thread=None):
    """``df.to_csv(path, sep="\\t", index=True, compression="gzip")`` for an all-float32 frame, with row blocks
    formatted and deflated in parallel (``scdrs_write_score_gz``); other frames fall back to pandas."""
    from . import _native

    if len(df.columns) and all(dt == np.float32 for dt in df.dtypes):
        _native.write_score_gz(path, df, n_thread=n_thread)
    else:
        df.to_csv(path, sep="\t", index=True, compression="gzip")


def zsc2pval(zsc):
    """One-sided p-value of a z-score (scdrs/util.py:430-435)."""
    from scipy.stats import norm

    return norm.cdf(-np.asarray(zsc, dtype=np.float64))


def pval2zsc(pval):
    """z-score of a one-sided p-value (scdrs/util.py:438-442)."""
    from scipy.stats import norm

    return -norm.ppf(np.asarray(pval, dtype=np.float64))


def fdr_bh(pvals):
    """Benjamini-Hochberg adjusted p-values (what ``multipletests(method='fdr_bh')[1]`` returns;
    used only for the per-trait log line of the CLI, bin/scdrs:289-303)."""
    p = np.asarray(pvals, dtype=np.float64)
    n = p.shape[0]
    order = np.argsort(p)
    adj = p[order] * n / np.arange(1, n + 1)
    adj = np.minimum.accumulate(adj[::-1])[::-1]
    out = np.empty(n)
    out[order] = np.minimum(adj, 1.0)
    return out
