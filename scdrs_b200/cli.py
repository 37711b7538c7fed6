"""``scdrs compute-score`` and ``scdrs perform-downstream`` (drop-ins for bin/scdrs:25-304 and :558-816 of
the reference).

python-fire is not installed here, so the flag conventions it gives the reference's CLI are
restated: ``compute-score`` / ``compute_score``, ``--h5ad-file`` / ``--h5ad_file``,
``--flag True`` / ``--flag=False``.  ``compute-score`` is the hot path, ``perform-downstream`` the step right
after it; ``munge-gs`` is outside this repository's scope.
"""
import os
import sys
import time

import pandas as pd

import scdrs_b200
from scdrs_b200 import util

_SIGNATURE = [
    ("h5ad_file", str, None), ("h5ad_species", str, None), ("gs_file", str, None), ("gs_species", str, None),
    ("out_folder", str, None), ("cov_file", str, None), ("ctrl_match_opt", str, "mean_var"),
    ("weight_opt", str, "vs"), ("adj_prop", str, None), ("flag_filter_data", bool, True),
    ("flag_raw_count", bool, True), ("n_ctrl", int, 1000), ("min_genes", int, 250), ("min_cells", int, 50),
    ("flag_return_ctrl_raw_score", bool, False), ("flag_return_ctrl_norm_score", bool, True),
]


_SIGNATURE_DOWNSTREAM = [
    ("h5ad_file", str, None), ("score_file", str, None), ("out_folder", str, None), ("group_analysis", str, None),
    ("corr_analysis", str, None), ("gene_analysis", str, None), ("flag_filter_data", bool, True),
    ("flag_raw_count", bool, True), ("min_genes", int, 250), ("min_cells", int, 50), ("knn_n_neighbors", int, 15),
    ("knn_n_pcs", int, 20),
]
_SIGNATURE_MUNGE = [
    ("out_file", str, None), ("pval_file", str, None), ("zscore_file", str, None), ("weight", str, "zscore"),
    ("fdr", float, None), ("fwer", float, None), ("n_min", int, 100), ("n_max", int, 1000),
]
_SIGNATURES = {"compute_score": _SIGNATURE, "perform_downstream": _SIGNATURE_DOWNSTREAM, "munge_gs": _SIGNATURE_MUNGE}


def get_cli_head():
    return (
        "******************************************************************************\n"
        "* Single-cell disease relevance score (scDRS) -- B200 scoring path\n"
        "* scdrs_b200 version %s\n"
        "******************************************************************************\n" % scdrs_b200.__version__
    )


def _to_bool(v):
    if isinstance(v, bool):
        return v
    if str(v) in ("True", "true", "1"):
        return True
    if str(v) in ("False", "false", "0"):
        return False
    raise ValueError("expected True/False, got %r" % (v,))


def parse_cli(argv):
    """argv (without the program name) -> (command, kwargs)."""
    if not argv:
        raise SystemExit("usage: scdrs compute-score --h5ad-file ... --gs-file ... --out-folder ...")
    command = argv[0].replace("-", "_")
    spec = {name: (typ, default) for name, typ, default in _SIGNATURES.get(command, _SIGNATURE)}
    kwargs = {name: default for name, (typ, default) in spec.items()}
    i = 1
    while i < len(argv):
        tok = argv[i]
        if not tok.startswith("--"):
            raise SystemExit("unexpected argument %r" % tok)
        if "=" in tok:
            key, val = tok[2:].split("=", 1)
            i += 1
        else:
            key = tok[2:]
            if i + 1 < len(argv) and not argv[i + 1].startswith("--"):
                val = argv[i + 1]
                i += 2
            else:
                val = "True"
                i += 1
        key = key.replace("-", "_")
        if key not in spec:
            raise SystemExit("unknown option --%s" % key.replace("_", "-"))
        typ = spec[key][0]
        kwargs[key] = _to_bool(val) if typ is bool else (None if val == "None" else typ(val))
    return command, kwargs


def compute_score(h5ad_file, h5ad_species, gs_file, gs_species, out_folder, cov_file=None,
                  ctrl_match_opt="mean_var", weight_opt="vs", adj_prop=None, flag_filter_data=True,
                  flag_raw_count=True, n_ctrl=1000, min_genes=250, min_cells=50,
                  flag_return_ctrl_raw_score=False, flag_return_ctrl_norm_score=True):
    """Score every trait of a ``.gs`` file and write ``<trait>.score.gz`` (first six columns) and, when
    control scores are kept, ``<trait>.full_score.gz`` (gzip TSV with the cell index)."""
    sys_start_time = time.time()
    if h5ad_species != gs_species:
        h5ad_species = util.convert_species_name(h5ad_species)
        gs_species = util.convert_species_name(gs_species)

    header = get_cli_head()
    header += "Call: scdrs compute-score \\\n"
    for key, val in [("h5ad-file", h5ad_file), ("h5ad-species", h5ad_species), ("cov-file", cov_file),
                     ("gs-file", gs_file), ("gs-species", gs_species), ("ctrl-match-opt", ctrl_match_opt),
                     ("weight-opt", weight_opt), ("adj-prop", adj_prop), ("flag-filter-data", flag_filter_data),
                     ("flag-raw-count", flag_raw_count), ("n-ctrl", n_ctrl), ("min-genes", min_genes),
                     ("min-cells", min_cells), ("flag-return-ctrl-raw-score", flag_return_ctrl_raw_score),
                     ("flag-return-ctrl-norm-score", flag_return_ctrl_norm_score)]:
        header += "--%s %s \\\n" % (key, val)
    header += "--out-folder %s\n" % out_folder
    print(header)

    if h5ad_species != gs_species:
        if h5ad_species not in ["mmusculus", "hsapiens"]:
            raise ValueError("--h5ad-species needs to be one of [mmusculus, hsapiens] "
                             "unless --h5ad-species==--gs-species")
        if gs_species not in ["mmusculus", "hsapiens"]:
            raise ValueError("--gs-species needs to be one of [mmusculus, hsapiens] "
                             "unless --h5ad-species==--gs-species")
    if ctrl_match_opt not in ["mean", "mean_var"]:
        raise ValueError("--ctrl-match-opt needs to be one of [mean, mean_var]")
    if weight_opt not in ["uniform", "vs", "inv_std", "od"]:
        raise ValueError("--weight-opt needs to be one of [uniform, vs, inv_std, od]")

    print("Loading data:")
    adata = util.load_h5ad(h5ad_file, flag_filter_data=flag_filter_data, flag_raw_count=flag_raw_count,
                           min_genes=min_genes, min_cells=min_cells)
    print("--h5ad-file loaded: n_cell=%d, n_gene=%d (sys_time=%0.1fs)"
          % (adata.shape[0], adata.shape[1], time.time() - sys_start_time))
    print("First 3 cells: %s" % (str(list(adata.obs_names[:3]))))
    print("First 5 genes: %s" % (str(list(adata.var_names[:5]))))
    if adj_prop is not None:
        assert adj_prop in adata.obs, "'adj_prop'=%s not in 'adata.obs.columns'" % adj_prop
        counts = adata.obs[adj_prop].value_counts()
        print("--adj-prop: %d categories" % counts.shape[0])
        print("Top 3: %s" % ", ".join("%s (%d)" % (x, counts[x]) for x in counts.index[:3]))
        print("Bottom 3: %s" % ", ".join("%s (%d)" % (x, counts[x]) for x in counts.index[-3:]))

    if cov_file is not None:
        df_cov = pd.read_csv(cov_file, sep="\t", index_col=0)
        df_cov.index = [str(x) for x in df_cov.index]
        print("--cov-file loaded: covariates=%s (sys_time=%0.1fs)"
              % (str(list(df_cov.columns)), time.time() - sys_start_time))
        print("n_cell=%d (%d in .h5ad)" % (df_cov.shape[0], len(set(df_cov.index) & set(adata.obs_names))))
        print("First 3 cells: %s" % (str(list(df_cov.index[:3]))))
        for col in df_cov.columns:
            print("First 5 values for '%s': %s" % (col, str(list(df_cov[col].values[:5]))))
    else:
        df_cov = None

    dict_gs = util.load_gs(gs_file, src_species=gs_species, dst_species=h5ad_species,
                           to_intersect=adata.var_names)
    print("--gs-file loaded: n_trait=%d (sys_time=%0.1fs)" % (len(dict_gs), time.time() - sys_start_time))
    print("Print info for first 3 traits:")
    for gs in list(dict_gs)[:3]:
        print("First 3 elements for '%s': %s, %s" % (gs, str(dict_gs[gs][0][:3]), str(dict_gs[gs][1][:3])))
    print("")

    t_loaded = time.time()
    print("Preprocessing:")
    scdrs_b200.preprocess(adata, cov=df_cov, adj_prop=adj_prop, n_mean_bin=20, n_var_bin=20, copy=False)
    print("")
    t_preprocessed = time.time()

    print("Computing scDRS score:")
    os.makedirs(out_folder, exist_ok=True)
    for trait in dict_gs:
        if len(dict_gs[trait][0]) < 10:
            print("trait=%s: skipped due to small size (n_gene=%d, sys_time=%0.1fs)"
                  % (trait, len(dict_gs[trait][0]), time.time() - sys_start_time))

    def write(trait, df_res):
        # same bytes (after gunzip) as DataFrame.to_csv(sep="\t", compression="gzip"), written by the
        # multi-threaded formatter/deflater of libscdrs_b200 (bin/scdrs:276-288 of the reference)
        util.write_score_file(os.path.join(out_folder, "%s.score.gz" % trait), df_res.iloc[:, 0:6])
        if "ctrl_norm_text" in df_res.attrs:
            # the control columns were formatted on the GPU (csrc/score_format.cu): only copied and deflated here
            from scdrs_b200 import _native

            text = df_res.attrs["ctrl_norm_text"]
            _native.write_score_gz_slots(os.path.join(out_folder, "%s.full_score.gz" % trait), df_res,
                                         text.columns, text.array)
        elif flag_return_ctrl_raw_score | flag_return_ctrl_norm_score:
            util.write_score_file(os.path.join(out_folder, "%s.full_score.gz" % trait), df_res)
        v_fdr = util.fdr_bh(df_res["pval"].values)
        print("Trait=%s, n_gene=%d: %d/%d FDR<0.1 cells, %d/%d FDR<0.2 cells (sys_time=%0.1fs)"
              % (trait, len(dict_gs[trait][0]), (v_fdr < 0.1).sum(), df_res.shape[0], (v_fdr < 0.2).sum(),
                 df_res.shape[0], time.time() - sys_start_time))

    t_write = [0.0]
    write_inner = write

    def write(trait, df_res):               # noqa: F811 -- timed wrapper around the writer above
        t0 = time.time()
        write_inner(trait, df_res)
        t_write[0] += time.time() - t0

    scored = scdrs_b200.score_traits(adata, dict_gs, ctrl_match_key=ctrl_match_opt, n_ctrl=n_ctrl,
                                     weight_opt=weight_opt, return_ctrl_raw_score=flag_return_ctrl_raw_score,
                                     return_ctrl_norm_score=flag_return_ctrl_norm_score, min_genes=10,
                                     on_result=write,
                                     text_output=(flag_return_ctrl_norm_score and not flag_return_ctrl_raw_score
                                                  and os.environ.get("SCDRS_B200_TEXT_OUTPUT", "1") != "0"))
    t_end = time.time()
    n_scored = max(1, len(scored))
    # scoring and file writing overlap (the writer runs beside the GPU): "score+write" is their wall clock
    print("Timing: load=%.2fs preprocess=%.2fs score+write=%.2fs (%.3fs per trait over %d traits; "
          "%.3fs per trait inside the score-file writer) total=%.2fs"
          % (t_loaded - sys_start_time, t_preprocessed - t_loaded, t_end - t_preprocessed,
             (t_end - t_preprocessed) / n_scored, len(scored), t_write[0] / n_scored, t_end - sys_start_time))


def perform_downstream(h5ad_file, score_file, out_folder, group_analysis=None, corr_analysis=None,
                       gene_analysis=None, flag_filter_data=True, flag_raw_count=True, min_genes=250, min_cells=50,
                       knn_n_neighbors=15, knn_n_pcs=20):
    """Group / cell-variable / gene analyses of precomputed ``.full_score.gz`` files (bin/scdrs:558-816):
    ``<trait>.scdrs_group.<annot>``, ``<trait>.scdrs_cell_corr``, ``<trait>.scdrs_gene`` in ``out_folder``.

    The group analysis needs the cells' neighbour graph.  The reference builds it with scanpy
    (``sc.pp.pca`` + ``sc.pp.neighbors``, bin/scdrs:748-760) when the file has none; scanpy is not part of this
    build, so the ``.h5ad`` must carry ``obsp["connectivities"]`` (any file that went through
    ``sc.pp.neighbors``)."""
    sys_start_time = time.time()
    header = get_cli_head()
    header += "Call: scdrs perform-downstream \\\n"
    assert sum([group_analysis is not None, corr_analysis is not None, gene_analysis is not None]) > 0, \
        "Expect at least one of `--group_analysis`, `--corr_analysis`, `--gene_analyis`"
    if group_analysis is not None:
        group_analysis = group_analysis.split(",") if isinstance(group_analysis, str) else list(group_analysis)
    if corr_analysis is not None:
        corr_analysis = corr_analysis.split(",") if isinstance(corr_analysis, str) else list(corr_analysis)
    for key, val in [("h5ad-file", h5ad_file), ("score-file", score_file), ("out-folder", out_folder),
                     ("group-analysis", None if group_analysis is None else ",".join(group_analysis)),
                     ("corr-analysis", None if corr_analysis is None else ",".join(corr_analysis)),
                     ("gene-analysis", gene_analysis), ("flag-filter-data", flag_filter_data),
                     ("flag-raw-count", flag_raw_count), ("min-genes", min_genes), ("min-cells", min_cells),
                     ("knn-n-neighbors", knn_n_neighbors)]:
        if val is not None:
            header += "--%s %s \\\n" % (key, val)
    header += "--knn-n-pcs %s\n" % knn_n_pcs
    print(header)

    adata = util.load_h5ad(h5ad_file=h5ad_file, flag_filter_data=flag_filter_data, flag_raw_count=flag_raw_count,
                           min_genes=min_genes, min_cells=min_cells)
    print("--h5ad-file loaded: n_cell=%d, n_gene=%d (sys_time=%0.1fs)"
          % (adata.shape[0], adata.shape[1], time.time() - sys_start_time))
    dict_score = util.load_scdrs_score(score_file)
    print("--score-file loaded: n_trait=%d (sys_time=%0.1fs)" % (len(dict_score), time.time() - sys_start_time))
    assert len(dict_score) > 0, "Expect at least 1 score file."
    if group_analysis is not None:
        assert len(set(group_analysis) - set(adata.obs)) == 0, \
            "Missing --group-analysis variables from `adata.obs.columns`."
        if "connectivities" not in adata.obsp:
            raise ValueError("`connectivities` not found in the .h5ad's obsp; run `sc.pp.neighbors(adata, n_neighbors=%d, "
                             "n_pcs=%d)` and save the file first (scanpy's graph construction is not part of scdrs_b200)"
                             % (knn_n_neighbors, knn_n_pcs))
    if corr_analysis is not None:
        assert len(set(corr_analysis) - set(adata.obs)) == 0, \
            "Missing --corr-analysis variables from `adata.obs.columns`."
    os.makedirs(out_folder, exist_ok=True)
    method = scdrs_b200.method

    if group_analysis is not None:
        print("Performing scDRS group-analysis")
        for trait in dict_score:
            dict_df_res = method.downstream_group_analysis(adata=adata, df_full_score=dict_score[trait],
                                                           group_cols=group_analysis)
            for group_col in group_analysis:
                name = "%s.scdrs_group.%s" % (trait, group_col.replace(" ", "_").replace(",", "_"))
                dict_df_res[group_col].to_csv(os.path.join(out_folder, name), sep="\t", index=True)
            print("Finish group-analysis for %s (sys_time=%0.1fs)" % (trait, time.time() - sys_start_time))
        print("")
    if corr_analysis is not None:
        print("Performing scDRS corr-analysis")
        for trait in dict_score:
            df_res = method.downstream_corr_analysis(adata=adata, df_full_score=dict_score[trait], var_cols=corr_analysis)
            df_res.to_csv(os.path.join(out_folder, "%s.scdrs_cell_corr" % trait), sep="\t", index=True)
            print("Finish corr-analysis for %s (sys_time=%0.1fs)" % (trait, time.time() - sys_start_time))
        print("")
    if gene_analysis is not None:
        print("Performing scDRS gene-analysis")
        for trait in dict_score:
            df_res = method.downstream_gene_analysis(adata=adata, df_full_score=dict_score[trait])
            df_res.to_csv(os.path.join(out_folder, "%s.scdrs_gene" % trait), sep="\t", index=True)
            print("Finish gene-analysis for %s (sys_time=%0.1fs)" % (trait, time.time() - sys_start_time))
        print("")


def _read_gene_table(path, what):
    """A GENE x trait table of floats; rows without a gene name dropped, duplicate names refused."""
    df = pd.read_csv(path, sep="\t", index_col=0)
    assert df.index.name == "GENE", "Header of the first column should be `GENE`."
    for col in df.columns:
        if df[col].dtype != float:
            raise ValueError(f"Column {col} contains non-float entries. Please check the input file.")
    print("--%s loaded: n_gene=%d, n_trait=%d" % (what, df.shape[0], df.shape[1]))
    if df.index.isnull().any():
        print("Removing %d rows with GENE=NaN" % df.index.isnull().sum())
        df = df.loc[~df.index.isnull()]
    if not df.index.is_unique:
        raise ValueError("Gene names in the provided %s are not unique: " % what,
                         ",".join(df.index[df.index.duplicated()]),
                         ". Please make sure the gene names are unique beforehand.")
    return df


def munge_gs(out_file, pval_file=None, zscore_file=None, weight="zscore", fdr=None, fwer=None, n_min=100, n_max=1000):
    """Gene-level p-values (or one-sided z-scores) per trait -> an scDRS ``.gs`` file (bin/scdrs:307-555): per
    trait the genes with the smallest p-values -- as many as pass the FDR (Benjamini-Hochberg) or FWER
    (Bonferroni) threshold, kept between ``n_min`` and ``n_max``, or the top ``n_max`` without a threshold --
    weighted by their z-score capped at 10 (``weight="zscore"``) or by 1 (``"uniform"``)."""
    import numpy as np

    assert (pval_file is not None) + (zscore_file is not None) == 1, "One of --pval-file and --zscore-file is expected."
    assert weight in ["zscore", "uniform"], "--weight needs to be one of 'zscore', 'uniform'"
    assert (fdr is None) or (fwer is None), "At most one of --fdr and --fwer is allowed."
    assert out_file is not None, "--out-file is expected."
    n_min = min(n_min, n_max)
    print(get_cli_head() + "Call: scdrs munge-gs --weight %s --n-min %s --n-max %s --out-file %s\n"
          % (weight, n_min, n_max, out_file))
    if pval_file is not None:
        df_pval = _read_gene_table(pval_file, "pval-file")
        if not (df_pval.min(skipna=True).min() > 0 and df_pval.max(skipna=True).max() < 1):
            print("Warning: pval-file are not all between 0 and 1.")
    else:
        df_zscore = _read_gene_table(zscore_file, "zscore-file")
        if not (df_zscore.min(skipna=True).min() < 0 and df_zscore.max(skipna=True).max() > 1):
            print("Warning: zscore-file values are all between 0 and 1.")
        df_pval = df_zscore.apply(lambda col: util.zsc2pval(col.values))

    rows = []
    for trait in sorted(df_pval.columns):
        p = df_pval[trait].dropna()
        if fdr is not None:
            n_gene = int((util.fdr_bh(p.values) <= fdr).sum())         # BH rejections at level fdr
        elif fwer is not None:
            n_gene = int((p.values <= fwer / max(len(p), 1)).sum())    # Bonferroni rejections at level fwer
        else:
            n_gene = n_max
        n_gene = max(min(n_gene, n_max), n_min)
        p = p.sort_values(ascending=True).iloc[:n_gene]
        pv = p.values.clip(min=1e-100)
        w = util.pval2zsc(pv).clip(max=10) if weight == "zscore" else np.ones(len(p))
        rows.append((trait, ",".join(f"{g}:{x:.5g}" for g, x in zip(p.index, w))))
    df_gs = pd.DataFrame(rows, columns=["TRAIT", "GENESET"])
    df_gs.to_csv(out_file, sep="\t", index=False)
    print("Finish munging %d gene sets" % df_gs.shape[0])


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    command, kwargs = parse_cli(argv)
    if command == "perform_downstream":
        missing = [k for k in ("h5ad_file", "score_file", "out_folder") if kwargs[k] is None]
        if missing:
            raise SystemExit("missing required option(s): " + ", ".join("--" + m.replace("_", "-") for m in missing))
        return perform_downstream(**kwargs)
    if command == "munge_gs":
        return munge_gs(**kwargs)
    if command != "compute_score":
        raise SystemExit("scdrs_b200 implements `compute-score`, `perform-downstream` and `munge-gs` (got %r)" % argv[0])
    missing = [k for k in ("h5ad_file", "h5ad_species", "gs_file", "gs_species", "out_folder") if kwargs[k] is None]
    if missing:
        raise SystemExit("missing required option(s): " + ", ".join("--" + m.replace("_", "-") for m in missing))
    compute_score(**kwargs)


if __name__ == "__main__":
    main()
