"""``scdrs compute-score`` (drop-in for bin/scdrs:25-304 of the reference).

python-fire is not installed here, so the flag conventions it gives the reference's CLI are
restated: ``compute-score`` / ``compute_score``, ``--h5ad-file`` / ``--h5ad_file``,
``--flag True`` / ``--flag=False``.  Only the ``compute-score`` sub-command is on the hot path;
``munge-gs`` and ``perform-downstream`` are outside this repository's scope.
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
    spec = {name: (typ, default) for name, typ, default in _SIGNATURE}
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

    print("Preprocessing:")
    scdrs_b200.preprocess(adata, cov=df_cov, adj_prop=adj_prop, n_mean_bin=20, n_var_bin=20, copy=False)
    print("")

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
        if flag_return_ctrl_raw_score | flag_return_ctrl_norm_score:
            util.write_score_file(os.path.join(out_folder, "%s.full_score.gz" % trait), df_res)
        v_fdr = util.fdr_bh(df_res["pval"].values)
        print("Trait=%s, n_gene=%d: %d/%d FDR<0.1 cells, %d/%d FDR<0.2 cells (sys_time=%0.1fs)"
              % (trait, len(dict_gs[trait][0]), (v_fdr < 0.1).sum(), df_res.shape[0], (v_fdr < 0.2).sum(),
                 df_res.shape[0], time.time() - sys_start_time))

    scdrs_b200.score_traits(adata, dict_gs, ctrl_match_key=ctrl_match_opt, n_ctrl=n_ctrl, weight_opt=weight_opt,
                            return_ctrl_raw_score=flag_return_ctrl_raw_score,
                            return_ctrl_norm_score=flag_return_ctrl_norm_score, min_genes=10, on_result=write)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    command, kwargs = parse_cli(argv)
    if command != "compute_score":
        raise SystemExit("scdrs_b200 implements `compute-score` only (got %r); munge-gs and "
                         "perform-downstream are outside the accelerated path" % argv[0])
    missing = [k for k in ("h5ad_file", "h5ad_species", "gs_file", "gs_species", "out_folder") if kwargs[k] is None]
    if missing:
        raise SystemExit("missing required option(s): " + ", ".join("--" + m.replace("_", "-") for m in missing))
    compute_score(**kwargs)


if __name__ == "__main__":
    main()
