"""Wall clock of `scdrs compute-score` with default flags (the n_cell x 1006 `.full_score.gz` of every trait is
written) on a synthetic data set of BASELINE config 2's shape: the input `.h5ad` / `.cov` / `.gs` files are
generated here (GPU generator of bench.py, scdrs_b200.h5lite.write_h5ad), then the CLI runs as a subprocess.

    python tests/scripts/bench_cli.py [--n-cell 110096] [--n-gene 20000] [--n-trait 8] [--out gpurun_out/r2_cli.json]

Prints / stores: load, preprocess, score+write seconds, seconds per trait, bytes written."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-cell", type=int, default=110096)
    ap.add_argument("--n-gene", type=int, default=20000)
    ap.add_argument("--n-trait", type=int, default=8)
    ap.add_argument("--gzip-level", type=int, default=None)
    ap.add_argument("--out", default=None)
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()

    import numpy as np
    import pandas as pd
    import torch
    from scipy import sparse

    import bench
    from scdrs_b200 import h5lite, synth
    from scdrs_b200.anndata_lite import AnnData

    work = tempfile.mkdtemp(prefix="scdrs_cli_")
    t0 = time.time()
    dev = torch.device("cuda", 0)
    indptr, indices, data = bench.make_csr_on_device(a.n_cell, a.n_gene, 0.10, bench.DATA_SEED, dev)
    X = sparse.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(a.n_cell, a.n_gene))
    del indptr, indices, data
    torch.cuda.empty_cache()
    obs = pd.DataFrame(index=["cell_%d" % i for i in range(a.n_cell)])
    var = pd.DataFrame(index=["gene_%05d" % i for i in range(a.n_gene)])
    adata = AnnData(X, obs, var)
    cov = bench.make_cov_frame(a.n_cell, np.diff(X.indptr), 0, a.n_cell, obs.index)
    h5ad, covf, gsf, outd = (os.path.join(work, n) for n in ("data.h5ad", "data.cov", "traits.gs", "out"))
    h5lite.write_h5ad(h5ad, adata)
    cov.to_csv(covf, sep="\t")
    traits = synth.make_traits(adata.var_names, n_trait=a.n_trait, n_gene_per_trait=1000, seed=1000)
    with open(gsf, "w") as fh:
        fh.write("TRAIT\tGENESET\n")
        for t, (genes, _w) in traits.items():
            fh.write("%s\t%s\n" % (t, ",".join(genes)))
    t_gen = time.time() - t0
    env = dict(os.environ)
    if a.gzip_level is not None:
        env["SCDRS_B200_GZIP_LEVEL"] = str(a.gzip_level)
    cmd = [sys.executable, os.path.join(ROOT, "bin", "scdrs"), "compute-score", "--h5ad-file", h5ad, "--h5ad-species",
           "mouse", "--gs-file", gsf, "--gs-species", "mouse", "--cov-file", covf, "--flag-filter-data", "False",
           "--flag-raw-count", "False", "--out-folder", outd]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    wall = time.time() - t0
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
        raise SystemExit(r.returncode)
    m = re.search(r"Timing: load=([\d.]+)s preprocess=([\d.]+)s score\+write=([\d.]+)s \(([\d.]+)s per trait over (\d+) traits; "
                  r"([\d.]+)s per trait inside the score-file writer\) total=([\d.]+)s", r.stdout)
    files = [os.path.join(outd, f) for f in os.listdir(outd)]
    res = {
        "what": "scdrs compute-score, default flags (flag_return_ctrl_norm_score=True: n_cell x 1006 float32 of gzip TSV "
                "per trait), flag_filter_data / flag_raw_count False (input already log-normalised)",
        "n_cell": a.n_cell, "n_gene": a.n_gene, "n_trait": a.n_trait, "n_ctrl": 1000,
        "gzip_level": int(env.get("SCDRS_B200_GZIP_LEVEL", "1")), "host_cores": os.cpu_count(),
        "wall_s_subprocess": round(wall, 2), "input_generation_s": round(t_gen, 1),
        "h5ad_bytes": os.path.getsize(h5ad), "bytes_written": int(sum(os.path.getsize(f) for f in files)),
        "n_files": len(files),
    }
    if os.environ.get("SCDRS_B200_TRACE", "0") == "1":
        sys.stderr.write(r.stderr[-6000:])
    if m:
        res.update(load_s=float(m.group(1)), preprocess_s=float(m.group(2)), score_write_s=float(m.group(3)),
                   s_per_trait=float(m.group(4)), writer_s_per_trait=float(m.group(6)), cli_total_s=float(m.group(7)))
    # steady state: differences of the per-trait "sys_time" stamps of the CLI's own log (the first traits carry the
    # one-time allocations of pinned staging memory)
    stamps = [float(x) for x in re.findall(r"^Trait=.*sys_time=([\d.]+)s\)", r.stdout, flags=re.M)]
    if len(stamps) >= 4:
        d = sorted(b - a_ for a_, b in zip(stamps[:-1], stamps[1:]))
        res["s_per_trait_median_of_deltas"] = round(d[len(d) // 2], 3)
        res["s_per_trait_deltas"] = [round(b - a_, 3) for a_, b in zip(stamps[:-1], stamps[1:])]
    # the files are what pandas reads back: spot-check one
    one = pd.read_csv(os.path.join(outd, "trait_00.full_score.gz"), sep="\t", index_col=0, nrows=5)
    res["full_score_columns"] = int(one.shape[1])
    print(json.dumps(res))
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)
    if not a.keep:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
