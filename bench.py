#!/usr/bin/env python
"""Benchmark of the scDRS scoring hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--config 2|3|4|5] [--n-cell C] [--n-gene G] [--n-trait T] [--n-ctrl R]
                    [--shard cells|traits] [--weak]

Metric (BASELINE.json): cell x trait scores per second at n_ctrl = 1000.
A step = scoring all T traits of the workload once (every trait is one pass of
spmm_score -> background correction -> pooled empirical null over ALL cells of the atlas).

Default workload: BASELINE.json configs[3] -- ONE synthetic atlas of 1,000,000 cells x 20,000 genes
(10 % stored entries), 74 binary 1000-gene traits, 4 covariates (implicit covariate correction),
n_ctrl = 1000.  The atlas is the same for every N (it is generated in 5,000-cell chunks seeded by the
global chunk number): with --gpus N its cells are SPLIT over the N ranks (strong scaling, the experiment
BASELINE.json's north_star names), the gene statistics come from a distributed preprocess over all
cells and the pooled null spans all ranks.  --config 2 is configs[1] (110,096 cells, one GPU),
--config 3 adds MAGMA-z-like gene weights and the adj_prop cell weights, --weak gives every rank its own
--n-cell cells (round 1's behaviour), --shard traits replicates X and deals the traits over the ranks.

  value : device-resident throughput -- X, covariates, gene statistics and the drawn control sets
          already in HBM; timed with CUDA events around K steps, max over ranks.
  e2e   : the same through the public API (``scdrs_b200.score_traits`` = ``score_cell`` per trait)
          with HOST inputs: every step uploads X / covariates from pinned host memory, draws the control
          sets on the host, uploads them, and copies every result back into float32 DataFrames.
  roofline     : the dominant kernel (the raw-score scatter) against the measured HBM peak.
  cpu_baseline : the reference's score_cell (oracle/_ref through oracle/ref_shim.py when staged, else
                 the numpy port oracle/scdrs_oracle.py) on a bounded sample of the same cells, 1 core.
  parity       : the CUDA path against that CPU run on the same sample (same SCDRS_PARAM).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    2: dict(n_cell=110096, n_gene=20000, n_ctrl=1000, weighted=False, adj_prop=False, name="BASELINE configs[1]"),
    3: dict(n_cell=110096, n_gene=20000, n_ctrl=1000, weighted=True, adj_prop=True, name="BASELINE configs[2]"),
    4: dict(n_cell=1000000, n_gene=20000, n_ctrl=1000, weighted=False, adj_prop=False, name="BASELINE configs[3]"),
    5: dict(n_cell=4000000, n_gene=30000, n_ctrl=5000, weighted=False, adj_prop=False, name="BASELINE configs[4]"),
}
CHUNK = 5000          # cells per generation chunk: the unit cells are dealt to ranks in
DATA_SEED = 20261019


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=sorted(CONFIGS))
    ap.add_argument("--n-cell", type=int, default=None, help="cells of the whole atlas (per rank with --weak)")
    ap.add_argument("--n-gene", type=int, default=None)
    ap.add_argument("--n-trait", type=int, default=74)
    ap.add_argument("--n-trait-gene", type=int, default=1000)
    ap.add_argument("--n-ctrl", type=int, default=None)
    ap.add_argument("--density", type=float, default=0.10)
    ap.add_argument("--weighted", action="store_true", help="MAGMA-z-like gene weights (BASELINE configs[2])")
    ap.add_argument("--adj-prop", action="store_true", help="cell-type proportion adjustment, 120 Zipf groups (configs[2])")
    ap.add_argument("--weak", action="store_true", help="every rank holds its own --n-cell cells (weak scaling)")
    ap.add_argument("--shard", default="cells", choices=["cells", "traits"],
                    help="cells: rows of X split over the ranks; traits: X replicated, traits dealt round-robin")
    ap.add_argument("--spmm", default="auto", choices=["auto", "f64", "fixed"],
                    help="arithmetic of the raw-score scatter (auto: exact fixed point when it applies)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-cells", type=int, default=8000)
    ap.add_argument("--cpu-sample-traits", type=int, default=2)
    args = ap.parse_args(argv)
    cfg = CONFIGS[args.config]
    for key in ("n_cell", "n_gene", "n_ctrl"):
        if getattr(args, key) is None:
            setattr(args, key, cfg[key])
    args.weighted = args.weighted or cfg["weighted"]
    args.adj_prop = args.adj_prop or cfg["adj_prop"]
    return args


def shard_range(n_cell, rank, world):
    """Cells [lo, hi) of rank `rank`: whole generation chunks, dealt as evenly as the chunking allows."""
    n_chunk = (n_cell + CHUNK - 1) // CHUNK
    c0, c1 = rank * n_chunk // world, (rank + 1) * n_chunk // world
    return min(c0 * CHUNK, n_cell), min(c1 * CHUNK, n_cell)


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def make_csr_on_device(n_cell_total, n_gene, density, seed, device, lo=0, hi=None):
    """Rows [lo, hi) of the TMS-FACS-shaped atlas, built on the GPU with torch (plumbing only): same recipe as
    scdrs_b200.synth.make_csr -- Bernoulli(q_g) expression, values log1p(Gamma(1, 3)).  Every CHUNK of cells
    has its own generator seed, so the atlas does not depend on how its cells are split over ranks."""
    import torch

    from scdrs_b200 import synth

    hi = n_cell_total if hi is None else hi
    assert lo % CHUNK == 0
    rng = np.random.default_rng(seed)
    q = torch.from_numpy(synth.gene_detection_prob(n_gene, density, rng).astype(np.float32)).to(device)
    gen = torch.Generator(device=device)
    counts, cols, vals = [], [], []
    for start in range(lo, hi, CHUNK):
        m = min(CHUNK, hi - start)
        gen.manual_seed(seed * 1000003 + start // CHUNK)
        mask = torch.rand((m, n_gene), device=device, generator=gen) < q
        counts.append(mask.sum(dim=1))
        c = mask.nonzero()[:, 1].to(torch.int32)
        del mask
        v = torch.empty(c.shape[0], dtype=torch.float32, device=device)
        v.exponential_(1.0 / 3.0, generator=gen)          # Gamma(1, scale 3)
        cols.append(c)
        vals.append(torch.log1p(v))
    indices = torch.cat(cols)
    data = torch.cat(vals)
    indptr = torch.zeros(hi - lo + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.cat(counts), 0)
    return indptr, indices, data


def pin_csr(adata):
    """Move the host copy of X into page-locked memory (the end-to-end number uploads its inputs from pinned
    host memory, as the bench contract states); the CSR matrix becomes a view of three pinned torch tensors."""
    import torch
    from scipy import sparse

    X = adata.X
    keep = [torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).pin_memory()
            for a, dt in ((X.data, np.float32), (X.indices, np.int32), (X.indptr, np.int64))]
    Xp = sparse.csr_matrix((keep[0].numpy(), keep[1].numpy(), keep[2].numpy()), shape=X.shape, copy=False)
    adata.X = Xp
    adata._pinned_host_buffers = keep


def make_cov_frame(n_cell_total, n_genes_local, lo, hi, index):
    """const, n_genes, sex_male, age for cells [lo, hi) (one seeded draw for the whole atlas, sliced)."""
    import pandas as pd

    rng = np.random.default_rng(7)
    sex = rng.integers(0, 2, size=n_cell_total).astype(np.float64)[lo:hi]
    age = rng.choice([3.0, 18.0, 24.0], size=n_cell_total)[lo:hi]
    return pd.DataFrame({"const": np.ones(hi - lo), "n_genes": n_genes_local.astype(np.float64),
                         "sex_male": sex, "age": age}, index=index)


def make_groups(n_cell_total, lo, hi, n_group=120):
    """120 cell types with Zipf(1.2)-like sizes floored at ~1 % of the largest."""
    rng = np.random.default_rng(11)
    size = 1.0 / np.arange(1, n_group + 1) ** 1.2
    size = np.maximum(size, 0.012 * size.max())
    lab = rng.choice(n_group, size=n_cell_total, p=size / size.sum())[lo:hi]
    return np.array(["ct_%03d" % i for i in range(n_group)], dtype=object)[lab]


def build_workload(args, rank, device, world=1, sharded_stats=False):
    """AnnData with HOST copies of everything + SCDRS_PARAM from the product's own preprocess.

    Strong scaling (default): the rank's rows of the one atlas; --weak or --shard traits: see the module doc."""
    import pandas as pd
    import torch
    from scipy import sparse

    import scdrs_b200
    from scdrs_b200 import synth
    from scdrs_b200.anndata_lite import AnnData

    weak = bool(getattr(args, "weak", False))
    by_traits = getattr(args, "shard", "cells") == "traits"
    if weak:
        n_total = args.n_cell * world
        lo, hi = rank * args.n_cell, (rank + 1) * args.n_cell
        assert args.n_cell % CHUNK == 0 or world == 1, "--weak shards must be whole chunks of %d cells" % CHUNK
    elif by_traits or world == 1:
        n_total, lo, hi = args.n_cell, 0, args.n_cell
    else:
        n_total = args.n_cell
        lo, hi = shard_range(n_total, rank, world)
    t0 = time.time()
    indptr, indices, data = make_csr_on_device(n_total, args.n_gene, args.density, DATA_SEED, device, lo, hi)
    torch.cuda.synchronize()
    h_indptr = indptr.cpu().numpy()
    h_indices = indices.cpu().numpy()
    h_data = data.cpu().numpy()
    del indptr, indices, data
    torch.cuda.empty_cache()
    n_local = hi - lo
    X = sparse.csr_matrix((h_data, h_indices, h_indptr), shape=(n_local, args.n_gene))
    obs = pd.DataFrame(index=["cell_%d" % i for i in range(lo, hi)])
    if getattr(args, "adj_prop", False):
        obs["cell_type"] = make_groups(n_total, lo, hi)
    var = pd.DataFrame(index=["gene_%05d" % i for i in range(args.n_gene)])
    adata = AnnData(X, obs, var)
    cov = make_cov_frame(n_total, np.diff(h_indptr), lo, hi, obs.index)
    t_gen = time.time() - t0
    pinned = os.environ.get("SCDRS_BENCH_PAGEABLE", "0") != "1"
    if pinned:
        pin_csr(adata)            # preprocess uploads X too: from page-locked memory, like every e2e step
    t0 = time.time()
    pp_kw = {}
    if getattr(args, "adj_prop", False):
        pp_kw["adj_prop"] = "cell_type"
    if sharded_stats:
        pp_kw["sharded"] = True
    scdrs_b200.preprocess(adata, cov=cov, **pp_kw)
    t_pp = time.time() - t0
    if pinned and not (getattr(adata, "_pinned_host_buffers", None) is not None
                       and np.shares_memory(adata.X.data, adata._pinned_host_buffers[0].numpy())):
        pin_csr(adata)            # preprocess replaced adata.X
    traits = synth.make_traits(adata.var_names, n_trait=args.n_trait, n_gene_per_trait=args.n_trait_gene,
                               weighted=bool(getattr(args, "weighted", False)), seed=1000)
    info = dict(generate_s=round(t_gen, 2), preprocess_s=round(t_pp, 2), nnz=int(X.nnz), n_cell_local=n_local,
                n_cell_total=n_total, cell_range=[lo, hi])
    timing = getattr(scdrs_b200.pp, "LAST_TIMING", None)
    if timing:
        info["preprocess_phases_s"] = {k: round(v, 4) for k, v in timing.items()}
    return adata, traits, info


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's score_cell (unmodified sources through the shim when staged under oracle/_ref,
# else the numpy port) on a bounded sample
# ------------------------------------------------------------------------------------------------
def cpu_scorer():
    """(kind, score_cell callable, preprocess callable) of the CPU arm."""
    import warnings

    warnings.filterwarnings("ignore")
    from oracle import ref_shim
    from oracle import scdrs_oracle as orc
    from scdrs_b200.loess1d import loess

    if ref_shim.available() and os.environ.get("SCDRS_BENCH_CPU_KIND", "") != "port":
        ref_pp, ref_method = ref_shim.load(loess_impl=loess)
        return "reference", ref_method.score_cell, lambda a, **kw: ref_pp.preprocess(a, **kw)
    return "port", orc.score_cell, lambda a, **kw: orc.preprocess(a, loess_impl=loess, **kw)


_CPU_STATE = {}


def _cpu_worker(job):
    """Score the given traits on the sample held by this process (forked after the sample was built)."""
    import warnings

    warnings.filterwarnings("ignore")
    adata, score_cell, n_ctrl = _CPU_STATE["adata"], _CPU_STATE["score_cell"], _CPU_STATE["n_ctrl"]
    t0 = time.time()
    for genes, w in job:
        score_cell(adata, genes, gene_weight=w, n_ctrl=n_ctrl)
    return time.time() - t0


def reference_sample(args, n_trait, kind_wanted, n_cell):
    """A bounded sample of the workload for the CPU arm: `n_cell` cells of the same recipe, same genes, same
    n_ctrl; SCDRS_PARAM from the CPU arm's own preprocess (nothing of the product on this path)."""
    from scdrs_b200 import synth

    n_cell = min(n_cell, args.n_cell)
    adata, cov = synth.make_adata(n_cell, args.n_gene, density=args.density, seed=DATA_SEED,
                                  n_group=120 if args.adj_prop else 0)
    os.environ["SCDRS_BENCH_CPU_KIND"] = "" if kind_wanted == "reference" else "port"
    kind, score_cell, preprocess = cpu_scorer()
    kw = {"adj_prop": "cell_type"} if args.adj_prop else {}
    t0 = time.time()
    preprocess(adata, cov=cov, **kw)
    t_pp = time.time() - t0
    traits = synth.make_traits(adata.var_names, n_trait=n_trait, n_gene_per_trait=args.n_trait_gene,
                               weighted=bool(args.weighted), seed=1000)
    return adata, list(traits.values()), n_cell, kind, score_cell, t_pp


def _describe(kind, n_cell, args, n_trait, n_proc, t_pp):
    return ("%d cells x %d genes, %d traits of %d genes per step, n_ctrl=%d, %s, %d processes (one trait each); "
            "sample built and preprocessed once (%.1f s, untimed)" % (
                n_cell, args.n_gene, n_trait, args.n_trait_gene, args.n_ctrl,
                "unmodified reference score_cell (oracle/_ref via ref_shim)" if kind == "reference"
                else "numpy port of score_cell (oracle/scdrs_oracle.py)", n_proc, t_pp))


def reference_arm(args):
    """--impl reference: the reference's CPU score_cell on the box's host cores, one process per core over traits
    (the reference's own SLURM-array idiom, experiments/job.compute_score/*.sh).  The sample is built ONCE; a step
    scores one trait per process.

    The unmodified reference spends ~20 s per trait on fixed Python overhead (1001 pandas label look-ups per
    gene set, scdrs/method.py:172-183) whatever the number of cells, so K = 20 steps of it do not fit "a few
    minutes".  One step of it is always run and reported (``reference_check``); the K timed steps use it when they
    fit the budget (SCDRS_BENCH_REF_BUDGET_S, default 560 s of a driver slot of ~800 s) and otherwise the numpy port, which is 3-7x FASTER
    than the reference per core -- the driver's ratio against it is conservative."""
    import multiprocessing as mp

    from oracle import ref_shim

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("TQDM_DISABLE", "1")
    t_arm = time.time()
    budget = float(os.environ.get("SCDRS_BENCH_REF_BUDGET_S", "560"))
    n_proc = max(1, min(os.cpu_count() or 1, 32))
    steps = max(1, args.steps)

    def run_steps(jobs, n_steps):
        out = []
        with mp.get_context("fork").Pool(n_proc) as pool:
            for _ in range(n_steps):
                t0 = time.time()
                pool.map(_cpu_worker, jobs)
                out.append(time.time() - t0)
        return out

    check = None
    times = None
    if ref_shim.available() and os.environ.get("SCDRS_BENCH_CPU_KIND", "") != "port":
        adata, traits, n_cell, kind, score_cell, t_pp = reference_sample(args, n_proc, "reference", args.cpu_sample_cells)
        _CPU_STATE.update(adata=adata, score_cell=score_cell, n_ctrl=args.n_ctrl)
        jobs = [[t] for t in traits]
        t1 = run_steps(jobs, 1)[0]                         # doubles as the (single) warm-up step
        check = {"value": n_cell * len(traits) / t1, "unit": "cell*trait/s", "cores": n_proc, "kind": "reference",
                 "s_per_step": t1, "sample": _describe(kind, n_cell, args, len(traits), n_proc, t_pp)}
        if (time.time() - t_arm) + steps * t1 * 1.1 <= budget:
            times = run_steps(jobs, steps)
    if times is None:
        adata, traits, n_cell, kind, score_cell, t_pp = reference_sample(args, n_proc, "port", 2 * args.cpu_sample_cells)
        _CPU_STATE.update(adata=adata, score_cell=score_cell, n_ctrl=args.n_ctrl)
        jobs = [[t] for t in traits]
        run_steps(jobs, max(0, min(args.warmup, 1)))       # a CPU loop has nothing to warm beyond the page cache
        times = run_steps(jobs, steps)
    dt = float(np.mean(times))
    value = n_cell * len(traits) / dt
    sample = _describe(kind, n_cell, args, len(traits), n_proc, t_pp)
    cfg = workload_config(args, 1)
    cfg["reference_sample"] = sample
    cfg["workload"] += " -- CPU arm timed on a %d-cell sample of it, %d traits per step" % (n_cell, len(traits))
    line = {
        "impl": "reference", "metric": "cell x trait scores/sec at n_ctrl=%d" % args.n_ctrl, "value": value,
        "unit": "cell*trait/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "strong" if not args.weak else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "cell*trait/s", "cores": n_proc, "kind": kind, "sample": sample},
        "reference_check": check,
        "e2e": {"value": value, "unit": "cell*trait/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "arm_wall_s": round(time.time() - t_arm, 1),
    }
    emit(line)


def workload_config(args, n_gpus):
    weak = bool(getattr(args, "weak", False))
    by_traits = getattr(args, "shard", "cells") == "traits"
    total = args.n_cell * (n_gpus if weak else 1)
    name = "custom shape"
    for cfg in CONFIGS.values():
        if (total, args.n_gene, args.n_ctrl, bool(args.weighted), bool(args.adj_prop)) == (
                cfg["n_cell"], cfg["n_gene"], cfg["n_ctrl"], cfg["weighted"], cfg["adj_prop"]):
            name = cfg["name"]
    per_gpu = total if by_traits else (args.n_cell if weak else -(-total // n_gpus))
    if by_traits:
        sharding = "X replicated on %d GPU(s), traits dealt round-robin, no collectives" % n_gpus
    else:
        sharding = "cells of ONE atlas split across %d GPU(s) (%s); gene statistics and pooled null over all cells" % (
            n_gpus, "weak: every rank adds its own cells" if weak else "strong: total cells fixed")
    return {
        "workload": "TMS-FACS-shaped synthetic (%s): %d cells x %d genes, %.0f%% nnz, %d %s %d-gene traits, "
                    "4 covariates (implicit covariate correction)%s, n_ctrl=%d, weight_opt=vs, ctrl_match_key=mean_var" % (
                        name, total, args.n_gene, 100 * args.density, args.n_trait,
                        "weighted" if args.weighted else "binary", args.n_trait_gene,
                        ", adj_prop over 120 groups" if args.adj_prop else "", args.n_ctrl),
        "n_cell_total": total, "n_cell_per_gpu": per_gpu, "n_gene": args.n_gene, "n_trait": args.n_trait,
        "n_ctrl": args.n_ctrl, "sharding": sharding,
        "l2": "inputs larger than L2 (X is %.2f GB per GPU, re-read for every trait)" % (
            8e-9 * per_gpu * args.n_gene * args.density),
    }


# ------------------------------------------------------------------------------------------------
_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner, warnings of
    worker processes) is sent to stderr, and the line itself goes to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parity_block(df, ref, n_cell, n_ctrl):
    """What north_star asks to be reported: float differences, and the cells whose pooled rank / Monte-Carlo
    count differs from the float64 CPU run (ties straddling the float32 rounding of a score)."""
    raw, raw_ref = df["raw_score"].values.astype(np.float64), ref["raw_score"].values.astype(np.float64)
    nrm, nrm_ref = df["norm_score"].values.astype(np.float64), ref["norm_score"].values.astype(np.float64)
    n_null = n_cell * n_ctrl
    r_got = np.round(df["pval"].values.astype(np.float64) * (n_null + 1))
    r_ref = np.round(ref["pval"].values.astype(np.float64) * (n_null + 1))
    d_rank = np.abs(r_got - r_ref)
    d_mc = np.abs(df["mc_pval"].values.astype(np.float64) - ref["mc_pval"].values.astype(np.float64)) * (1 + n_ctrl)
    z_same = d_rank == 0
    return {
        "max_rel_raw": float(np.max(np.abs(raw - raw_ref) / np.maximum(np.abs(raw_ref), 1e-30))),
        "max_abs_norm": float(np.max(np.abs(nrm - nrm_ref))),
        "max_abs_zscore_where_rank_equal": float(np.max(np.abs(
            df["zscore"].values.astype(np.float64)[z_same] - ref["zscore"].values.astype(np.float64)[z_same]))) if z_same.any() else None,
        "n_rank_diff": int((d_rank > 0).sum()), "max_rank_diff": int(d_rank.max()),
        "n_mc_diff": int((d_mc > 0.5).sum()), "n_cell": int(n_cell), "n_null": int(n_null),
    }


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import scdrs_b200
    from scdrs_b200 import _device
    from scdrs_b200 import method as M
    from scdrs_b200.distributed import CudaPhases, ShardedScorer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    scdrs_b200.set_default_device(local_rank)
    scdrs_b200.set_spmm_mode(args.spmm)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    by_traits = args.shard == "traits"
    cell_sharded = world > 1 and not by_traits

    adata, traits, info = build_workload(args, rank, device, world, sharded_stats=cell_sharded)
    names = list(traits)
    n_local, n_total, n_ctrl = info["n_cell_local"], info["n_cell_total"], args.n_ctrl
    my_names = names[rank::world] if (by_traits and world > 1) else names
    state = _device.get_state(adata)
    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, True)

    # ---- control sets drawn once (host) and made resident: inputs of the device-resident leg
    from concurrent.futures import ThreadPoolExecutor

    tab = M._gene_table(adata, state)
    bins = M._GeneBins(tab.df_gene, "mean_var", 200)
    with ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) // max(1, world) - 1))) as pool:
        preps = list(pool.map(lambda t: M._prepare_trait(adata, state, traits[t][0], traits[t][1], "mean_var",
                                                         n_ctrl, 200, 0, tab, bins), my_names))
    ctx.set_gene_stats(preps[0]["var"], preps[0]["var_tech"])
    for p in preps:
        p["d_ctrl"] = torch.from_numpy(np.ascontiguousarray(p["ctrl_cols"], dtype=np.int32)).to(device)
    # kernels run on torch's current stream (CudaPhases binds it) so that torch CUDA events bracket them
    scorer = ShardedScorer(CudaPhases(ctx, device), device, group=None if cell_sharded else False)
    ctx.set_resident_sets(True)       # the control sets of this leg sit in HBM before the first step starts

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    spmm_ms = []

    def step(record):
        # two traits in flight: trait i + 1 is queued before the results of trait i are fetched
        pending = None
        for p in preps:
            ticket = scorer.submit(p["trait_cols"], p["trait_gw"], p["d_ctrl"], p["ctrl_gw"], "vs", True)
            if pending is not None:
                scorer.collect(pending)
            pending = ticket
            if record == "spmm":
                spmm_ms.append(ctx.last_timing()[0])        # synchronises: only in the separate roofline pass
        if pending is not None:
            scorer.collect(pending)

    for _ in range(args.warmup):
        step(None)
    launches0 = ctx.launch_count()
    barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step(None)
        ev1.record()
        barrier()
        wall = time.perf_counter() - t0
    launches = ctx.launch_count() - launches0
    # CUDA events on the launching stream (max over ranks); the wall clock is kept as a cross-check
    dev_s = ev0.elapsed_time(ev1) * 1e-3
    t_val = torch.tensor([dev_s, wall], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_val, op=dist.ReduceOp.MAX)
    wall_check = float(t_val[1].item())
    wall = float(t_val[0].item())
    work_cells = n_total                      # cells x traits scored by the whole job per step
    value = work_cells * len(names) * args.steps / wall

    ctx.set_resident_sets(False)
    # ---- roofline of the dominant kernel: its own CUDA events (ctx.ev[0..1] around the scatter), one more pass
    step("spmm")
    barrier()
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    nnz = info["nnz"]
    k = 4
    alg_bytes = 8 * nnz + 8 * (n_local + 1) + 4 * k * n_local + 4 * n_local * (1 + n_ctrl)
    k1_ms = float(np.mean(spmm_ms)) if spmm_ms else float("nan")
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    spmm_kind, spmm_shift = ctx.last_spmm_kind()
    traffic, tsrc = None, None
    for cand in ("r2_spmm_traffic.json", "r1_spmm_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", cand)
        if os.path.exists(tpath):
            prof = json.load(open(tpath))
            # ncu's dram bytes of one launch at config 2, scaled by the stored entries of this run's launch
            traffic = float(prof["dram_bytes_total"]) * (alg_bytes / float(prof.get("algorithmic_bytes", 2198722416)))
            tsrc = "profiles/%s (ncu --set full, config 2) scaled by algorithmic bytes" % cand
            break
    roofline = {"bound": "hbm", "kernel": "spmm_score_fx_kernel" if spmm_kind == "fixed" else "spmm_score_kernel",
                "arithmetic": ("exact integer sums of round(x*base*2^%d), 32-bit shared-memory words, 64-bit totals" % spmm_shift)
                if spmm_kind == "fixed" else "fp64 accumulators in shared memory",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k1_ms,
                "scatter_adds_per_launch": float(nnz) * (1 + n_ctrl) * args.n_trait_gene / args.n_gene,
                "note": "bound by the SM's LSU / shared-memory data pipe (one conflict-free 32-bit load + store per "
                        "accumulate, ~50 accumulates per stored entry), not by HBM; see DESIGN.md section 4"}

    # ---- per-phase device times of a trait: CUDA events between the phases and around every collective
    scorer.timing = {}
    for p in preps[: min(3, len(preps))]:
        scorer.collect(scorer.submit(p["trait_cols"], p["trait_gw"], p["d_ctrl"], p["ctrl_gw"], "vs", True))
    phases = scorer.timing_summary()
    scorer.timing = None
    if phases is not None:
        mat_bytes = 4.0 * n_local * (1 + n_ctrl)
        phases["scatter_ms"] = k1_ms
        phases["rowstats_GBps_of_one_read"] = mat_bytes / (phases["rowstats_ms"] * 1e-3) / 1e9
        phases["count_GBps_of_one_read"] = mat_bytes / (phases["count_ms"] * 1e-3) / 1e9
        phases["note"] = ("raw = list building + matching + scatter; count = sort of all queries + grid + one pass "
                          "over the local null; *_allreduce / gather = NCCL collectives timed by CUDA events on the "
                          "compute stream (max over ranks)")
        if world > 1:
            keys = sorted(k_ for k_ in phases if k_.endswith("_ms"))
            t = torch.tensor([phases[k_] for k_ in keys], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            phases.update({k_: float(v) for k_, v in zip(keys, t.tolist())})

    # ---- end to end through the public API with host inputs
    e2e = None
    if not args.no_e2e:
        my_traits = {t: traits[t] for t in my_names}
        for p in preps:
            p.pop("d_ctrl", None)
        torch.cuda.empty_cache()

        def e2e_step():
            # every step uploads X and the covariate terms again (from pinned host memory) and downloads its
            # results; the context keeps its device / pinned working buffers between steps
            scdrs_b200.invalidate_uploads()
            st_ = _device._STATES.get(id(adata))
            if st_ is not None:
                st_.h2d_bytes = 0          # counted per step
            return scdrs_b200.score_traits(adata, my_traits, n_ctrl=n_ctrl, sharded=cell_sharded)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            res = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        h2d = _device.get_state(adata).h2d_bytes + len(my_names) * (
            preps[0]["ctrl_cols"].nbytes + 3 * 8 * args.n_trait_gene + 2 * 8 * args.n_gene)
        d2h = len(my_names) * n_local * (8 + 4 + 4 + 8)
        t_e2e = torch.tensor([dt, float(h2d), float(d2h)], dtype=torch.float64, device=device)
        if world > 1:
            t_max = t_e2e.clone()
            dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
            dist.all_reduce(t_e2e, op=dist.ReduceOp.SUM)
            dt, h2d, d2h = float(t_max[0].item()), float(t_e2e[1].item()), float(t_e2e[2].item())
        e2e = {"value": work_cells * len(names) / dt, "unit": "cell*trait/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "s_per_step": dt,
               "api": "scdrs_b200.score_traits (score_cell per trait, host control-gene draws overlapped; X uploaded "
                      "from pinned host memory every step, working buffers kept between steps)"
                      + ("; one process per GPU, sharded=True" if cell_sharded else "")
                      + ("; one process per GPU, each scoring its share of the traits" if by_traits and world > 1 else "")}
        del res

    # ---- CPU baseline + parity on a bounded sample of the same cells (rank 0 at N = 1 only)
    cpu, parity = None, None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        cpu, parity = cpu_baseline_and_parity(args, adata, traits)

    if rank == 0:
        line = {
            "metric": "cell x trait scores/sec at n_ctrl=%d" % n_ctrl, "value": value, "unit": "cell*trait/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
            "scaling": "weak" if args.weak else "strong",
            "vs_baseline": None, "dtype": "s32 fixed point (exact sums) + f64" if spmm_kind == "fixed" else "f64",
            "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "phases": phases, "setup": info,
            "device_ms_per_trait": 1e3 * wall / args.steps / len(preps), "wall_s_check": wall_check,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline_and_parity(args, adata, traits):
    """The first `cpu_sample_cells` cells of the benchmarked atlas with the atlas's own gene-level parameters:
    the CPU score_cell is timed on them (1 core), the CUDA path scores the same inputs, and the two frames are
    compared."""
    import pandas as pd
    import warnings

    import scdrs_b200
    from scdrs_b200.anndata_lite import AnnData

    warnings.filterwarnings("ignore")
    kind, score_cell, _ = cpu_scorer()
    n_s = min(args.cpu_sample_cells, adata.shape[0])
    Xs = adata.X[:n_s].tocsr().copy()
    sub = AnnData(Xs, pd.DataFrame(index=adata.obs_names[:n_s]), pd.DataFrame(index=adata.var_names))
    param = dict(adata.uns["SCDRS_PARAM"])
    param["COV_MAT"] = param["COV_MAT"].iloc[:n_s]
    if "CELL_STATS" in param:
        param["CELL_STATS"] = param["CELL_STATS"].iloc[:n_s]
    sub.uns["SCDRS_PARAM"] = param
    names = list(traits)[: args.cpu_sample_traits]
    t0 = time.time()
    refs = [score_cell(sub, traits[t][0], gene_weight=traits[t][1], n_ctrl=args.n_ctrl) for t in names]
    dt = time.time() - t0
    cpu = {"value": n_s * len(names) / dt, "unit": "cell*trait/s", "cores": 1, "kind": kind,
           "sample": "first %d cells of the benchmarked atlas x %d genes (gene statistics of the whole atlas), %d traits, "
                     "n_ctrl=%d, %s, %.1f s" % (n_s, args.n_gene, len(names), args.n_ctrl,
                                                  "unmodified reference score_cell" if kind == "reference"
                                                  else "numpy port of score_cell", dt)}
    blocks = []
    for t, ref in zip(names, refs):
        df = scdrs_b200.score_cell(sub, traits[t][0], gene_weight=traits[t][1], n_ctrl=args.n_ctrl)
        blocks.append(parity_block(df, ref, n_s, args.n_ctrl))
    parity = {k_: (max(b[k_] for b in blocks) if k_.startswith("max") else blocks[0][k_] if k_ in ("n_cell", "n_null")
                   else sum(b[k_] for b in blocks)) for k_ in blocks[0]}
    parity["n_trait"] = len(blocks)
    parity["against"] = cpu["sample"]
    parity["tolerance"] = "north_star: raw / norm / z within 1e-5, ranks differ only where ties straddle it"
    return cpu, parity


if __name__ == "__main__":
    main()
