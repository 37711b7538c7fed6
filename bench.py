#!/usr/bin/env python
"""Benchmark of the scDRS scoring hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--n-cell C] [--n-gene G] [--n-trait T] [--n-ctrl R]

Metric (BASELINE.json): cell x trait scores per second at n_ctrl = 1000.
A step = scoring all T traits of the workload once (every trait is one pass of
spmm_score -> background correction -> pooled empirical null over all cells).

  value : device-resident throughput -- X, covariates and gene statistics already in HBM, control
          gene sets already drawn; timed with CUDA events around K steps.
  e2e   : the same through the public API (``scdrs_b200.score_traits`` = ``score_cell`` per trait)
          with HOST inputs: every step uploads X / covariates, draws the control sets on the host,
          uploads them, and copies every result back into float32 DataFrames.
  roofline     : the dominant kernel (spmm_score_kernel) against the measured HBM peak.
  cpu_baseline : the oracle port of the reference (oracle/scdrs_oracle.py) on a bounded sample.

Default workload at N = 1: BASELINE.json configs[1] -- 110,096 cells x 20,000 genes, 10 % nnz, 74
binary 1000-gene traits, 4 covariates (implicit covariate correction), n_ctrl = 1000.
With N > 1 (torchrun) every rank holds its own 110,096-cell shard (weak scaling); the pooled null
spans all ranks' cells (all-reduce of column statistics and of the rank histogram).
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-cell", type=int, default=110096)
    ap.add_argument("--n-gene", type=int, default=20000)
    ap.add_argument("--n-trait", type=int, default=74)
    ap.add_argument("--n-trait-gene", type=int, default=1000)
    ap.add_argument("--n-ctrl", type=int, default=1000)
    ap.add_argument("--density", type=float, default=0.10)
    ap.add_argument("--weighted", action="store_true", help="MAGMA-z-like gene weights (BASELINE configs[2])")
    ap.add_argument("--spmm", default="auto", choices=["auto", "f64", "fixed"],
                    help="arithmetic of the raw-score scatter (auto: exact fixed point when it applies)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-cells", type=int, default=30000)
    ap.add_argument("--cpu-sample-traits", type=int, default=2)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic workload
# ------------------------------------------------------------------------------------------------
def make_csr_on_device(n_cell, n_gene, density, seed, device):
    """TMS-FACS-shaped CSR built on the GPU with torch (plumbing only): same recipe as
    scdrs_b200.synth.make_csr -- Bernoulli(q_g) expression, values log1p(Gamma(1, 3))."""
    import torch

    from scdrs_b200 import synth

    rng = np.random.default_rng(seed)
    q = torch.from_numpy(synth.gene_detection_prob(n_gene, density, rng).astype(np.float32)).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    counts, cols = [], []
    chunk = 8192
    for start in range(0, n_cell, chunk):
        m = min(chunk, n_cell - start)
        mask = torch.rand((m, n_gene), device=device, generator=gen) < q
        counts.append(mask.sum(dim=1))
        cols.append(mask.nonzero()[:, 1].to(torch.int32))
        del mask
    indices = torch.cat(cols)
    indptr = torch.zeros(n_cell + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.cat(counts), 0)
    data = torch.empty(indices.shape[0], dtype=torch.float32, device=device)
    data.exponential_(1.0 / 3.0, generator=gen)          # Gamma(1, scale 3)
    data = torch.log1p(data)
    return indptr, indices, data


def pin_csr(adata):
    """Move the host copy of X into page-locked memory (the end-to-end number uploads its inputs from pinned
    host memory, as the bench contract states); the CSR matrix becomes a view of three pinned torch tensors."""
    import numpy as np
    import torch
    from scipy import sparse

    X = adata.X
    keep = [torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).pin_memory()
            for a, dt in ((X.data, np.float32), (X.indices, np.int32), (X.indptr, np.int64))]
    Xp = sparse.csr_matrix((keep[0].numpy(), keep[1].numpy(), keep[2].numpy()), shape=X.shape, copy=False)
    adata.X = Xp
    adata._pinned_host_buffers = keep


def build_workload(args, rank, device):
    """AnnData with HOST copies of everything + SCDRS_PARAM from the product's own preprocess."""
    import pandas as pd
    import torch
    from scipy import sparse

    import scdrs_b200
    from scdrs_b200 import synth
    from scdrs_b200.anndata_lite import AnnData

    t0 = time.time()
    indptr, indices, data = make_csr_on_device(args.n_cell, args.n_gene, args.density, 20261019 + rank, device)
    torch.cuda.synchronize()
    h_indptr = indptr.cpu().numpy()
    h_indices = indices.cpu().numpy()
    h_data = data.cpu().numpy()
    del indptr, indices, data
    torch.cuda.empty_cache()
    X = sparse.csr_matrix((h_data, h_indices, h_indptr), shape=(args.n_cell, args.n_gene))
    obs = pd.DataFrame(index=["cell_%d_%d" % (rank, i) for i in range(args.n_cell)])
    var = pd.DataFrame(index=["gene_%05d" % i for i in range(args.n_gene)])
    adata = AnnData(X, obs, var)
    cov = synth.make_covariates(X, seed=7 + rank)
    cov.index = obs.index
    t_gen = time.time() - t0
    t0 = time.time()
    scdrs_b200.preprocess(adata, cov=cov)
    t_pp = time.time() - t0
    if os.environ.get("SCDRS_BENCH_PAGEABLE", "0") != "1":
        pin_csr(adata)
    traits = synth.make_traits(adata.var_names, n_trait=args.n_trait, n_gene_per_trait=args.n_trait_gene,
                               weighted=bool(getattr(args, "weighted", False)), seed=1000)
    return adata, traits, dict(generate_s=round(t_gen, 2), preprocess_s=round(t_pp, 2), nnz=int(X.nnz))


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
# CPU arm: the oracle port of the reference on a bounded sample
# ------------------------------------------------------------------------------------------------
def _cpu_worker(payload):
    """Score `traits` on the sample with the oracle (one process)."""
    import warnings

    warnings.filterwarnings("ignore")
    from oracle import scdrs_oracle as orc

    adata, traits, n_ctrl = payload
    t0 = time.time()
    for genes, w in traits:
        orc.score_cell(adata, genes, gene_weight=w, n_ctrl=n_ctrl)
    return time.time() - t0


def cpu_sample(args, n_proc):
    """A bounded sample of the workload: the first `cpu_sample_cells` cells, same genes, n_ctrl."""
    import warnings

    warnings.filterwarnings("ignore")
    from oracle import scdrs_oracle as orc
    from scdrs_b200 import synth
    from scdrs_b200.loess1d import loess

    n_cell = min(args.cpu_sample_cells, args.n_cell)
    adata, cov = synth.make_adata(n_cell, args.n_gene, density=args.density, seed=20261019)
    # SCDRS_PARAM for the sample: gene statistics from one sparse pass (the dense oracle preprocess
    # would need n_cell x n_gene float64); the scoring arm under test is score_cell, not preprocess
    _cheap_param(adata, cov, loess)
    traits = synth.make_traits(adata.var_names, n_trait=max(n_proc, args.cpu_sample_traits) if n_proc > 1 else args.cpu_sample_traits,
                               n_gene_per_trait=args.n_trait_gene, weighted=False, seed=1000)
    return adata, list(traits.values()), n_cell


def _cheap_param(adata, cov, loess):
    """SCDRS_PARAM via sparse moments (host numpy; bench-only helper for the CPU arm)."""
    import pandas as pd

    X = adata.X.tocsr().astype(np.float64)
    n = X.shape[0]
    C = cov.values.astype(np.float64)
    gmean = np.asarray(X.mean(axis=0)).reshape(-1)
    beta = np.linalg.solve(C.T @ C / n, np.asarray((X.T @ C).T) / n)
    B = -beta                                     # k x n_gene (COV_BETA.T)
    # moments of v = x + C B + mu via dense-part closed forms is overkill here: use X's own moments
    mean = gmean
    var = np.asarray(X.power(2).mean(axis=0)).reshape(-1) - mean ** 2
    ct = X.copy(); ct.data = np.expm1(ct.data)
    ct_mean = np.asarray(ct.mean(axis=0)).reshape(-1)
    ct_var = np.asarray(ct.power(2).mean(axis=0)).reshape(-1) - ct_mean ** 2
    ok = (ct_var > 0) & (ct_mean > 0)
    est = np.zeros(X.shape[1])
    m = loess(np.log10(ct_mean[ok]), np.log10(ct_var[ok]), span=0.3, degree=2); m.fit()
    est[ok] = m.outputs.fitted_values
    with np.errstate(all="ignore"):
        var_tech = var * 10 ** est / ct_var
    var_tech[np.isnan(var_tech)] = 0
    df = pd.DataFrame({"mean": mean, "var": var, "var_tech": var_tech}, index=adata.var_names)
    mb = pd.qcut(df["mean"], 20, labels=False, duplicates="drop")
    lab = np.full(X.shape[1], "", dtype=object)
    for b in sorted(set(mb)):
        sel = (mb == b).values
        vb = pd.qcut(df.loc[sel, "var"], 20, labels=False, duplicates="drop")
        lab[sel] = ["%d.%d" % (b, x) for x in vb]
    df["mean_var"] = lab
    adata.uns["SCDRS_PARAM"] = {
        "FLAG_SPARSE": True, "FLAG_COV": True, "FLAG_ADJ_PROP": False, "GENE_STATS": df,
        "COV_MAT": cov, "COV_BETA": pd.DataFrame(B.T, index=adata.var_names, columns=cov.columns),
        "COV_GENE_MEAN": pd.Series(gmean, index=adata.var_names),
    }


def run_cpu_arm(args, n_proc):
    """Time the oracle on the sample with n_proc processes (traits spread over processes)."""
    import multiprocessing as mp

    adata, traits, n_cell = cpu_sample(args, n_proc)
    if n_proc <= 1:
        dt = _cpu_worker((adata, traits, args.n_ctrl))
        n_tr = len(traits)
    else:
        per = [traits[i::n_proc] for i in range(n_proc)]
        per = [p for p in per if p]
        t0 = time.time()
        with mp.get_context("fork").Pool(len(per)) as pool:
            pool.map(_cpu_worker, [(adata, p, args.n_ctrl) for p in per])
        dt = time.time() - t0
        n_tr = len(traits)
    return n_cell * n_tr / dt, dt, n_cell, n_tr


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_proc = max(1, min(os.cpu_count() or 1, 32))
    vals = []
    for _ in range(max(0, min(args.warmup, 1))):
        run_cpu_arm(args, n_proc)
    t_all = 0.0
    for _ in range(max(1, args.steps)):
        v, dt, n_cell, n_tr = run_cpu_arm(args, n_proc)
        vals.append(v)
        t_all += dt
    value = float(np.mean(vals))
    sample = "%d cells x %d genes, %d traits of %d genes, n_ctrl=%d, oracle port of score_cell, %d processes over traits" % (
        n_cell, args.n_gene, n_tr, args.n_trait_gene, args.n_ctrl, n_proc)
    line = {
        "impl": "reference", "metric": "cell x trait scores/sec at n_ctrl=%d" % args.n_ctrl, "value": value,
        "unit": "cell*trait/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_all / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "cell*trait/s", "cores": n_proc, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "cell*trait/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def baseline_config_name(args, n_gpus):
    """Which of BASELINE.json's configs the synthetic shape corresponds to."""
    total = args.n_cell * n_gpus
    if args.n_gene == 30000 and args.n_ctrl == 5000 and total >= 3500000:
        return "BASELINE configs[4]"
    if args.n_gene == 20000 and args.n_ctrl == 1000 and total >= 1000000 and args.n_cell != 110096:
        return "BASELINE configs[3]"
    if (args.n_cell, args.n_gene, args.n_ctrl) == (110096, 20000, 1000):
        return "BASELINE configs[2]" if getattr(args, "weighted", False) else "BASELINE configs[1]"
    return "custom shape"


def workload_config(args, n_gpus):
    return {
        "workload": "TMS-FACS-shaped synthetic (%s): %d cells x %d genes per GPU, %.0f%% nnz, "
                    "%d %s %d-gene traits, 4 covariates (implicit covariate correction), n_ctrl=%d, "
                    "weight_opt=vs, ctrl_match_key=mean_var" % (baseline_config_name(args, n_gpus), args.n_cell, args.n_gene, 100 * args.density,
                                                               args.n_trait, "weighted" if getattr(args, "weighted", False) else "binary",
                                                               args.n_trait_gene, args.n_ctrl),
        "n_cell_per_gpu": args.n_cell, "n_gene": args.n_gene, "n_trait": args.n_trait, "n_ctrl": args.n_ctrl,
        "sharding": "cells across %d GPU(s); pooled null over all cells" % n_gpus,
        "l2": "inputs larger than L2 (X is %.2f GB per GPU, re-read for every trait)" % (
            8e-9 * args.n_cell * args.n_gene * args.density),
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


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import scdrs_b200
    from scdrs_b200 import _device, _native
    from scdrs_b200 import method as M

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

    adata, traits, info = build_workload(args, rank, device)
    if world > 1:
        # One atlas, cells sharded: the gene-level parameters (GENE_STATS incl. the mean_var bins,
        # COV_BETA, COV_GENE_MEAN) must be the same on every rank or the ranks would draw different
        # control sets.  The bench takes rank 0's (a distributed preprocess is outside this bench).
        keys = ["GENE_STATS", "COV_BETA", "COV_GENE_MEAN"]
        box = [{k: adata.uns["SCDRS_PARAM"][k] for k in keys} if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        adata.uns["SCDRS_PARAM"].update(box[0])
    names = list(traits)
    n_cell, n_ctrl = args.n_cell, args.n_ctrl
    state = _device.get_state(adata)
    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, True)
    # run the kernels on torch's current stream so that torch CUDA events bracket them
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)

    # ---- control sets drawn once (host), as inputs of the device-resident leg
    from concurrent.futures import ThreadPoolExecutor

    tab = M._gene_table(adata, state)
    bins = M._GeneBins(tab.df_gene, "mean_var", 200)
    with ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1))) as pool:
        preps = list(pool.map(lambda t: M._prepare_trait(adata, state, traits[t][0], traits[t][1], "mean_var",
                                                         n_ctrl, 200, 0, tab, bins), names))
    ctx.set_gene_stats(preps[0]["var"], preps[0]["var_tech"])

    if world > 1:
        from scdrs_b200.distributed import CudaPhases, ShardedScorer

        scorer = ShardedScorer(CudaPhases(ctx, device), device)
        def run_trait(p):
            return scorer.score(p["trait_cols"], p["trait_gw"], p["ctrl_cols"], p["ctrl_gw"], "vs", True)
    else:
        def run_trait(p):
            return ctx.score_trait(p["trait_cols"], p["trait_gw"], p["ctrl_cols"], p["ctrl_gw"], "vs", True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    spmm_ms = []

    def step(record):
        for p in preps:
            run_trait(p)
            if record:
                spmm_ms.append(ctx.last_timing()[0])

    for _ in range(args.warmup):
        step(False)
    launches0 = ctx.launch_count()
    barrier()
    with ClockSampler(local_rank) as clocks:
        t0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step(True)
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
    total_cells = n_cell * world
    value = total_cells * len(names) * args.steps / wall

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    nnz = info["nnz"]
    k = 4
    alg_bytes = 8 * nnz + 8 * (n_cell + 1) + 4 * k * n_cell + 4 * n_cell * (1 + n_ctrl)
    k1_ms = float(np.mean(spmm_ms)) if spmm_ms else float("nan")
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    traffic, binding = None, None
    tpath = os.path.join(ROOT, "profiles", "r1_spmm_traffic.json")
    if os.path.exists(tpath) and (args.n_cell, args.n_gene, args.n_ctrl) == (110096, 20000, 1000):
        prof = json.load(open(tpath))
        traffic = float(prof["dram_bytes_total"])   # from the committed ncu capture
        if "lsu_data_pipe_pct_of_peak" in prof:
            # the resource that actually binds the scatter: the SM's LSU data pipe takes one wavefront per
            # clock; floor = wavefronts of the ncu capture / (SMs x clock), against the live launch time
            wf = float(prof["lsu_data_pipe_pct_of_peak"]) / 100.0 * float(prof.get("capture_ms", 9.227)) * 1e-3 * 1.965e9 * 148
            floor_ms = wf / (148 * 1.965e9) * 1e3
            binding = {"resource": "LSU data pipe (shared-memory + L1 wavefronts, one per SM per clock)",
                       "wavefronts_per_launch": wf, "floor_ms": floor_ms,
                       "frac_of_floor": None, "source": "profiles/r1_spmm_fx_ncu_summary.md"}
    spmm_kind, spmm_shift = ctx.last_spmm_kind()
    roofline = {"bound": "hbm", "kernel": "spmm_score_fx_kernel" if spmm_kind == "fixed" else "spmm_score_kernel",
                "arithmetic": ("exact integer sums of round(x*base*2^%d), 32-bit shared-memory words, 64-bit totals" % spmm_shift)
                if spmm_kind == "fixed" else "fp64 accumulators in shared memory",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k1_ms, "binding_resource": binding,
                "scatter_fma_per_launch": float(nnz) * (1 + n_ctrl) * args.n_trait_gene / args.n_gene,
                "note": ("bound by the LSU / shared-memory data pipe (86% busy in ncu: one conflict-free 32-bit load + "
                         "store per accumulate, 1.1e10 accumulates per launch), not by HBM; see DESIGN.md section 4")
                if spmm_kind == "fixed" else
                ("bound by the shared-memory/L1 data pipe (89% busy in ncu: one 64-bit load+store per accumulate, "
                 "1.1e10 accumulates per launch), not by HBM; see DESIGN.md section 4")}

    if binding is not None and spmm_kind == "fixed" and not getattr(args, "weighted", False):
        binding["frac_of_floor"] = binding["floor_ms"] / k1_ms
    else:
        roofline["binding_resource"] = None

    # ---- per-phase device times of one trait (CUDA events between the C-ABI phases), single GPU only
    phases = None
    if world == 1:
        ncol = 1 + n_ctrl
        t_cs = torch.empty(ncol, dtype=torch.float64, device=device)
        t_c2, t_cm = torch.empty_like(t_cs), torch.empty_like(t_cs)
        t_q = torch.empty(n_cell, dtype=torch.float32, device=device)
        t_h = torch.empty(n_cell + 1, dtype=torch.int64, device=device)
        acc = np.zeros(6)
        reps = preps[: min(3, len(preps))]
        for p in reps:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            ctx.trait_raw(p["trait_cols"], p["trait_gw"], p["ctrl_cols"], p["ctrl_gw"], "vs", True, t_cs.data_ptr()); ev[1].record()
            ctx.trait_rowstats(t_cs.data_ptr(), n_cell, t_c2.data_ptr(), t_cm.data_ptr()); ev[2].record()
            ctx.trait_queries(t_c2.data_ptr(), t_cm.data_ptr(), t_q.data_ptr()); ev[3].record()
            ctx.trait_count(t_q.data_ptr(), n_cell, t_h.data_ptr()); ev[4].record()
            ctx.trait_finish(t_h.data_ptr(), n_cell, 0, n_cell * n_ctrl, n_ctrl); ev[5].record()
            torch.cuda.synchronize()
            acc[:5] += [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
            acc[5] += ctx.last_timing()[0]
        acc /= len(reps)
        mat_bytes = 4.0 * n_cell * (1 + n_ctrl)
        phases = {
            "raw_ms": acc[0], "scatter_ms": acc[5], "raw_other_ms": acc[0] - acc[5], "rowstats_ms": acc[1],
            "queries_ms": acc[2], "count_ms": acc[3], "finish_ms": acc[4],
            "note": "raw_other = set upload + gene-major lists + matching + column sums; count = sort of the queries + "
                    "grid + one pass over the null; finish = suffix sums + result download",
            "rowstats_GBps_of_one_read": mat_bytes / (acc[1] * 1e-3) / 1e9,
            "count_GBps_of_one_read": mat_bytes / (acc[3] * 1e-3) / 1e9,
        }

    # ---- end to end through the public API with host inputs
    e2e = None
    if not args.no_e2e:
        scdrs_b200.clear_device_cache()

        def e2e_step():
            # every step uploads X and the covariate terms again (from pinned host memory) and downloads its
            # results; the context keeps its device / pinned working buffers between steps
            scdrs_b200.invalidate_uploads()
            st_ = _device._STATES.get(id(adata))
            if st_ is not None:
                st_.h2d_bytes = 0          # counted per step
            res = scdrs_b200.score_traits(adata, traits, n_ctrl=n_ctrl, sharded=world > 1)
            return res
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 2))
        for _ in range(n_e2e):
            res = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t_e2e = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
            dt = float(t_e2e.item())
        st = _device.get_state(adata)
        h2d = st.h2d_bytes + len(names) * (preps[0]["ctrl_cols"].nbytes + 3 * 8 * args.n_trait_gene + 2 * 8 * args.n_gene)
        d2h = len(names) * n_cell * (8 + 4 + 4 + 8)
        e2e = {"value": total_cells * len(names) / dt, "unit": "cell*trait/s", "h2d_bytes_per_step": int(h2d) * world,
               "d2h_bytes_per_step": int(d2h) * world, "s_per_step": dt,
               "api": "scdrs_b200.score_traits (score_cell per trait, host control-gene draws overlapped; X uploaded "
                      "from pinned host memory every step, working buffers kept between steps)"
                      + ("; one process per GPU, sharded=True" if world > 1 else "")}
        del res

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        v, dt, sc_cells, sc_traits = run_cpu_arm(args, 1)
        cpu = {"value": v, "unit": "cell*trait/s", "cores": 1, "kind": "port",
               "sample": "%d cells x %d genes, %d traits, n_ctrl=%d, oracle port of score_cell, %.1f s" % (
                   sc_cells, args.n_gene, sc_traits, n_ctrl, dt)}

    if rank == 0:
        line = {
            "metric": "cell x trait scores/sec at n_ctrl=%d" % n_ctrl, "value": value, "unit": "cell*trait/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s32 fixed point (exact sums) + f64" if spmm_kind == "fixed" else "f64",
            "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu, "phases": phases, "setup": info,
            "device_ms_per_trait": 1e3 * wall / args.steps / len(names), "wall_s_check": wall_check,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
