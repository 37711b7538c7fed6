"""Probe (historical): needed the twin-context experiment, which was measured and reverted; kept for the record."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scdrs_b200 import _device, method as M
class A: n_cell, n_gene, density, n_trait, n_trait_gene, weighted = 110096, 20000, 0.10, 2, 1000, False
dev = torch.device("cuda", 0)
adata, traits, info = bench.build_workload(A, 0, dev)
state = _device.get_state(adata); ctx = state.ensure_matrix(adata); state.ensure_cov(adata, True)
sA = torch.cuda.Stream(); sB = torch.cuda.Stream()
ctx.set_stream(sA.cuda_stream)
tab = M._gene_table(adata, state); bins = M._GeneBins(tab.df_gene, "mean_var", 200)
p = M._prepare_trait(adata, state, traits["trait_00"][0], traits["trait_00"][1], "mean_var", 1000, 200, 0, tab, bins)
ctx.set_gene_stats(p["var"], p["var_tech"])
colsum = torch.empty(1001, dtype=torch.float64, device=dev)
x = torch.rand(64 * 1024 * 1024, device=dev)      # 256 MB elementwise op: ~0.1 ms alone
for with_twin in (False, True):
    if with_twin:
        tw = state.twin_ctx()      # makes the scatter run with 18 warps per SM
    for rep in range(3):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); eA = torch.cuda.Event(enable_timing=True); eB = torch.cuda.Event(enable_timing=True); eB0 = torch.cuda.Event(enable_timing=True)
        e0.record(sA)
        ctx.trait_raw(p["trait_cols"], p["trait_gw"], p["ctrl_cols"], p["ctrl_gw"], "vs", True, colsum.data_ptr())
        eA.record(sA)
        time.sleep(0.003)            # the scatter is running now
        with torch.cuda.stream(sB):
            eB0.record(sB)
            y = x * 2.0
            eB.record(sB)
        torch.cuda.synchronize()
        print("twin=%s rep %d: scatter phase %.2f ms; small kernel launched at +%.2f ms, finished at +%.2f ms" % (
            with_twin, rep, e0.elapsed_time(eA), e0.elapsed_time(eB0), e0.elapsed_time(eB)))
