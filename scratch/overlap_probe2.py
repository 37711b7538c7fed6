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
tw = state.twin_ctx(); tw.set_gene_stats(p["var"], p["var_tech"]); tw.set_stream(sB.cuda_stream)
n = A.n_cell
def bufs():
    return (torch.empty(1001, dtype=torch.float64, device=dev), torch.empty(1001, dtype=torch.float64, device=dev),
            torch.empty(1001, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float32, device=dev),
            torch.empty(n + 1, dtype=torch.int64, device=dev))
csA, _, _, _, _ = bufs(); csB, c2B, cmB, qB, hB = bufs()
args = (p["trait_cols"], p["trait_gw"], p["ctrl_cols"], p["ctrl_gw"], "vs", True)
x = torch.rand(16 * 1024 * 1024, device=dev)
for rep in range(3):
    tw.trait_raw(*args, csB.data_ptr())
    tw.trait_rowstats(csB.data_ptr(), n, c2B.data_ptr(), cmB.data_ptr())
    tw.trait_queries(c2B.data_ptr(), cmB.data_ptr(), qB.data_ptr())
    tw.trait_count(qB.data_ptr(), n, hB.data_ptr())
    torch.cuda.synchronize()
    names = ["a0", "a1", "b0", "torch", "queries", "count", "finish"]
    ev = {k: torch.cuda.Event(enable_timing=True) for k in names}
    ev["a0"].record(sA)
    ctx.trait_raw(*args, csA.data_ptr())
    ev["a1"].record(sA)
    time.sleep(0.003)
    with torch.cuda.stream(sB):
        ev["b0"].record(sB)
        y = x * 2.0
        ev["torch"].record(sB)
    tw.trait_queries(c2B.data_ptr(), cmB.data_ptr(), qB.data_ptr()); ev["queries"].record(sB)
    tw.trait_count(qB.data_ptr(), n, hB.data_ptr()); ev["count"].record(sB)
    tw.trait_finish(hB.data_ptr(), n, 0, n * 1000, 1000); ev["finish"].record(sB)
    torch.cuda.synchronize()
    print("rep %d: main raw ends %.2f | " % (rep, ev["a0"].elapsed_time(ev["a1"])) + "  ".join("%s %.2f" % (k, ev["a0"].elapsed_time(ev[k])) for k in names[2:]))
