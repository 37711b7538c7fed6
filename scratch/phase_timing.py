"""Per-phase device timing of one trait (CUDA events on the context's stream) -- scratch diagnostics."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from scdrs_b200 import _device, method as M
import scdrs_b200

class A: n_cell, n_gene, density, n_trait, n_trait_gene, weighted = 110096, 20000, 0.10, 6, 1000, False
dev = torch.device("cuda", 0)
adata, traits, info = bench.build_workload(A, 0, dev)
state = _device.get_state(adata); ctx = state.ensure_matrix(adata); state.ensure_cov(adata, True)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
tab = M._gene_table(adata, state); bins = M._GeneBins(tab.df_gene, "mean_var", 200)
preps = [M._prepare_trait(adata, state, traits[t][0], traits[t][1], "mean_var", 1000, 200, 0, tab, bins) for t in traits]
ctx.set_gene_stats(preps[0]["var"], preps[0]["var_tech"])
n = A.n_cell; ncol = 1001
colsum = torch.empty(ncol, dtype=torch.float64, device=dev); col2 = torch.empty_like(colsum); colmin = torch.empty_like(colsum)
q = torch.empty(n, dtype=torch.float32, device=dev); hist = torch.empty(n + 1, dtype=torch.int64, device=dev)
def E(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it, p in enumerate(preps):
    torch.cuda.synchronize(); w0 = time.perf_counter()
    e = [E()]
    ctx.trait_raw(p["trait_cols"], p["trait_gw"], p["ctrl_cols"], p["ctrl_gw"], "vs", True, colsum.data_ptr()); e.append(E())
    ctx.trait_rowstats(colsum.data_ptr(), n, col2.data_ptr(), colmin.data_ptr()); e.append(E())
    ctx.trait_queries(col2.data_ptr(), colmin.data_ptr(), q.data_ptr()); e.append(E())
    ctx.trait_count(q.data_ptr(), n, hist.data_ptr()); e.append(E())
    out = ctx.trait_finish(hist.data_ptr(), n, 0, n * 1000, 1000); e.append(E())
    torch.cuda.synchronize(); w1 = time.perf_counter()
    k1 = ctx.last_timing()[0]
    ms = [e[i].elapsed_time(e[i + 1]) for i in range(5)]
    print("trait %d: raw %.2f (K1 %.2f, rest %.2f) rowstats %.2f queries %.2f count %.2f finish %.2f | device %.2f wall %.2f" % (
        it, ms[0], k1, ms[0] - k1, ms[1], ms[2], ms[3], ms[4], sum(ms), 1e3 * (w1 - w0)))
