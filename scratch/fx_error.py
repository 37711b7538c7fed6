"""Difference between the fixed-point and the fp64 scatter at config-2 size (scratch diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, scdrs_b200
from scdrs_b200 import _device
for weighted in (False, True):
    class A: n_cell, n_gene, density, n_trait, n_trait_gene = 110096, 20000, 0.10, 1, 1000
    A.weighted = weighted
    adata, traits, _ = bench.build_workload(A, 0, torch.device("cuda", 0))
    g, w = traits["trait_00"]
    out = {}
    for mode in ("fixed", "f64"):
        _device.set_spmm_mode(mode)
        out[mode] = scdrs_b200.score_cell(adata, g, gene_weight=w, n_ctrl=1000, return_ctrl_raw_score=True, return_ctrl_norm_score=True)
    _device.set_spmm_mode("auto")
    a, b = out["fixed"], out["f64"]
    cr = [c for c in a.columns if c.startswith("ctrl_raw")]; cn = [c for c in a.columns if c.startswith("ctrl_norm")]
    rel_raw = np.abs(a[cr].values - b[cr].values).max() / np.abs(b[cr].values).max()
    dn = np.abs(a[cn].values.astype(np.float64) - b[cn].values)
    dq = np.abs(a["norm_score"].values.astype(np.float64) - b["norm_score"].values)
    dp = (a["pval"].values != b["pval"].values).sum()
    print("weighted=%s: max|d raw|/max|raw| = %.2e, ctrl_norm: max %.2e rms %.2e, norm_score: max %.2e, cells with a different pooled p: %d of %d, mc_pval differs: %d" % (
        weighted, rel_raw, dn.max(), np.sqrt((dn ** 2).mean()), dq.max(), dp, len(a), (a["mc_pval"].values != b["mc_pval"].values).sum()))
    scdrs_b200.clear_device_cache()
