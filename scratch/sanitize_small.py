"""Small end-to-end pass over every kernel of the scoring path, for compute-sanitizer memcheck."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from scipy import sparse
import scdrs_b200
from scdrs_b200 import synth, util, _device
from scdrs_b200.anndata_lite import AnnData
adata, cov = synth.make_adata(300, 700, density=0.12, seed=3)
# count preprocessing on raw-count-like data, then the resident matrix goes through preprocess + scoring
X = adata.X.copy(); X.data = np.floor(np.expm1(X.data) * 3 + 1).astype(np.float32)
ad = util.preprocess_counts(AnnData(X, adata.obs.copy(), adata.var.copy()), True, True, min_genes=40, min_cells=10)
cov2 = cov.loc[ad.obs_names]
scdrs_b200.preprocess(ad, cov=cov2)
for weighted in (False, True):
    genes, w = synth.make_traits(ad.var_names, 1, 90, weighted=weighted, seed=5)["trait_00"]
    for n_ctrl in (30, 1100):
        for mode in ("auto", "f64"):
            _device.set_spmm_mode(mode)
            df = scdrs_b200.score_cell(ad, genes, gene_weight=w, n_ctrl=n_ctrl, return_ctrl_norm_score=True, return_ctrl_raw_score=True)
            assert np.isfinite(df["norm_score"].values).all()
_device.set_spmm_mode("auto")
res = scdrs_b200.score_traits(ad, synth.make_traits(ad.var_names, 3, 60, seed=8), n_ctrl=50)
print("ok", ad.shape, len(res))
