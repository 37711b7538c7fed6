"""load_h5ad's filters + normalisation at config-2 size: device path vs the host restatement (same box)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, pandas as pd, torch
from scipy import sparse
import bench, scdrs_b200
from oracle import h5ad_oracle
from scdrs_b200 import util, _device
from scdrs_b200.anndata_lite import AnnData
dev = torch.device("cuda", 0)
n_cell, n_gene = 110096, 20000
indptr, indices, data = bench.make_csr_on_device(n_cell, n_gene, 0.10, 20261019, dev)
counts = torch.floor(torch.expm1(data) * 3.0) + 1.0           # raw-count-like positive integers
X = sparse.csr_matrix((counts.cpu().numpy().astype(np.float32), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(n_cell, n_gene))
del indptr, indices, data, counts; torch.cuda.empty_cache()
obs = pd.DataFrame(index=["c%d" % i for i in range(n_cell)]); var = pd.DataFrame(index=["g%d" % i for i in range(n_gene)])
res = {"shape": [n_cell, n_gene], "nnz": int(X.nnz), "min_genes": 1900, "min_cells": 50}
for rep in range(2):
    t0 = time.perf_counter()
    ad = util.preprocess_counts(AnnData(X, obs, var), flag_filter_data=True, flag_raw_count=True, min_genes=1900, min_cells=50)
    t_dev = time.perf_counter() - t0
    st = _device.get_state(ad); ctx = st.ctx
    _device.clear_device_cache()
t0 = time.perf_counter()
Y, kc, kg, *_ = h5ad_oracle.preprocess_counts(X, flag_filter_data=True, flag_raw_count=True, min_genes=1900, min_cells=50)
t_host = time.perf_counter() - t0
assert ad.shape == Y.shape and np.array_equal(ad.X.indices, Y.indices) and np.allclose(ad.X.data, Y.data, rtol=2e-6)
res.update(out_shape=list(Y.shape), device_path_s=t_dev, host_restatement_s=t_host,
           note="device path = upload of the raw CSR + five kernels + download of the processed CSR into adata.X; "
                "the matrix stays resident for preprocess / score_cell")
print(json.dumps(res))
