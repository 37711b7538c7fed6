import sys; sys.path.insert(0,'.')
import numpy as np, warnings, time
warnings.filterwarnings('ignore')
from scdrs_b200 import synth
from oracle import scdrs_oracle as orc
from scdrs_b200.loess1d import loess
import bench
class A: pass
n_cell=400; n_gene=20000
adata, cov = synth.make_adata(n_cell, n_gene, density=0.1, seed=20261019)
bench._cheap_param(adata, cov, loess)
genes, w = synth.make_traits(adata.var_names, 1, 1000, seed=1000)["trait_00"]
df, info = orc.score_cell(adata, genes, gene_weight=w, n_ctrl=1000, return_internals=True)
X = adata.X.tocsr()
param = adata.uns['SCDRS_PARAM']
# weights per set
gs = param['GENE_STATS']
sets = np.vstack([info['trait_cols'][None,:], info['ctrl_cols']])
W64 = np.vstack([info['trait_w'][None,:], info['ctrl_w'].T])   # (1001, 1000)
# dense W (n_gene x 1001) float32
Wd = np.zeros((n_gene, 1001), dtype=np.float32)
for j in range(1001):
    Wd[sets[j], j] = W64[j].astype(np.float32)
C = param['COV_MAT'].values; B = param['COV_BETA'].values; mu = param['COV_GENE_MEAN'].values
raw64 = np.zeros((n_cell,1001)); raw32 = np.zeros((n_cell,1001))
for c in range(n_cell):
    idx = X.indices[X.indptr[c]:X.indptr[c+1]]; x = X.data[X.indptr[c]:X.indptr[c+1]]
    P = x[:,None].astype(np.float64) * Wd[idx].astype(np.float64)     # exact products
    raw64[c] = P.sum(axis=0)
    # sequential fp32 accumulation of exact products (fma): emulate by cumsum float32 of float32-rounded partial? use np.cumsum on float64 products but rounding each step to f32
    acc = np.zeros(1001, dtype=np.float32)
    for i in range(len(idx)):
        acc = (acc.astype(np.float64) + P[i]).astype(np.float32)
    raw32[c] = acc
covterm = np.zeros((n_cell,1001))
for j in range(1001):
    wj = Wd[sets[j], j].astype(np.float64)
    covterm[:, j] = C @ (B[sets[j]].T @ wj) + mu[sets[j]] @ wj
ratio = np.concatenate([[1.0], info['ratio']])
def norm(raw):
    t, c = orc.correct_background(raw[:,0], raw[:,1:], ratio[1:])
    return t, c
t64, c64 = norm(raw64 + covterm); t32, c32 = norm(raw32.astype(np.float64) + covterm)
print('raw rel err max', np.abs(raw32/raw64-1).max())
print('trait norm abs err max', np.abs(t32-t64).max(), 'ctrl norm abs err max', np.abs(c32-c64).max(), 'rms', np.sqrt(np.mean((c32-c64)**2)))
print('std/mean of raw rows', np.median(raw64[:,1:].std(axis=1)/raw64[:,1:].mean(axis=1)))
