import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, warnings
warnings.filterwarnings('ignore')
import scdrs_b200 as sc
from oracle import scdrs_oracle as orc
from scdrs_b200 import synth
from scdrs_b200.loess1d import loess
adata, _ = synth.make_adata(300, 1200, density=0.1, seed=6)
orc.preprocess(adata, cov=None, loess_impl=loess)
genes, _ = synth.make_traits(adata.var_names, 1, 50, seed=4)["trait_00"]
df = sc.score_cell(adata, genes, n_ctrl=1, return_ctrl_norm_score=True, return_ctrl_raw_score=True)
print(df.head(8)); print(np.isnan(df.norm_score).sum(), np.isinf(df.norm_score).sum())
