import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, pandas as pd, warnings
warnings.filterwarnings('ignore')
import helpers, scdrs_b200 as sc
from oracle import scdrs_oracle as orc
from scdrs_b200.loess1d import loess
for cov in (True, False):
    a1, df_cov, df_gs, ref, stats = helpers.load_toy()
    a2 = a1.copy()
    sc.preprocess(a1, cov=df_cov if cov else None)
    orc.preprocess(a2, cov=df_cov if cov else None, loess_impl=loess)
    g1, g2 = a1.uns['SCDRS_PARAM']['GENE_STATS'], a2.uns['SCDRS_PARAM']['GENE_STATS']
    for c in ['mean','var','ct_mean','ct_var','ct_var_tech','var_tech']:
        x, y = g1[c].values.astype(float), g2[c].values.astype(float)
        print(cov, c, 'max rel', np.nanmax(np.abs(x-y)/(np.abs(y)+1e-300)), 'argmax', np.nanargmax(np.abs(x-y)/(np.abs(y)+1e-300)))
    print('bins', (g1['mean_var'].astype(str).values == g2['mean_var'].astype(str).values).mean())
    c1, c2 = a1.uns['SCDRS_PARAM']['CELL_STATS'], a2.uns['SCDRS_PARAM']['CELL_STATS']
    print('cell', np.abs(c1.values.astype(float)/c2.values.astype(float)-1).max(axis=0))
    if cov:
        print('beta', np.abs(a1.uns['SCDRS_PARAM']['COV_BETA'].values - a2.uns['SCDRS_PARAM']['COV_BETA'].values).max())
