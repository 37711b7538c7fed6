import sys, time; sys.path.insert(0,'.')
import numpy as np, torch, warnings, cProfile, pstats
warnings.filterwarnings('ignore')
import bench, scdrs_b200
args = bench.parse_args()
adata, traits, info = bench.build_workload(args, 0, torch.device('cuda',0))
print(info)
from scdrs_b200 import synth
cov = synth.make_covariates(adata.X, seed=7); cov.index = adata.obs.index
scdrs_b200.clear_device_cache()
pr=cProfile.Profile(); pr.enable(); scdrs_b200.preprocess(adata, cov=cov); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
