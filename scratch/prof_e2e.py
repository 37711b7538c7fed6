import sys, time; sys.path.insert(0,'.')
import numpy as np, torch, warnings
warnings.filterwarnings('ignore')
import bench, scdrs_b200
from scdrs_b200 import method as M, _device
class A: pass
args = bench.parse_args()
dev = torch.device('cuda',0)
adata, traits, info = bench.build_workload(args, 0, dev)
print(info)
names = list(traits)
t=time.perf_counter(); scdrs_b200.clear_device_cache(); st=_device.get_state(adata); ctx=st.ensure_matrix(adata); st.ensure_cov(adata, True); ctx.sync(); print('upload X+cov', time.perf_counter()-t)
t=time.perf_counter(); p=M._prepare_trait(adata, st, *traits[names[0]], "mean_var", 1000, 200, 0); print('prep first (cold caches)', time.perf_counter()-t)
ts=[]
for n in names[1:6]:
    t=time.perf_counter(); p=M._prepare_trait(adata, st, *traits[n], "mean_var", 1000, 200, 0); ts.append(time.perf_counter()-t)
print('prep warm', np.mean(ts))
import cProfile, pstats
pr=cProfile.Profile(); pr.enable(); p=M._prepare_trait(adata, st, *traits[names[7]], "mean_var", 1000, 200, 0); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
ctx.set_gene_stats(p['var'], p['var_tech'])
ts=[]; tf=[]
for i in range(5):
    t=time.perf_counter(); out=ctx.score_trait(p["trait_cols"], p["trait_gw"], p["ctrl_cols"], p["ctrl_gw"], "vs", True); ts.append(time.perf_counter()-t)
    t=time.perf_counter(); df=M._result_frame(adata, out, 1000, False, False); tf.append(time.perf_counter()-t)
print('score_trait call', np.mean(ts), 'result_frame', np.mean(tf), 'timing', ctx.last_timing())
t=time.perf_counter(); res=scdrs_b200.score_traits(adata, traits, n_ctrl=1000); print('score_traits total', time.perf_counter()-t)
for nj in (4, 8, 15):
    t=time.perf_counter(); res=scdrs_b200.score_traits(adata, traits, n_ctrl=1000, n_jobs=nj); print('score_traits n_jobs', nj, time.perf_counter()-t)
