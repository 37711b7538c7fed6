"""Where does a cold context spend its first submits?  SCDRS_B200_TRACE=1 python scratch/trace_cold.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench, scdrs_b200

args = bench.parse_args()
adata, traits, info = bench.build_workload(args, 0, torch.device("cuda", 0))
for rep in range(2):
    scdrs_b200.clear_device_cache()
    t = time.perf_counter()
    scdrs_b200.score_traits(adata, traits, n_ctrl=args.n_ctrl)
    sys.stderr.write("cold run %d: %.3f s\n" % (rep, time.perf_counter() - t))
