"""Device timeline of the trait loop (SCDRS_B200_TIMELINE=1): when does the list building of trait i + 1 run relative
to the scatter of trait i?  Config 2, device-resident control sets, the loop of bench.py's device leg."""
import os
import sys

os.environ["SCDRS_B200_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    import bench
    from scdrs_b200 import _device, method as M
    from scdrs_b200.distributed import CudaPhases, ShardedScorer

    args = bench.parse_args(["--config", "2", "--n-trait", "24"])
    device = torch.device("cuda", 0)
    adata, traits, info = bench.build_workload(args, 0, device)
    state = _device.get_state(adata)
    ctx = state.ensure_matrix(adata)
    state.ensure_cov(adata, True)
    tab = M._gene_table(adata, state)
    bins = M._GeneBins(tab.df_gene, "mean_var", 200)
    preps = [M._prepare_trait(adata, state, traits[t][0], traits[t][1], "mean_var", args.n_ctrl, 200, 0, tab, bins)
             for t in traits]
    ctx.set_gene_stats(preps[0]["var"], preps[0]["var_tech"])
    for p in preps:
        p["d_ctrl"] = torch.from_numpy(np.ascontiguousarray(p["ctrl_cols"], dtype=np.int32)).to(device)
    scorer = ShardedScorer(CudaPhases(ctx, device), device, group=False)
    ctx.set_resident_sets(True)
    for rep in range(2):
        pending = None
        for p in preps:
            ticket = scorer.submit(p["trait_cols"], p["trait_gw"], p["d_ctrl"], p["ctrl_gw"], "vs", True)
            if pending is not None:
                scorer.collect(pending)
            pending = ticket
        scorer.collect(pending)
    torch.cuda.synchronize()
    ctx.sync()


if __name__ == "__main__":
    main()
