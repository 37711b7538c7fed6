"""torchrun -N 2: score_traits(sharded=True) with draws dealt over ranks must equal every rank drawing everything."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import scdrs_b200
from scdrs_b200 import synth
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
scdrs_b200.set_default_device(lr)
dist.init_process_group("nccl", device_id=dev)
adata, cov = synth.make_adata(3000, 4000, density=0.1, seed=50 + rank)
scdrs_b200.preprocess(adata, cov=cov)
box = [{k: adata.uns["SCDRS_PARAM"][k] for k in ["GENE_STATS", "COV_BETA", "COV_GENE_MEAN"]} if rank == 0 else None]
dist.broadcast_object_list(box, src=0); adata.uns["SCDRS_PARAM"].update(box[0])
traits = synth.make_traits(adata.var_names, 7, 300, weighted=False, seed=3)
traits.update({"w" + k: v for k, v in synth.make_traits(adata.var_names, 3, 250, weighted=True, seed=4).items()})
os.environ["SCDRS_B200_LOCAL_DRAWS"] = "1"
a = scdrs_b200.score_traits(adata, traits, n_ctrl=200, sharded=True)
os.environ["SCDRS_B200_LOCAL_DRAWS"] = "0"
b = scdrs_b200.score_traits(adata, traits, n_ctrl=200, sharded=True)
for t in traits:
    assert np.array_equal(a[t].values, b[t].values, equal_nan=True), t
print("rank", rank, "shared draws == local draws for", len(traits), "traits")
dist.destroy_process_group()
