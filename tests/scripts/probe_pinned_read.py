"""Probe: how fast do HOST threads read page-locked memory that the GPU filled by DMA, against ordinary memory?
(The slot writer of the score files reads 1.76 GB of such memory per trait.)"""
import time

import numpy as np
import torch

n = 1 << 28          # 256 MiB
dev = torch.empty(n, dtype=torch.uint8, device="cuda").random_(0, 255)
pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
pinned.copy_(dev)
torch.cuda.synchronize()
plain = torch.empty(n, dtype=torch.uint8)
plain.copy_(dev)
reg = torch.empty(n, dtype=torch.uint8)
rc = torch.cuda.cudart().cudaHostRegister(reg.data_ptr(), n, 0)
reg.copy_(dev)
torch.cuda.synchronize()
out = np.empty(n, dtype=np.uint8)
for name, t in (("cudaHostAlloc (pin_memory)", pinned), ("pageable", plain), ("malloc + cudaHostRegister rc=%s" % rc, reg)):
    a = t.numpy()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        np.copyto(out, a)
        best = min(best, time.perf_counter() - t0)
    # strided 16-byte gathers like the slot writer: view as [n/16, 16], read byte 15 of each
    v = a.reshape(-1, 16)
    t0 = time.perf_counter()
    s = int(v[:, 15].astype(np.int64).sum())
    t1 = time.perf_counter() - t0
    print("%-40s sequential copy %.2f GB/s, strided byte reads %.1f M/s" % (name, n / best / 1e9, v.shape[0] / t1 / 1e6))
