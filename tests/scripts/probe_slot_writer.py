"""Probe: float writer against slot writer at config-2 size, with the slots in ordinary and in page-locked memory."""
import os
import time

import numpy as np
import pandas as pd
import torch

from scdrs_b200 import _native

n, m = 110096, 1000
rng = np.random.default_rng(0)
vals = rng.normal(size=(n, m)).astype(np.float32)
lead = rng.normal(size=(n, 6)).astype(np.float32)
idx = ["cell_%d" % i for i in range(n)]
cols = ["c%d" % i for i in range(m)]
df_lead = pd.DataFrame(lead, index=idx, columns=["l%d" % i for i in range(6)])
df_full = pd.DataFrame(np.concatenate([lead, vals], 1), index=idx, columns=list(df_lead.columns) + cols)
slots = np.zeros((n, m, 16), np.uint8)
nt = os.cpu_count()
t0 = time.time()
_native.lib().scdrs_format_slots_host(n * m, vals.ctypes.data, slots.ctypes.data, nt)
print("host slot formatting on %d threads: %.2f s" % (nt, time.time() - t0))
pinned_t = torch.empty(slots.shape, dtype=torch.uint8, pin_memory=True)
pinned = pinned_t.numpy()
pinned[:] = slots
dev = torch.from_numpy(slots).cuda()
dma_t = torch.empty(slots.shape, dtype=torch.uint8, pin_memory=True)
dma_t.copy_(dev)
torch.cuda.synchronize()
dma = dma_t.numpy()
for rep in range(2):
    t0 = time.time(); _native.write_score_gz("/tmp/f.gz", df_full); a = time.time() - t0
    t0 = time.time(); _native.write_score_gz_slots("/tmp/s1.gz", df_lead, cols, slots); b = time.time() - t0
    t0 = time.time(); _native.write_score_gz_slots("/tmp/s2.gz", df_lead, cols, pinned); c = time.time() - t0
    t0 = time.time(); _native.write_score_gz_slots("/tmp/s3.gz", df_lead, cols, dma); d = time.time() - t0
    print("float writer %.2f s | slot writer: ordinary memory %.2f s, pinned (CPU-filled) %.2f s, pinned (DMA-filled) %.2f s" % (a, b, c, d))
