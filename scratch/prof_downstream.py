"""One group analysis + one correlation analysis at config-2 size (for an ncu launch list)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import helpers  # noqa: E402
import scdrs_b200  # noqa: E402

adata, df = helpers.make_downstream_inputs(110096, 1000, 120, k=15, seed=1, dtype=np.float32)
m = scdrs_b200.method
for _ in range(int(os.environ.get("REPS", "1"))):
    res = m.downstream_group_analysis(adata, df, ["grp"])["grp"]
    corr = m.downstream_corr_analysis(adata, df, ["v1"])
print(res.head(3), corr, sep="\n")
