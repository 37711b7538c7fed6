"""The reference's ``save_intermediate`` dumps for one committed golden case, from the UNMODIFIED reference sources.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_intermediate.py
Writes tests/golden/ref_intermediate_sparse_cov_vs.npz: the ten ``.tsv.gz`` files of scdrs/method.py:507-596,
loaded back as arrays (keys = the file suffixes)."""
import glob
import os
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402
from scdrs_b200.loess1d import loess  # noqa: E402
import helpers  # noqa: E402


def main():
    pp, method = ref_shim.load(loess_impl=loess)
    path = os.path.join(HERE, "ref_case_sparse_cov_vs.npz")
    adata, cov, case = helpers.load_case(path)
    with tempfile.TemporaryDirectory() as tmp:
        prefix = os.path.join(tmp, "dump")
        method.score_cell(adata, case["gene_list"], gene_weight=case["gene_weight"], ctrl_match_key=case["ctrl_match_key"],
                          n_ctrl=case["n_ctrl"], weight_opt=case["weight_opt"], save_intermediate=prefix)
        out = {}
        for f in sorted(glob.glob(prefix + ".*.tsv.gz")):
            out[os.path.basename(f)[len("dump."):-len(".tsv.gz")]] = np.loadtxt(f, delimiter="\t")
    assert len(out) == 10, sorted(out)
    dst = os.path.join(HERE, "ref_intermediate_sparse_cov_vs.npz")
    np.savez_compressed(dst, **out)
    print(dst, os.path.getsize(dst) // 1024, "KiB", sorted(out))


if __name__ == "__main__":
    main()
