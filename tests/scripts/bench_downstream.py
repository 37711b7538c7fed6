"""Timing of the downstream analyses at config-2 size (110,096 cells, n_ctrl = 1000, 120 cell groups, a
symmetric 15-neighbour graph): device path of scdrs_b200.method.downstream_* against the numpy port
(oracle/downstream_oracle.py, vectorised -- the reference itself loops over cells in Python,
scdrs/method.py:1073-1079) on a bounded sample of the groups.  Writes one JSON object.

    python tests/scripts/bench_downstream.py [--n-cell N] [--n-ctrl C] [--n-group G] [--cpu-groups K]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import helpers  # noqa: E402
import scdrs_b200  # noqa: E402
from scdrs_b200 import _native  # noqa: E402
from oracle import downstream_oracle as dso  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-cell", type=int, default=110096)
    ap.add_argument("--n-ctrl", type=int, default=1000)
    ap.add_argument("--n-group", type=int, default=120)
    ap.add_argument("--cpu-groups", type=int, default=3)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    t0 = time.perf_counter()
    adata, df = helpers.make_downstream_inputs(a.n_cell, a.n_ctrl, a.n_group, k=15, seed=1, dtype=np.float32)
    gen_s = time.perf_counter() - t0
    G = adata.obsp["connectivities"]

    calls = {}
    for name in ("ds_set_scores", "ds_group_stats", "ds_corr"):
        orig = getattr(_native.Context, name)

        def timed(self, *args, _orig=orig, _name=name, **kw):
            t = time.perf_counter()
            out = _orig(self, *args, **kw)
            calls.setdefault(_name, []).append(time.perf_counter() - t)
            return out

        setattr(_native.Context, name, timed)

    m = scdrs_b200.method
    res = {}
    for label, fn in (("group_analysis", lambda: m.downstream_group_analysis(adata, df, ["grp"])),
                      ("corr_analysis", lambda: m.downstream_corr_analysis(adata, df, ["v1"]))):
        times = []
        for _ in range(a.reps):
            calls.clear()
            t = time.perf_counter()
            out = fn()
            times.append(time.perf_counter() - t)
            last_calls = {k: float(np.sum(v)) for k, v in calls.items()}
        res[label] = {"wall_s_best": min(times), "wall_s_all": times, "c_abi_calls_s_last": last_calls}
    got = m.downstream_group_analysis(adata, df, ["grp"])["grp"]

    # CPU port on the first few (largest) groups, scaled by the cells x edges they cover
    order = sorted(adata.obs_names)
    sub, dfo = adata[order], df.loc[order]
    labels = sub.obs["grp"].values
    ctrl = dfo[[c for c in df.columns if c.startswith("ctrl_norm_score")]].values
    groups = sorted(set(labels))
    sizes = np.array([(labels == g).sum() for g in groups])
    pick = [groups[i] for i in np.argsort(-sizes)[1:1 + a.cpu_groups]]        # skip the giant first group
    mask = np.isin(labels, pick)
    t = time.perf_counter()
    want = dso.group_analysis(sub.obsp["connectivities"][mask][:, mask], labels[mask], dfo["norm_score"].values[mask],
                              ctrl[mask], dfo["pval"].values[mask])
    cpu_s = time.perf_counter() - t
    frac = mask.sum() / len(labels)
    err = float(np.nanmax(np.abs(got.loc[want.index, ["assoc_mcz", "hetero_mcz"]].values.astype(float) /
                                  want[["assoc_mcz", "hetero_mcz"]].values.astype(float) - 1)))
    out = {
        "workload": "%d cells x (1 + %d) score columns, %d groups (largest %d cells), graph %d edges" % (
            a.n_cell, a.n_ctrl, len(groups), sizes.max(), G.nnz),
        "device": res, "generate_s": gen_s,
        "cpu_port": {"seconds": cpu_s, "groups": len(pick), "cells": int(mask.sum()), "fraction_of_cells": frac,
                     "extrapolated_full_s": cpu_s / frac, "cores": 1,
                     "note": "oracle/downstream_oracle.group_analysis (vectorised numpy); the reference's gearys_c is a Python loop over cells"},
        "max_rel_diff_mcz_vs_port": err,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
