"""The hot path once, for profiling (ncu launch list / --set full captures; see profiles/README.md): the synthetic
data set of BASELINE config 2 (110,096 x 20,000), preprocess (K4 kernels), then `--n-trait` traits through
score_traits with the score-file text produced on the device."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-trait", type=int, default=2)
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--weighted", action="store_true")
    a = ap.parse_args()
    import torch

    import bench
    import scdrs_b200

    argv = ["--config", str(a.config), "--n-trait", str(a.n_trait)] + (["--weighted"] if a.weighted else [])
    args = bench.parse_args(argv)
    adata, traits, info = bench.build_workload(args, 0, torch.device("cuda", 0))
    seen = []
    scdrs_b200.score_traits(adata, traits, n_ctrl=args.n_ctrl, return_ctrl_norm_score=True, text_output=True,
                            on_result=lambda t, df: seen.append((t, float(df["norm_score"].mean()))))
    print("profiled path ok:", info["preprocess_phases_s"], seen)


if __name__ == "__main__":
    main()
