#!/usr/bin/env python
"""Host-side profile of the configs[3] loop (cProfile over `workloads.config4` on a few queries): where the wall time of a
(query, strategy) result goes that is not GPU time.  python tools/pipeline_profile.py [--queries 32] [--docs 20000]"""
import argparse
import cProfile
import io
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import sdag_b200.sdag as sdag  # noqa: E402
from tools import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--queries", type=int, default=32)
    ap.add_argument("--docs", type=int, default=20000)
    args = ap.parse_args()
    dev = torch.device("cuda")
    model = sdag.prepare_model(bench.build_llama(32, dev))
    pipe = workloads.Pipeline(dev, 0, 1, n_docs=args.docs)
    workloads.config4(pipe, model, 8)  # warm everything
    pr = cProfile.Profile()
    pr.enable()
    r = workloads.config4(pipe, model, args.queries)
    pr.disable()
    print({k: v["stage_seconds"] for k, v in r["per_strategy"].items()}, r["seconds"])
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(70)
    print(s.getvalue()[:14000])
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
    print(s.getvalue()[:7000])


if __name__ == "__main__":
    main()
