#!/usr/bin/env python
"""Dense top-k throughput vs query-batch size on one GPU (the reference pipeline calls the search with
32 queries at a time, where the problem is HBM-bound: one pass over the bf16 corpus per batch).

    python tools/ip_bench.py [--rows 10000000] [--queries 1,8,32,128,512,4096]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdag_b200.retrieval import FlatIPIndex  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--queries", default="1,8,32,128,512,4096")
    ap.add_argument("--topk", type=int, default=10)
    args = ap.parse_args()
    dev, d = torch.device("cuda"), 1024
    gen = torch.Generator(device=dev).manual_seed(42)
    x = torch.empty((args.rows, d), dtype=torch.float32, device=dev)
    for r0 in range(0, args.rows, 1 << 20):
        r1 = min(args.rows, r0 + (1 << 20))
        x[r0:r1].normal_(generator=gen)
        x[r0:r1] /= x[r0:r1].norm(dim=1, keepdim=True)
    index = FlatIPIndex.from_device(x)
    for Q in [int(q) for q in args.queries.split(",")]:
        xq = torch.randn(Q, d, generator=gen, device=dev)
        xq /= xq.norm(dim=1, keepdim=True)
        for _ in range(2):
            index.search_device(xq, args.topk)
        torch.cuda.synchronize()
        n = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            index.search_device(xq, args.topk)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps(dict(Q=Q, rows=args.rows, ms=round(ms, 3), qps=round(Q / ms * 1e3, 1),
                              corpus_gbs=round(args.rows * d * 2 / ms / 1e6, 1),
                              tflops=round(2.0 * Q * args.rows * d / ms / 1e9, 1))), flush=True)


if __name__ == "__main__":
    main()
