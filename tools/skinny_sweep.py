#!/usr/bin/env python
"""Bring-up: llm_skinny_linear streaming rate per layer shape under the SDAG_SKINNY_{SK,R,KS,STAGES,SMEM_KB,NOMMA} knobs (read once per process)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdag_b200._lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shapes = {"qkv": (4096, 6144, 1, 0), "o": (4096, 4096, 0, 0), "gate_up": (4096, 28672, 1, 1), "down": (14336, 4096, 0, 0),
          "head": (4096, 128256, 1, 0), "head_nonorm": (4096, 128256, 0, 0)}
which = sys.argv[2].split(",") if len(sys.argv) > 2 else list(shapes)
for name in which:
    K, N, norm, pair = shapes[name]
    copies = max(2, min(16, int(400e6 // (K * N * 2)) + 1))
    Ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(copies)]
    x = torch.randn(B, K, device=dev, dtype=torch.bfloat16)
    nw = torch.ones(K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(B, N // 2 if pair else N, device=dev, dtype=torch.bfloat16)
    reps = 2 * copies

    def run(i):
        L.check(lib.llm_skinny_linear(x.data_ptr(), K, Ws[i % copies].data_ptr(), 0, nw.data_ptr() if norm else 0, 1e-5, 0, 0,
                                      out.data_ptr(), out.stride(0), B, N, K, pair, L.stream_ptr(dev)), "skinny")

    for i in range(copies):
        run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            run(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    us = sorted(ts)[len(ts) // 2] * 1e3
    print(json.dumps(dict(layer=name, B=B, SK=os.environ.get("SDAG_SKINNY_SK", ""), R=os.environ.get("SDAG_SKINNY_R", ""), KS=os.environ.get("SDAG_SKINNY_KS", ""),
                          ST=os.environ.get("SDAG_SKINNY_STAGES", ""), nomma=os.environ.get("SDAG_SKINNY_NOMMA", ""),
                          us=round(us, 1), gbs=round(K * N * 2 / us / 1e3, 1))), flush=True)
    del Ws
