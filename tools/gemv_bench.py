#!/usr/bin/env python
"""Bring-up tool: llm_gemv vs torch F.linear (cuBLAS) per decode-step layer shape of the Llama-3-8B geometry, each timed as
16 back-to-back calls on distinct weight copies inside one CUDA graph (weights never come from L2)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdag_b200._lib as L  # noqa: E402

lib = L.load()
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
shapes = [("qkv+norm", 4096, 6144, 1, 0), ("o_proj+res", 4096, 4096, 0, 0), ("gate_up+norm+swiglu", 4096, 28672, 1, 1),
          ("down+res", 14336, 4096, 0, 0), ("lm_head+norm", 4096, 128256, 1, 0)]
for name, K, N, pro, pair in shapes:
    copies = max(2, min(16, int(400e6 // (K * N * 2)) + 1))
    Ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(copies)]
    x = torch.randn(B, 2 * K if pro == 2 else K, device=dev, dtype=torch.bfloat16)
    xin = torch.randn(B, K, device=dev, dtype=torch.bfloat16)
    nw = torch.ones(K, device=dev, dtype=torch.bfloat16)
    res = torch.randn(B, N, device=dev, dtype=torch.bfloat16)
    out = torch.empty(B, N, device=dev, dtype=torch.bfloat16)
    reps = 2 * copies

    def ours(i):
        L.check(lib.llm_gemv(x.data_ptr(), Ws[i % copies].data_ptr(), 0, nw.data_ptr() if pro == 1 else 0,
                             res.data_ptr() if (pro != 1 and not pair) else 0, out.data_ptr(), B, K, N, pro, 1e-5, pair, L.stream_ptr(dev)), "gemv")

    def cublas(i):
        F.linear(xin, Ws[i % copies], out=None)

    out_sk = torch.empty(B, N // 2 if pair else N, device=dev, dtype=torch.bfloat16)

    def skinny(i):
        L.check(lib.llm_skinny_linear(x.data_ptr(), x.stride(0), Ws[i % copies].data_ptr(), 0, nw.data_ptr() if pro == 1 else 0, 1e-5,
                                      res.data_ptr() if (pro != 1 and not pair) else 0, N, out_sk.data_ptr(), out_sk.stride(0), B, N, K,
                                      pair, L.stream_ptr(dev)), "skinny")

    row = dict(layer=name, B=B, K=K, N=N, MB=round(K * N * 2 / 1e6, 1))
    # the skinny kernel against an fp32 product of the same bf16 operands (first weight copy)
    skinny(0)
    xf = x.float()
    if pro == 1:
        xf = (nw.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16).float()).to(torch.bfloat16).float()
    ref = xf @ Ws[0].float().T
    if pair:
        gte, up = ref[:, : N // 2].to(torch.bfloat16).float(), ref[:, N // 2:].to(torch.bfloat16).float()
        ref = F.silu(gte).to(torch.bfloat16).float() * up
    elif pro != 1:
        ref = ref.to(torch.bfloat16).float() + res.float()
    row["skinny_max_abs_err"] = round(float((out_sk.float() - ref).abs().max()), 5)
    row["skinny_ref_max"] = round(float(ref.abs().max()), 3)
    for label, fn in (("skinny", skinny), ("gemv", ours), ("cublas", cublas)):
        if label == "gemv" and B > 8:
            continue
        for i in range(copies):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                fn(i)
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
        row[label + "_us"] = round(us, 1)
        row[label + "_gbs"] = round(K * N * 2 / us / 1e3, 1)
    print(json.dumps(row), flush=True)
    del Ws
