#!/usr/bin/env python
"""Bring-up tool: sub-step event trace of the prefill kernel (library built with
SDAG_EXTRA_NVCC_FLAGS=-DSDAG_DEBUG_TIMING).  Prints, for the first CTAs, the mean clock counts between the events of
softmax warp 0 (wait for S, TMEM load, arithmetic + P store) and of the MMA issuer (waits before Q.K^T and P.V)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdag_b200._lib as L  # noqa: E402
import sdag_b200.sdag as sdag  # noqa: E402
from tools.attn_bench import CASES, layout  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "causal8k"
c = CASES[name]
Lq, S, spans, qa = c["raw"] if "raw" in c else layout(*c["lay"])
B, H, Hkv, D = c["B"], 32, 8, 128
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
q = torch.randn(B, Lq, H, D, generator=g, device=dev).to(torch.bfloat16).transpose(1, 2)
k = torch.randn(B, Hkv, Lq, D, generator=g, device=dev).to(torch.bfloat16)
v = torch.randn(B, Hkv, Lq, D, generator=g, device=dev).to(torch.bfloat16)
bm = sdag.causal_blockmap(Lq, dev) if c["causal"] else sdag.build_blockmap([sdag.SdagLayout(Lq, S, spans, qa)], dev)
lib = L.load()
CTAS, EV = 8, 8192
buf = torch.zeros(CTAS * 2 * EV, dtype=torch.int64, device=dev)
for _ in range(2):
    sdag.sdag_prefill_attention(q, k, v, bm)
lib.sdag_debug_set_trace.argtypes = [ctypes.c_void_p]
assert lib.sdag_debug_set_trace(buf.data_ptr()) == 0
torch.cuda.synchronize()
sdag.sdag_prefill_attention(q, k, v, bm)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(CTAS, 2, EV)
for cta in range(min(CTAS, 4)):
    for role, label in ((0, "softmax warp 0"), (1, "MMA issuer")):
        row = t[cta, role]
        sm = int(row[0])
        ev = row[1:][row[1:] != 0]
        code, clk = ev & 0xFF, ev >> 8
        print(f"CTA {cta} (SM {sm}) {label}: {len(ev)} events, span {clk[-1] - clk[0]} clk")
        pairs = {}
        for a, b, d in zip(code[:-1], code[1:], np.diff(clk)):
            pairs.setdefault((int(a), int(b)), []).append(int(d))
        for (a, b), ds in sorted(pairs.items()):
            ds = np.array(ds)
            print(f"   {a:2d} -> {b:2d}: n {len(ds):5d}  mean {ds.mean():8.1f}  p50 {np.median(ds):7.0f}  p95 {np.percentile(ds, 95):7.0f}  total {ds.sum()}")
    if len(sys.argv) > 2:
        row = t[cta, 0]
        ev = row[1:][row[1:] != 0][: int(sys.argv[2])]
        print("   first softmax events:", [(int(e & 0xFF), int((e >> 8) - (ev[0] >> 8))) for e in ev])
        row = t[cta, 1]
        ev2 = row[1:][row[1:] != 0][: int(sys.argv[2])]
        print("   first MMA events    :", [(int(e & 0xFF), int((e >> 8) - (ev[0] >> 8))) for e in ev2])
