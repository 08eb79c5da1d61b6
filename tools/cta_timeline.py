#!/usr/bin/env python
"""Bring-up tool: per-item stage timestamps of the persistent prefill kernel (needs a library built with
SDAG_EXTRA_NVCC_FLAGS=-DSDAG_DEBUG_TIMING).  Stamps are taken by softmax thread 0 of the CTA that ran the item."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdag_b200._lib as L  # noqa: E402
import sdag_b200.sdag as sdag  # noqa: E402
from tools.attn_bench import CASES, layout  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
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
n_items = bm.launch_qtiles * H * B
buf = torch.zeros(n_items * 8, dtype=torch.int64, device=dev)
for _ in range(2):
    sdag.sdag_prefill_attention(q, k, v, bm)
lib.sdag_debug_set_timing.argtypes = [ctypes.c_void_p]
assert lib.sdag_debug_set_timing(buf.data_ptr()) == 0
torch.cuda.synchronize()
sdag.sdag_prefill_attention(q, k, v, bm)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(n_items, 8).astype(np.int64)
real = t[:, 4] > 0
tr = t[real]
t0 = t[t[:, 0] > 0, 0].min()
print(f"{name}: items {n_items}, with tensor work {real.sum()}, span {(tr[:, 4].max() - t0) / 1e3:.1f} us")


def stat(x):
    return f"mean {x.mean() / 1e3:6.2f} us  p50 {np.median(x) / 1e3:6.2f}  p95 {np.percentile(x, 95) / 1e3:6.2f}"


for nk in sorted(set(tr[:, 5])):
    w = tr[tr[:, 5] == nk]
    print(f"-- items with {nk} kv tiles: {len(w)}")
    print("   item start -> first scores ready :", stat(w[:, 1] - w[:, 0]))
    print("   softmax loop                     :", stat(w[:, 2] - w[:, 1]))
    print("   wait for last P.V                :", stat(w[:, 3] - w[:, 2]))
    print("   epilogue                         :", stat(w[:, 4] - w[:, 3]))
    print("   whole item                       :", stat(w[:, 4] - w[:, 0]))
# per-CTA busy fraction
ctas = tr[:, 6]
busy = np.zeros(int(ctas.max()) + 1)
for cta in np.unique(ctas):
    w = tr[ctas == cta]
    busy[int(cta)] = (w[:, 4] - w[:, 0]).sum()
span = tr[:, 4].max() - t0
print(f"per-CTA time inside items / kernel span: mean {busy.mean() / span:.2f}, min {busy.min() / span:.2f}")
