#!/usr/bin/env python
"""Bring-up tool: per-CTA stage timestamps of the prefill kernel (needs a library built with
SDAG_EXTRA_NVCC_FLAGS=-DSDAG_DEBUG_TIMING)."""
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
grid = bm.launch_qtiles * H * B
buf = torch.zeros(grid * 8, dtype=torch.int64, device=dev)
for _ in range(2):
    sdag.sdag_prefill_attention(q, k, v, bm)
lib.sdag_debug_set_timing.argtypes = [ctypes.c_void_p]
assert lib.sdag_debug_set_timing(buf.data_ptr()) == 0
torch.cuda.synchronize()
sdag.sdag_prefill_attention(q, k, v, bm)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(grid, 8).astype(np.int64)
real = t[:, 1] > 0
t0 = t[t[:, 0] > 0, 0].min()
print(f"{name}: grid {grid}, real CTAs {real.sum()}, kernel span {(t[real][:, 6].max() - t0) / 1e3:.1f} us")
tr = t[real]
n_kv_known = tr[:, 2] > 0
def stat(x):
    return f"mean {x.mean() / 1e3:7.2f} us  p50 {np.median(x) / 1e3:7.2f}  p95 {np.percentile(x, 95) / 1e3:7.2f}"
print("setup (start -> after tmem alloc+sync):", stat(tr[:, 1] - tr[:, 0]))
w = tr[n_kv_known]
print("first S ready (after setup):          ", stat(w[:, 2] - w[:, 1]))
print("first softmax:                        ", stat(w[:, 3] - w[:, 2]))
print("rest of loop until O ready:           ", stat(w[:, 4] - w[:, 3]))
print("epilogue:                             ", stat(tr[:, 5] - tr[:, 4]))
print("teardown:                             ", stat(tr[:, 6] - tr[:, 5]))
print("whole CTA:                            ", stat(tr[:, 6] - tr[:, 0]))
starts = np.sort(tr[:, 0] - t0)
print("CTA start times (us) percentiles 10/50/90/100:", [round(float(np.percentile(starts, p)) / 1e3, 1) for p in (10, 50, 90, 100)])
