#!/usr/bin/env python
"""Micro-benchmark of the SDAG prefill / decode attention kernels alone (CUDA events, q/k/v ~ N(0,1) bf16).

    python tools/attn_bench.py [--iters 20] [--cases cfg2,cfg2_causal,cfg5,...]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdag_b200.sdag as sdag  # noqa: E402


def layout(S, n_docs, doc_len, gap, qa):
    spans, pos = [], S
    for _ in range(n_docs):
        spans.append((pos, pos + doc_len))
        pos += doc_len + gap
    return pos + qa, S, spans, pos


def knn_lists(n, k, seed=42):
    g = torch.Generator().manual_seed(seed)
    e = torch.randn(n, 64, generator=g)
    e /= e.norm(dim=1, keepdim=True)
    s = e @ e.T
    s.fill_diagonal_(-9)
    return s.topk(k, dim=1).indices.tolist()


CASES = {
    "cfg2": dict(lay=(48, 10, 256, 1, 24), B=8, causal=False, K=0),
    "cfg2_b1": dict(lay=(48, 10, 256, 1, 24), B=1, causal=False, K=0),
    "cfg2_b32": dict(lay=(48, 10, 256, 1, 24), B=32, causal=False, K=0),
    "cfg2_causal": dict(lay=(48, 10, 256, 1, 24), B=8, causal=True, K=0),
    "cfg5_strict": dict(lay=(48, 50, 640, 1, 24), B=1, causal=False, K=0),
    "cfg5_K2": dict(lay=(48, 50, 640, 1, 24), B=1, causal=False, K=2),
    "cfg5_causal": dict(lay=(48, 50, 640, 1, 24), B=1, causal=True, K=0),
    "causal8k": dict(lay=(8192, 0, 0, 0, 0), B=4, causal=True, K=0),
    # overhead probes: CTAs with 0 / 1 kv tiles only
    "gaps_only": dict(raw=(2560, 0, [], 2560), B=8, causal=False, K=0),
    "docs128": dict(lay=(0, 20, 128, 0, 0), B=8, causal=False, K=0),
    "docs256": dict(lay=(0, 10, 256, 0, 0), B=8, causal=False, K=0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cases", default="cfg2,cfg2_b1,cfg2_b32,cfg2_causal,cfg5_strict,cfg5_K2,cfg5_causal,causal8k")
    ap.add_argument("--decode", action="store_true")
    args = ap.parse_args()
    H, Hkv, D = 32, 8, 128
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in args.cases.split(","):
        c = CASES[name]
        L, S, spans, qa = c["raw"] if "raw" in c else layout(*c["lay"])
        B = c["B"]
        g = torch.Generator(device=dev).manual_seed(0)
        q = torch.randn(B, L, H, D, generator=g, device=dev).to(torch.bfloat16).transpose(1, 2)
        k = torch.randn(B, Hkv, L, D, generator=g, device=dev).to(torch.bfloat16)
        v = torch.randn(B, Hkv, L, D, generator=g, device=dev).to(torch.bfloat16)
        if c["causal"]:
            bm = sdag.causal_blockmap(L, dev)
        else:
            nb = knn_lists(len(spans), c["K"]) if c["K"] else None
            bm = sdag.build_blockmap([sdag.SdagLayout(L, S, spans, qa, nb)], dev)
        hdr = bm.header(0)
        for _ in range(3):
            sdag.sdag_prefill_attention(q, k, v, bm)
        torch.cuda.synchronize()
        times = []
        for _ in range(args.iters):
            flush.zero_()  # evict L2 between timed launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sdag.sdag_prefill_attention(q, k, v, bm)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        ms = times[len(times) // 2]
        flops = 4.0 * D * hdr["pairs"] * H * B
        print(json.dumps(dict(case=name, L=L, B=B, pairs=hdr["pairs"], qtiles=hdr["n_qtiles"], kvtiles=hdr["n_kvtiles"],
                              ms=round(ms, 4), algo_tflops=round(flops / ms / 1e9, 1),
                              tile_tflops=round(4.0 * D * 128 * 128 * hdr["n_kvtiles"] * H * B / ms / 1e9, 1))), flush=True)
        if args.decode:
            # Launch-latency-free timing: `reps` decode calls captured in one CUDA graph, rotating over enough cache
            # copies that consecutive launches cannot hit in L2 (> 256 MB of distinct K/V between reuses).
            qd = q[:, :, -1:, :].contiguous()
            nbytes = 2.0 * B * Hkv * L * D * 2
            copies = max(2, min(8, int(300e6 // nbytes) + 1))
            ks = [k] + [k.clone() for _ in range(copies - 1)]
            vs = [v] + [v.clone() for _ in range(copies - 1)]
            reps = 2 * copies
            out = torch.empty((B, 1, H, D), dtype=torch.bfloat16, device=dev)
            for i in range(copies):
                sdag.sdag_decode_attention(qd, ks[i], vs[i], out=out)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for i in range(reps):
                    sdag.sdag_decode_attention(qd, ks[i % copies], vs[i % copies], out=out)
            graph.replay()
            torch.cuda.synchronize()
            times = []
            for _ in range(args.iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1) / reps)
            times.sort()
            ms = times[len(times) // 2]
            print(json.dumps(dict(case=name + "/decode", L=L, B=B, us_per_call=round(ms * 1e3, 2), gbs=round(nbytes / ms / 1e6, 1),
                                  timing=f"{reps} calls (split + merge) per graph replay, {copies} cache copies")), flush=True)
            del ks, vs

if __name__ == "__main__":
    main()
