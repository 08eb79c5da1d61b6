#!/usr/bin/env python
"""Generation micro-benchmark (Llama-3-8B geometry, random init, bf16, batch of 1 like the reference loop):
prefill latency and decode ms/token for the fused engine (CUDA-graph decode) vs the per-token Hugging Face loop with
the same attention kernels.  python tools/gen_bench.py [--cases cfg2,cfg5] [--new 64] [--layers 32]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import sdag_b200.sdag as sdag  # noqa: E402
from sdag_b200.llama_runner import get_runner  # noqa: E402
from tools.attn_bench import knn_lists, layout  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="cfg2,cfg5")
    ap.add_argument("--new", type=int, default=64)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--hf", action="store_true", help="also time the per-token HF loop")
    ap.add_argument("--batch", type=int, default=1, help="prompts per call (3-8: the cuBLAS / SDAG_DECODE_GEMV step)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    model = sdag.prepare_model(bench.build_llama(args.layers, dev))
    runner = get_runner(model)
    assert runner is not None
    for name in args.cases.split(","):
        lay, K = {"cfg2": ((48, 10, 256, 1, 24), 0), "cfg5": ((48, 50, 640, 1, 24), 2)}[name]
        L, S, spans, qa = layout(*lay)
        nb = knn_lists(len(spans), K) if K else None
        Bn = args.batch
        ids = torch.randint(0, 128000, (Bn, L), generator=torch.Generator().manual_seed(1)).to(dev)
        bm = sdag.build_blockmap([sdag.SdagLayout(L, S, spans, qa, nb)] * Bn, dev)
        res = dict(case=name, L=L, batch=Bn, new_tokens=args.new, layers=args.layers, decode_gemv=os.environ.get("SDAG_DECODE_GEMV", ""))

        def timed(n_new):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            runner.generate(ids, bm, n_new, None)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) * 1e3

        timed(args.new)  # builds the decode context (cache, captured graph) for this bucket
        timed(args.new)
        res["prefill_ms"] = min(timed(0) for _ in range(3))
        res["generate_ms"] = min(timed(args.new) for _ in range(3))
        res["decode_ms_per_token"] = (res["generate_ms"] - res["prefill_ms"]) / args.new
        res["graph"] = runner._last_ctx["graph"] is not None
        kv_bytes = 2 * L * 8 * 128 * 2 * args.layers * Bn
        w_bytes = sum(p.numel() for p in model.parameters()) * 2 - model.model.embed_tokens.weight.numel() * 2
        res["decode_hbm_gbs"] = (kv_bytes + w_bytes) / res["decode_ms_per_token"] / 1e6
        if args.hf:
            sdag.USE_FUSED_PREFILL = False
            try:
                with torch.no_grad():
                    for rep in range(2):
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        logits, past = sdag.prefill_forward(model, ids, bm)
                        torch.cuda.synchronize()
                        t1 = time.perf_counter()
                        nxt = logits.argmax(-1, keepdim=True)
                        for _ in range(args.new):
                            o = model(input_ids=nxt, past_key_values=past, use_cache=True)
                            past = o.past_key_values
                            nxt = o.logits[:, -1].argmax(-1, keepdim=True)
                        torch.cuda.synchronize()
                        t2_ = time.perf_counter()
                res["hf_prefill_ms"] = (t1 - t0) * 1e3
                res["hf_decode_ms_per_token"] = (t2_ - t1) * 1e3 / args.new
            finally:
                sdag.USE_FUSED_PREFILL = True
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
