#!/usr/bin/env python
"""Throughput of the sentence-encoder path: BERT-large geometry (e5-large-v2: 24 layers, hidden
1024, 16 heads, FFN 4096), random-init bf16, synthetic documents of 64-256 tokens (config 1's corpus law), embeddings
pooled + normalised straight into an index shard.  Prints docs/s and tokens/s for this library's runner and for the
Hugging Face forward (sdpa attention, padded batches of 32 sorted by length — what sentence-transformers runs) on the
same GPU.

    python tools/encode_bench.py [--docs 8192] [--layers 24]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdag_b200.encoder import B200Ranker  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=8192)
    ap.add_argument("--layers", type=int, default=24)
    ap.add_argument("--max-batch-tokens", type=int, default=32768)
    args = ap.parse_args()
    from transformers import BertConfig, BertModel

    cfg = BertConfig(vocab_size=30522, hidden_size=1024, num_hidden_layers=args.layers, num_attention_heads=16,
                     intermediate_size=4096, max_position_embeddings=512, hidden_dropout_prob=0.0,
                     attention_probs_dropout_prob=0.0)
    torch.manual_seed(42)
    with torch.device("cuda"):
        bert = BertModel(cfg, add_pooling_layer=False).to(torch.bfloat16).eval()

    class Tok:  # token lists are synthetic: only the pad id is needed
        pad_token_id = 0

    ranker = B200Ranker(bert, Tok(), max_batch_tokens=args.max_batch_tokens)
    g = torch.Generator().manual_seed(42)
    lens = torch.randint(64, 257, (args.docs,), generator=g).tolist()
    toks = [torch.randint(1000, 30000, (n,), generator=g).tolist() for n in lens]
    n_tok = sum(lens)
    out = torch.empty((args.docs, 1024), dtype=torch.float32, device="cuda")
    shadow = torch.empty((args.docs, 1024), dtype=torch.bfloat16, device="cuda")

    def ours():
        ranker.encode_tokens_into(toks, out, shadow)

    @torch.no_grad()
    def hf():
        order = sorted(range(len(toks)), key=lambda i: -len(toks[i]))
        for i in range(0, len(order), 32):
            idx = order[i : i + 32]
            w = len(toks[idx[0]])
            ids = torch.zeros((len(idx), w), dtype=torch.long)
            mask = torch.zeros((len(idx), w), dtype=torch.long)
            for r, j in enumerate(idx):
                ids[r, : lens[j]] = torch.tensor(toks[j])
                mask[r, : lens[j]] = 1
            ids, mask = ids.cuda(non_blocking=True), mask.cuda(non_blocking=True)
            h = bert(input_ids=ids, attention_mask=mask).last_hidden_state
            m = mask.unsqueeze(-1).to(h.dtype)
            e = (h * m).sum(1) / m.sum(1).clamp(min=1e-9)
            out[torch.tensor(idx, device="cuda")] = torch.nn.functional.normalize(e.float(), p=2, dim=1)

    res = {}
    for name, fn in (("sdag_b200", ours), ("hf_sdpa_padded32", hf)):
        if name != "sdag_b200":
            bert.set_attn_implementation("sdpa")
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        res[name] = dict(ms=round(ms, 1), docs_per_s=round(args.docs / ms * 1e3, 1), tokens_per_s=round(n_tok / ms * 1e3, 1))
    # 2 * params-in-linear-layers FLOP per token (attention FLOPs excluded): 24 * (4 h^2 + 2 h * 4h) * 2
    flops_tok = args.layers * (4 * 1024 * 1024 + 2 * 1024 * 4096) * 2
    res["linear_tflops_sdag_b200"] = round(flops_tok * n_tok / res["sdag_b200"]["ms"] / 1e9, 1)
    res.update(docs=args.docs, tokens=n_tok, layers=args.layers)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
