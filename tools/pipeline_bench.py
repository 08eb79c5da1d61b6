#!/usr/bin/env python
"""BASELINE config 4 in small: the chained hot paths, query-parallel — encode corpus -> dense retrieval (TOP_K = 10) ->
near-centroid selection of one adversarial document -> SDAG generation, every stage on this library's kernels.

    python tools/pipeline_bench.py [--docs 20000] [--queries 64] [--new-tokens 32] [--gen-batch 8]
    python -m torch.distributed.run --nproc-per-node N tools/pipeline_bench.py ...    # queries split over N GPUs

Synthetic everything (no network): byte-level tokenizer (tests/synth.py), template documents of ~250 byte-tokens, a
BERT-large-geometry encoder and a Llama-3-8B-geometry generator with random weights.  Weights and corpus are replicated,
queries are partitioned, no collective on the data path; rank 0 prints one JSON line with per-stage times.
"""
import argparse
import json
import os
import random
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402

import sdag_b200.sdag as sdag  # noqa: E402
from sdag_b200.encoder import B200Ranker, encode_corpus  # noqa: E402
from sdag_b200.retrieval import DenseRetriever  # noqa: E402
from sdag_b200.selection import select_malicious_docs_for_batch  # noqa: E402

WORDS = ("alpha beta gamma delta river mountain ocean tower bridge city museum garden engine signal orbit planet crystal "
         "forest harbor market library station valley desert island canyon glacier meadow village castle").split()


def make_doc(rnd, n_bytes):
    out = []
    while sum(len(w) + 1 for w in out) < n_bytes:
        out.append(rnd.choice(WORDS))
    return " ".join(out)[:n_bytes]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=20000)
    ap.add_argument("--queries", type=int, default=64, help="per GPU")
    ap.add_argument("--new-tokens", type=int, default=32)
    ap.add_argument("--gen-batch", type=int, default=8)
    ap.add_argument("--llm-layers", type=int, default=32)
    ap.add_argument("--enc-layers", type=int, default=24)
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from transformers import BertConfig, BertModel, LlamaConfig, LlamaForCausalLM

    tok = synth.make_tokenizer()
    rnd = random.Random(42)
    docs = [make_doc(rnd, rnd.randint(200, 256)) for _ in range(args.docs)]
    meta = [{"id": f"d{i}", "text": d} for i, d in enumerate(docs)]
    qrnd = random.Random(1000 + rank)
    queries = [f"what is known about the {qrnd.choice(WORDS)} near the {qrnd.choice(WORDS)}?" for _ in range(args.queries)]
    cands = [[make_doc(qrnd, 200) for _ in range(5)] for _ in range(args.queries)]

    torch.manual_seed(42)
    with torch.device(dev):
        bert = BertModel(BertConfig(vocab_size=30522, hidden_size=1024, num_hidden_layers=args.enc_layers, num_attention_heads=16,
                                    intermediate_size=4096, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
                         add_pooling_layer=False).to(torch.bfloat16)
        prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.bfloat16)
        llm = LlamaForCausalLM(LlamaConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336,
                                           num_hidden_layers=args.llm_layers, num_attention_heads=32, num_key_value_heads=8,
                                           max_position_embeddings=8192, rope_theta=500000.0)).eval()
        torch.set_default_dtype(prev)
    ranker = B200Ranker(bert, tok, device=dev)
    sdag.prepare_model(llm)
    sdag.TEMPERATURE = 0

    def timed(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, time.perf_counter() - t0

    index, t_enc = timed(lambda: encode_corpus(ranker, docs))
    retriever = DenseRetriever(ranker, index, meta)
    stages = dict(retrieve=0.0, select=0.0, generate=0.0)
    answers = []
    # warm-up of the generation engine (graph capture) outside the timed loop
    sdag.run_rag_with_doc_isolation_batch(llm, tok, queries[: args.gen_batch], [[c[0]] for c in cands[: args.gen_batch]],
                                          [docs[:10]] * args.gen_batch, args.new_tokens, str(dev), 1, ranker, 0)
    t_loop0 = time.perf_counter()
    for q0 in range(0, args.queries, 32):  # BATCH_SIZE_EMBED_Q
        qs, cs = queries[q0 : q0 + 32], cands[q0 : q0 + 32]
        rb, dt = timed(lambda: retriever.retrieve_batch(qs, 10, 32))
        stages["retrieve"] += dt
        chosen, dt = timed(lambda: select_malicious_docs_for_batch(ranker, rb.docs_texts_full, cs, "closest_to_centroid", 1,
                                                                   device=dev, retrieved_rows_batch_full=rb.rows_full, index=index))
        stages["select"] += dt
        for g0 in range(0, len(qs), args.gen_batch):
            sl = slice(g0, g0 + args.gen_batch)
            out, dt = timed(lambda: sdag.run_rag_with_doc_isolation_batch(llm, tok, qs[sl], chosen[sl], rb.docs_texts_full[sl],
                                                                          args.new_tokens, str(dev), 1, ranker, 0))
            stages["generate"] += dt
            answers += out
    t_loop = time.perf_counter() - t_loop0
    qps = torch.tensor([args.queries / t_loop], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(qps)
    if rank == 0:
        L = len(tok(tok.apply_chat_template([{"role": "user", "content": "x"}], tokenize=False)).input_ids)
        print(json.dumps(dict(
            config="configs[3] in small: encode corpus -> dense top-10 -> closest_to_centroid (1 doc) -> SDAG generation",
            n_gpus=world, docs=args.docs, queries_per_gpu=args.queries, new_tokens=args.new_tokens, gen_batch=args.gen_batch,
            encode_corpus_s=round(t_enc, 2), encode_docs_per_s=round(args.docs / t_enc, 1),
            stage_seconds={k: round(v, 3) for k, v in stages.items()}, loop_seconds=round(t_loop, 3),
            queries_per_s_all_gpus=round(float(qps), 2), answers=len(answers), chat_overhead_tokens=L)), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
