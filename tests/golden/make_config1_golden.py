#!/usr/bin/env python
"""Golden outputs for BASELINE configs[0], produced by the REFERENCE code on the CPU:

    50 queries = first 50 of random.Random(42).sample over the 1000 NQ queries, a 1 k-document slice of the synthetic
    corpus encoded by the CPU sentence-encoder oracle (e5-large-v2 geometry, random init seed 42, fp32), dense TOP_K = 5
    through the reference's own ``dense.py`` functions over the restated IndexFlatIP, one adversarial document per query
    picked by the reference's ``select_malicious_docs_for_batch`` (closest_to_centroid) from the 5 CSV candidates, then
    ``run_rag_with_doc_isolation`` (SDAG strict, DOC_NEIGHBORS_K = 0, rank 1, greedy, max_new_tokens 16) on a
    Qwen2.5-0.5B-geometry model (24 layers, hidden 896, 14 q / 2 kv heads, FFN 4864, vocab 151 936, tied; random init seed
    42, fp32).

    python tests/golden/make_config1_golden.py          # ~20 min on 8 cores; needs /root/reference

Writes tests/golden/config1.json (picked queries, retrieved ids / scores, chosen adversarial documents, token ids, answers
and the fp32 top-2 logit gap of every greedy decision) and tests/golden/config1_emb.npz (fp16 copies of the oracle
embeddings of the 50 queries and of the first 64 corpus documents, for the encoder parity check).

Weights: both models are created on the CPU under ``torch.manual_seed`` (CPU generator: reproducible on any machine) by
``tests/synth.py:config1_encoder`` / ``config1_generator``; the GPU test builds them the same way and moves them over.
"""
import json
import os
import random
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), "/root/reference"):
    if p not in sys.path:
        sys.path.insert(0, p)
sys.modules.setdefault("sentence_transformers", types.SimpleNamespace(SentenceTransformer=object))
sys.modules.setdefault("faiss", types.SimpleNamespace(read_index=None))

import synth  # noqa: E402
from oracle.encoder import encode_oracle  # noqa: E402
from oracle.flat_ip import FlatIPOracle  # noqa: E402
from src.pipeline.attack import malicious_selection as ref_sel  # noqa: E402
from src.pipeline.retrieval import dense as ref_dense  # noqa: E402
from src.pipeline.sparse_attention_RAG import SDAG as ref_sdag  # noqa: E402
from tools.workloads import load_nq_fixture, make_corpus  # noqa: E402

N_QUERIES, N_DOCS, TOP_K, NEW_TOKENS, N_EMB_DOCS = 50, 1000, 5, 16, 64


class OracleRanker:
    """Duck-typed SentenceTransformer over the CPU encoder oracle (the three call shapes the reference uses)."""

    def __init__(self, bert, tok):
        self.bert, self.tok = bert, tok

    def get_sentence_embedding_dimension(self):
        return self.bert.config.hidden_size

    def encode(self, texts, convert_to_tensor=True, normalize_embeddings=True, batch_size=32, **_):
        toks = [self.tok(t)["input_ids"][:512] for t in list(texts)]
        emb = encode_oracle(self.bert, toks, pad_id=self.tok.pad_token_id, normalize=normalize_embeddings)
        return torch.from_numpy(emb) if convert_to_tensor else emb


class Top2Gaps:
    def __init__(self, model):
        self.calls = []
        self.handle = model.register_forward_hook(self._hook)

    def _hook(self, module, args, output):
        top = output.logits[:, -1].float().topk(2, dim=-1).values
        self.calls.append(round(float(top[0, 0] - top[0, 1]), 6))

    def take(self):
        calls, self.calls = self.calls, []
        return calls


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    tok = synth.make_tokenizer()
    queries = load_nq_fixture()
    picked = random.Random(42).sample(range(len(queries)), len(queries))[:N_QUERIES]
    docs = make_corpus(100_000)[:N_DOCS]
    meta = [{"id": f"d{i}", "text": d} for i, d in enumerate(docs)]
    ref_dense.RANKER_MODEL_NAME = "intfloat/e5-large-v2"  # e5 prefixes (dense.py:59, malicious_selection.py:32)
    ref_sel.RANKER_MODEL_NAME = "intfloat/e5-large-v2"

    bert = synth.config1_encoder()
    ranker = OracleRanker(bert, tok)
    doc_emb = ranker.encode(["passage: " + d for d in docs], convert_to_tensor=False)
    print(f"corpus slice encoded: {doc_emb.shape} in {time.time() - t0:.0f}s", flush=True)
    index = FlatIPOracle(doc_emb.shape[1])
    index.add(doc_emb)
    retriever = ref_dense.DenseRetriever(ranker, index, meta)
    q_texts = [queries[i]["question"] for i in picked]
    batch = retriever.retrieve_batch(q_texts, max_k_needed=TOP_K, embed_batch_size=32)  # reference code, unmodified
    q_emb = np.stack(batch.q_embs)
    print(f"retrieval done in {time.time() - t0:.0f}s", flush=True)
    random.seed(42)
    chosen = ref_sel.select_malicious_docs_for_batch(ranker, batch.docs_texts_full, [queries[i]["malicious_docs"] for i in picked],
                                                     "closest_to_centroid", 1)
    print(f"selection done in {time.time() - t0:.0f}s", flush=True)

    ref_sdag.TEMPERATURE = 0  # greedy (module-level import-time copy, SDAG.py:9)
    llm = synth.config1_generator()
    gaps = Top2Gaps(llm)
    records = []
    for n, qi in enumerate(picked):
        seen = {}
        orig = tok.decode

        def spy(ids, **kw):
            seen["ids"] = [int(x) for x in ids]
            return orig(ids, **kw)

        tok.decode = spy
        ans = ref_sdag.run_rag_with_doc_isolation(model=llm, tokenizer=tok, query=q_texts[n], malicious_docs=chosen[n],
                                                  retrieved_docs=batch.docs_texts_full[n], max_new_tokens=NEW_TOKENS, device="cpu",
                                                  add_attack_in_rank=1, ranker=ranker, doc_neighbors_k=0)
        tok.decode = orig
        g = gaps.take()
        chat, S, spans, qa, ranked = ref_sdag.build_rag_chat_and_spans(tok, ref_sdag.SYSTEM_PROMPT_RAG, ref_sdag.USER_RAG_PROMPT,
                                                                       q_texts[n], chosen[n], batch.docs_texts_full[n], 1)
        L = len(tok(chat).input_ids)
        records.append(dict(query_index=qi, question=q_texts[n], retrieved_ids=batch.ids_full[n], retrieved_scores=batch.scores_full[n],
                            chosen_malicious=[queries[qi]["malicious_docs"].index(c) for c in chosen[n]], L=L, S=S,
                            spans=[list(s) for s in spans], qa_start=qa, new_token_ids=seen["ids"], top2_gaps=g[: len(seen["ids"])],
                            answer=ans))
        print(f"[{n + 1}/{len(picked)}] L={L} tokens={seen['ids'][:6]}... min gap {min(g):.4f} ({time.time() - t0:.0f}s)", flush=True)
    with open(os.path.join(HERE, "config1.json"), "w") as f:
        json.dump(dict(n_docs=N_DOCS, top_k=TOP_K, new_tokens=NEW_TOKENS, picked=picked, records=records), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "config1_emb.npz"), queries=q_emb.astype(np.float16),
                        docs=doc_emb[:N_EMB_DOCS].astype(np.float16))
    print("config1 goldens written", time.time() - t0)


if __name__ == "__main__":
    main()
