#!/usr/bin/env python
"""Generate the golden fixtures by running the REFERENCE's own code on seeded synthetic inputs.

Run in the authoring container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference modules are imported unmodified from ``/root/reference``; the two packages they import
that are not installable here are replaced by stubs *before* import:

* ``sentence_transformers`` — only the name ``SentenceTransformer`` is needed at import (SDAG.py:7,
  dense.py:8, malicious_selection.py:7); the ranker object passed in is ``tests/synth.HashRanker``.
* ``faiss`` — ``dense.py:6`` imports it; the index object passed in is ``oracle.flat_ip.FlatIPOracle``
  (so the search arithmetic itself stays "parity unpinned", see oracle/__init__.py).

Outputs (all small): masks.npz, attention.npz, generation.json, knn.json, selector.json, retrieval.json,
ranked_list.json, hybrid.json, harness.json, encoder.npz, carg.json, pipeline.json.
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")

sys.modules.setdefault("sentence_transformers", types.SimpleNamespace(SentenceTransformer=object))
sys.modules.setdefault("faiss", types.SimpleNamespace(read_index=None))

import synth  # noqa: E402
from oracle.flat_ip import FlatIPOracle  # noqa: E402
from oracle.sdag_mask import synthetic_layout  # noqa: E402
from src.pipeline.attack import malicious_selection as ref_sel  # noqa: E402
from src.pipeline.retrieval import dense as ref_dense  # noqa: E402
from src.pipeline.sparse_attention_RAG import SDAG as ref_sdag  # noqa: E402
from src.pipeline.utils import ranked_list as ref_rl  # noqa: E402
from src.pipeline.utils.prompts import SYSTEM_PROMPT_RAG, USER_RAG_PROMPT  # noqa: E402


def mask_cases():
    """(name, L, S, spans, qa_start, neighbours) — layouts incl. the quirks of SURVEY.md Appendix A."""
    cases = []
    L, S, spans, qa = synthetic_layout(5, 3, 7, 2, 6)
    cases.append(("strict_small", L, S, spans, qa, None))
    cases.append(("nbrs_directed", L, S, spans, qa, [[2], [0, 2], []]))
    cases.append(("nbrs_len_mismatch_off", L, S, spans, qa, [[1], [0]]))
    cases.append(("nbrs_oob_self_dup", L, S, spans, qa, [[0, 0, 5, -1], [1, 2, 2], [7]]))
    cases.append(("ragged", 61, 9, [(9, 12), (14, 33), (33, 40), (45, 46)], 50, [[3], [], [1, 0], [2]]))
    cases.append(("no_docs", 20, 8, [], 12, None))
    cases.append(("no_qa", 30, 4, [(4, 14), (16, 28)], 30, None))
    cases.append(("no_gaps", 5 + 3 * 8 + 4, 5, [(5, 13), (13, 21), (21, 29)], 29, [[1], [2], [0]]))
    cases.append(("empty_neighbour_lists", L, S, spans, qa, [[], [], []]))
    L, S, spans, qa = synthetic_layout(48, 5, 256, 1, 24)
    cases.append(("cfg1_k5x256", L, S, spans, qa, None))
    L, S, spans, qa = synthetic_layout(48, 10, 256, 1, 24)
    cases.append(("cfg2_k10x256", L, S, spans, qa, None))
    L, S, spans, qa = synthetic_layout(48, 10, 256, 0, 24)
    cases.append(("cfg2_nogap", L, S, spans, qa, None))
    L, S, spans, qa = synthetic_layout(130, 6, 200, 2, 40)
    rng = np.random.default_rng(42)
    emb = rng.standard_normal((6, 32)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    sims = emb @ emb.T
    nb = [[int(j) for j in np.argsort(-sims[r]) if j != r][:2] for r in range(6)]
    cases.append(("mid_k6x200_K2", L, S, spans, qa, nb))
    return cases


def gen_masks():
    small, meta = {}, []
    for name, L, S, spans, qa, nb in mask_cases():
        m = ref_sdag.build_blocked_causal_mask_full(
            base_seq_len=L, max_new_tokens=500, sys_user_len=S, doc_token_spans=spans, qa_start=qa,
            device="cpu", doc_neighbors=nb,
        ).numpy()
        packed = np.packbits(m, axis=1)
        entry = dict(name=name, L=L, S=S, spans=[list(s) for s in spans], qa_start=qa, neighbors=nb,
                     pairs=int(m.sum()), gap_rows=int((~m.any(axis=1)).sum()),
                     sha256=hashlib.sha256(packed.tobytes()).hexdigest())
        meta.append(entry)
        if L <= 256:
            small[name] = packed
    # config 5 (L = 32 122): count only (the bool mask is ~1 GiB)
    L, S, spans, qa = synthetic_layout(48, 50, 640, 1, 24)
    rng = np.random.default_rng(42)
    emb = rng.standard_normal((50, 64)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    sims = emb @ emb.T
    nb = [[int(j) for j in np.argsort(-sims[r]) if j != r][:2] for r in range(50)]
    for name, nbrs in (("cfg5_k50x640_strict", None), ("cfg5_k50x640_K2", nb)):
        m = ref_sdag.build_blocked_causal_mask_full(L, 500, S, spans, qa, "cpu", nbrs)
        meta.append(dict(name=name, L=L, S=S, spans=[list(s) for s in spans], qa_start=qa, neighbors=nbrs,
                         pairs=int(m.sum()), gap_rows=int((~m.any(dim=1)).sum()), sha256=None))
        del m
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **small)
    with open(os.path.join(HERE, "masks.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("masks:", [(e["name"], e["pairs"]) for e in meta])


def gen_attention():
    """HF eager attention fed exactly as SDAG.py:161-176 feeds it (additive finfo.min bias), fp32."""
    from transformers.models.llama.modeling_llama import eager_attention_forward

    out = {}
    for name, L, S, spans, qa, nb in mask_cases():
        if L > 256:
            continue
        H, Hkv, D = 4, 2, 16
        g = torch.Generator().manual_seed(abs(hash(name)) % (2**31))
        g = torch.Generator().manual_seed(int(hashlib.sha256(name.encode()).hexdigest()[:8], 16))
        q = torch.randn(1, H, L, D, generator=g)
        k = torch.randn(1, Hkv, L, D, generator=g)
        v = torch.randn(1, Hkv, L, D, generator=g)
        mask = ref_sdag.build_blocked_causal_mask_full(L, 500, S, spans, qa, "cpu", nb)
        neg = torch.finfo(torch.float32).min
        attn = torch.zeros_like(mask, dtype=torch.float32).masked_fill(~mask, neg).unsqueeze(0).unsqueeze(1)
        module = types.SimpleNamespace(num_key_value_groups=H // Hkv, training=False)
        o, _ = eager_attention_forward(module, q, k, v, attn, scaling=D ** -0.5)
        out[name + "/q"] = q[0].numpy()
        out[name + "/k"] = k[0].numpy()
        out[name + "/v"] = v[0].numpy()
        out[name + "/out"] = o[0].numpy()  # [L, H, D]
    np.savez_compressed(os.path.join(HERE, "attention.npz"), **out)
    print("attention cases:", len(out) // 4)


class Top2Gaps:
    """Forward hook on the reference's model: the fp32 gap between the best and the second-best logit at the last position
    of every forward call — i.e. of every greedy decision, per batch row.  Stored next to the golden token ids so that the
    GPU tests can hold every prompt to the reference's tokens and excuse a deviation only where the reference's own
    decision was a near-tie (bf16 attention moves logits by ~1e-2)."""

    def __init__(self, model):
        self.calls = []
        self.handle = model.register_forward_hook(self._hook)

    def _hook(self, module, args, output):
        top = output.logits[:, -1].float().topk(2, dim=-1).values
        self.calls.append([round(float(x), 6) for x in (top[:, 0] - top[:, 1])])

    def take(self):
        calls, self.calls = self.calls, []
        return calls

    def remove(self):
        self.handle.remove()


def gen_generation():
    ref_sdag.TEMPERATURE = 0  # greedy (module-level import-time copy)
    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok))
    ranker = synth.HashRanker()
    docs = [d for d in synth.SAMPLE_DOCS]
    records = []
    for case, retrieved, mal, rank, nbr_k in (
        ("strict_rank1", docs[:5], synth.SAMPLE_MAL[:1], 1, 0),
        ("nbrs2_rank2_no_empty", [d for d in docs if d.strip()][:5], synth.SAMPLE_MAL[:2], 2, 2),
        ("nbrs2_with_empty_doc_disables", docs[:5], synth.SAMPLE_MAL[:1], 1, 2),
        ("no_attack", docs[3:6], [], 0, 0),
    ):
        chat, S, spans, qa, ranked = ref_sdag.build_rag_chat_and_spans(
            tok, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT, synth.SAMPLE_QUERY, mal, retrieved, rank)
        seen = {}
        orig_decode = tok.decode

        def spy(ids, **kw):
            seen["ids"] = [int(x) for x in ids]
            return orig_decode(ids, **kw)

        tok.decode = spy
        gaps = Top2Gaps(model)
        ans = ref_sdag.run_rag_with_doc_isolation(
            model=model, tokenizer=tok, query=synth.SAMPLE_QUERY, malicious_docs=mal, retrieved_docs=retrieved,
            max_new_tokens=6, device="cpu", add_attack_in_rank=rank, ranker=ranker, doc_neighbors_k=nbr_k)
        top2_gaps = [c[0] for c in gaps.take()]  # one forward call per generated token (prefill, then the loop)
        gaps.remove()
        tok.decode = orig_decode
        nbrs = ref_sdag.compute_doc_knn_for_docs(ranker, ranked, nbr_k) if nbr_k > 0 else None
        ids = tok(chat, return_tensors="pt").input_ids
        L = ids.shape[1]
        # prefill logits of the last position, exactly as SDAG.py:161-181 obtains them
        mask = ref_sdag.build_blocked_causal_mask_full(L, 500, S, spans, qa, "cpu", nbrs)
        attn = torch.zeros_like(mask, dtype=torch.float32).masked_fill(~mask, torch.finfo(torch.float32).min)
        with torch.no_grad():
            logits = model(input_ids=ids, attention_mask=attn[None, None], use_cache=True).logits[0, -1]
        top = torch.topk(logits, 8)
        records.append(dict(case=case, retrieved=retrieved, malicious=mal, rank=rank, doc_neighbors_k=nbr_k,
                            chat_sha256=hashlib.sha256(chat.encode()).hexdigest(), L=L, S=S,
                            spans=[list(s) for s in spans], qa_start=qa, ranked_docs=ranked, neighbors=nbrs,
                            new_token_ids=seen["ids"], top2_gaps=top2_gaps[: len(seen["ids"])], answer=ans,
                            prefill_top_ids=top.indices.tolist(), prefill_top_logits=top.values.tolist(),
                            prefill_logits_absmax=float(logits.abs().max())))
    with open(os.path.join(HERE, "generation.json"), "w") as f:
        json.dump(records, f, indent=1)
    print("generation:", [(r["case"], r["L"], r["new_token_ids"]) for r in records])


def gen_knn():
    ranker = synth.HashRanker()
    recs = []
    for docs, k in ((synth.SAMPLE_DOCS, 2), ([d for d in synth.SAMPLE_DOCS if d.strip()], 3),
                    (synth.SAMPLE_DOCS[:1], 2), (synth.SAMPLE_DOCS, 0), (["  ", "x"], 1)):
        recs.append(dict(docs=docs, k=k, neighbors=ref_sdag.compute_doc_knn_for_docs(ranker, docs, k)))
    with open(os.path.join(HERE, "knn.json"), "w") as f:
        json.dump(recs, f, indent=1)


def gen_selector():
    ranker = synth.HashRanker()
    retrieved = [[d for d in synth.SAMPLE_DOCS], synth.SAMPLE_DOCS[3:6], [" ", ""], synth.SAMPLE_DOCS[:2]]
    cands = [synth.SAMPLE_MAL, synth.SAMPLE_MAL[:3], synth.SAMPLE_MAL[:2], []]
    recs = []
    for strategy in ("closest_to_centroid", "furthest_from_centroid"):
        for max_docs in (1, 3, None, -1, 99):
            random.seed(1234)
            chosen = ref_sel.select_malicious_docs_for_batch(ranker, retrieved, cands, strategy, max_docs)
            recs.append(dict(strategy=strategy, max_docs=max_docs, chosen=chosen))
    for strategy in ("random", "something_else"):
        for max_docs in (1, 2):
            random.seed(1234)
            chosen = ref_sel.select_malicious_docs_for_batch(ranker, retrieved, cands, strategy, max_docs)
            recs.append(dict(strategy=strategy, max_docs=max_docs, chosen=chosen, seed=1234))
    with open(os.path.join(HERE, "selector.json"), "w") as f:
        json.dump(dict(retrieved=retrieved, candidates=cands, records=recs), f, indent=1)


def gen_retrieval():
    ranker = synth.HashRanker()
    corpus = [f"document number {i} about topic {i % 7}" for i in range(40)]
    meta = [dict(id=f"doc{i}", text=t) for i, t in enumerate(corpus)]
    meta[5] = dict(text="no id here")  # id missing -> 'NA' (dense.py:120)
    emb = np.stack([synth.text_embedding("passage: " + t) for t in corpus])
    recs = []
    for n_rows, k in ((40, 6), (4, 6)):
        index = FlatIPOracle(emb.shape[1])
        index.add(emb[:n_rows])
        retr = ref_dense.DenseRetriever(ranker, index, meta[:n_rows])
        queries = ["topic 3 documents", "document number 17", synth.SAMPLE_QUERY]
        rb = retr.retrieve_batch(queries, max_k_needed=k, embed_batch_size=2)
        recs.append(dict(n_rows=n_rows, k=k, queries=queries, docs_texts_full=rb.docs_texts_full,
                         ids_full=rb.ids_full, scores_full=rb.scores_full,
                         q_embs=[np.asarray(e).tolist() for e in rb.q_embs]))
    with open(os.path.join(HERE, "retrieval.json"), "w") as f:
        json.dump(dict(corpus=corpus, meta=meta, records=recs), f, indent=1)


def gen_pipeline():
    """The reference's own functions chained the way main.py:764-796 + 467-496 chains them (config-1 shape in
    miniature): retrieve_batch -> select_malicious_docs_for_batch -> pool = first k+1 -> run_rag_with_doc_isolation
    per query, greedy, on the tiny fp32 Llama.  The retriever's index is the oracle (faiss is not installable)."""
    ref_sdag.TEMPERATURE = 0
    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok))
    ranker = synth.HashRanker()
    rng = random.Random(42)
    topics = ["rivers", "volcanoes", "composers", "planets", "bridges", "vaccines", "dynasties", "alloys"]
    corpus = [f"Passage {i}: notes on {topics[i % 8]} number {i}, with detail {rng.randint(0, 999)}." for i in range(300)]
    meta = [dict(id=f"d{i}", text=t) for i, t in enumerate(corpus)]
    emb = np.stack([synth.text_embedding("passage: " + t) for t in corpus])
    index = FlatIPOracle(emb.shape[1])
    index.add(emb)
    retr = ref_dense.DenseRetriever(ranker, index, meta)
    queries = [f"What is known about {t}?" for t in topics[:6]]
    cands = [[f"Fabricated claim {j} about {t}: the answer is {j * 7}." for j in range(5)] for t in topics[:6]]
    top_k, strategy, max_docs, rank, nbr_k = 5, "closest_to_centroid", 1, 1, 0
    rb = retr.retrieve_batch(queries, max_k_needed=top_k + 1, embed_batch_size=32)
    chosen = ref_sel.select_malicious_docs_for_batch(ranker, rb.docs_texts_full, cands, strategy, max_docs)
    answers, token_ids = [], []
    orig_decode = tok.decode
    for q, docs, mal in zip(queries, rb.docs_texts_full, chosen):
        seen = {}
        tok.decode = lambda ids, _s=seen, **kw: (_s.__setitem__("ids", [int(x) for x in ids]), orig_decode(ids, **kw))[1]
        ans = ref_sdag.run_rag_with_doc_isolation(model=model, tokenizer=tok, query=q, malicious_docs=mal,
                                                  retrieved_docs=docs[: top_k + 1], max_new_tokens=5, device="cpu",
                                                  add_attack_in_rank=rank, ranker=ranker, doc_neighbors_k=nbr_k)
        tok.decode = orig_decode
        answers.append(ans)
        token_ids.append(seen["ids"])
    with open(os.path.join(HERE, "pipeline.json"), "w") as f:
        json.dump(dict(corpus=corpus, queries=queries, candidates=cands, top_k=top_k, strategy=strategy, max_docs=max_docs,
                       rank=rank, doc_neighbors_k=nbr_k, ids_full=rb.ids_full, chosen=chosen, answers=answers,
                       token_ids=token_ids), f, indent=1)
    print("pipeline:", token_ids)


def gen_ranked_list():
    base = ["a", "b", "c", "d"]
    recs = []
    for mal, pos in ((["M1"], 1), (["M1", "M2"], 2), (["M1"], 99), (["M1"], 0), ([], 3), (["M1", "M2"], [3, 1]),
                     (["M1", "M2", "M3"], [2]), (["M1"], -1), (["M1", "M2"], [None, 2]), (["M1"], -5)):
        random.seed(7)
        recs.append(dict(base=base, malicious=mal, pos=pos, seed=7,
                         out=ref_rl.inject_malicious_docs_into_ranked_list(base, mal, pos)))
    with open(os.path.join(HERE, "ranked_list.json"), "w") as f:
        json.dump(recs, f, indent=1)


def gen_carg():
    """CARG: the reference's generate_batch_seq2seq (attack/content_generation.py:44-94, HF ``model.generate`` on a
    left-padded batch) and generate_noiso_batch (main.py:308-378; main.py imports pyserini at module level, stubbed),
    greedy, on ragged prompts."""
    for name in ("pyserini", "pyserini.search", "pyserini.search.lucene"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pyserini.search.lucene"].LuceneSearcher = object
    from src.pipeline import config as ref_config
    from src.pipeline import main as ref_main
    from src.pipeline.attack import content_generation as ref_cg

    tok = synth.make_tokenizer()
    tok.padding_side = "left"  # load_llm, content_generation.py:40
    model = synth.tiny_llama(len(tok))
    seen = []
    orig_decode = tok.decode

    def spy(ids, **kw):
        seen.append([int(x) for x in ids])
        return orig_decode(ids, **kw)

    tok.decode = spy
    raw_prompts = ["Paris is", "The highest mountain above sea level is called", "Honey", "Octopuses have three hearts and"]
    gaps = Top2Gaps(model)
    raw_out = ref_cg.generate_batch_seq2seq(tok, model, raw_prompts, max_tokens=7, temperature=0, top_p=1.0)
    raw_gaps = [[call[b] for call in gaps.calls] for b in range(len(raw_prompts))]  # [prompt][step]
    gaps.take()
    raw_ids, seen[:] = list(seen), []
    ref_config.ORACLE, ref_config.LLM_BATCH_SIZE, ref_config.MAX_GEN_TOKENS_RAG = True, 2, 6
    ref_config.TEMPERATURE, ref_config.TOP_P = 0, 1.0
    docs = [d for d in synth.SAMPLE_DOCS]
    queries = [synth.SAMPLE_QUERY, "Which ocean is the largest?", "How many hearts does an octopus have?"]
    defended = [docs[:3], docs[3:6], docs[4:8]]
    mals = [synth.SAMPLE_MAL[:1], [], synth.SAMPLE_MAL[1:3]]
    random.seed(5)
    answers = ref_main.generate_noiso_batch(tok, model, queries, defended, mals, attacker_pos=2, order_mode="bottom_up")
    # sub-batches of LLM_BATCH_SIZE = 2 prompts, each a run of consecutive forward calls (a call's batch size tells them apart)
    noiso_gaps, batch_start = [], 0
    calls = gaps.take()
    gaps.remove()
    ci = 0
    while batch_start < len(queries):
        nb = min(2, len(queries) - batch_start)
        steps = []
        while ci < len(calls) and len(calls[ci]) == nb and len(steps) < 6:
            steps.append(calls[ci])
            ci += 1
        noiso_gaps += [[st[b] for st in steps] for b in range(nb)]
        batch_start += nb
    tok.decode = orig_decode
    with open(os.path.join(HERE, "carg.json"), "w") as f:
        json.dump(dict(raw_prompts=raw_prompts, raw_max_tokens=7, raw_answers=raw_out, raw_token_ids=raw_ids,
                       queries=queries, defended=defended, malicious=mals, attacker_pos=2, order_mode="bottom_up",
                       llm_batch_size=2, max_gen_tokens=6, answers=answers, token_ids=list(seen),
                       raw_top2_gaps=raw_gaps, top2_gaps=noiso_gaps), f, indent=1)
    print("carg:", raw_ids, list(seen))


def gen_encoder():
    """Sentence embeddings of the e5 module stack restated on HF BertModel (oracle/encoder.py): random-init tiny BERT,
    fp32, eager attention, CPU — ragged inputs from one token to several tiles long."""
    from oracle.encoder import encode_oracle

    tok = synth.make_tokenizer()
    bert = synth.tiny_bert(len(tok))
    toks = [tok(t)["input_ids"][:512] for t in synth.ENCODER_TEXTS]
    emb = encode_oracle(bert, toks, pad_id=tok.pad_token_id)
    raw = encode_oracle(bert, toks, pad_id=tok.pad_token_id, normalize=False)
    np.savez_compressed(os.path.join(HERE, "encoder.npz"), emb=emb.astype(np.float32), raw=raw.astype(np.float32),
                        lengths=np.array([len(t) for t in toks], dtype=np.int32))
    print("encoder:", emb.shape, [len(t) for t in toks])


def gen_hybrid():
    """Reference RRF fusion (retrieval/hybrid.py) and ranked-list ordering on seeded synthetic lists."""
    from src.pipeline.retrieval import hybrid as ref_hybrid

    rnd = random.Random(11)
    pool = [f"doc {i} text" for i in range(30)]
    cases = []
    for top_k in (1, 4, 5, 10, 11):
        sp_t, sp_i, de_t, de_i = [], [], [], []
        for _q in range(4):
            s_idx = rnd.sample(range(30), 12)
            d_idx = rnd.sample(range(30), 12)
            sp_t.append([pool[i] for i in s_idx])
            de_t.append([pool[i] for i in d_idx])
            # ids: mostly real, some "NA" / "" (fall back to text keys), one duplicated id inside a list
            sp_i.append([("NA" if i % 7 == 0 else "" if i % 11 == 0 else f"id{i}") for i in s_idx])
            de_i.append([("NA" if i % 7 == 0 else "" if i % 11 == 0 else f"id{i}") for i in d_idx])
            sp_i[-1][3] = sp_i[-1][1]
        t, i, sc = ref_hybrid.fuse_sparse_and_dense_batch(sp_t, sp_i, de_t, de_i, top_k=top_k, seed=42, k0=60)
        cases.append(dict(top_k=top_k, seed=42, k0=60, sparse_texts=sp_t, sparse_ids=sp_i, dense_texts=de_t, dense_ids=de_i,
                          fused_texts=t, fused_ids=i, fused_scores=sc))
    splits = []
    rng = random.Random(5)
    for top_k in (0, 1, 2, 3, 7, 10, 11):
        splits.append(dict(top_k=top_k, split=list(ref_hybrid.split_k_between_sparse_and_dense(top_k, rng))))
    orders = []
    for mode in ("top_down", "bottom_up", "random", "other"):
        random.seed(3)
        orders.append(dict(mode=mode, seed=3, docs=list("abcdefg"), out=ref_rl.apply_ranked_list_order(list("abcdefg"), mode)))
    with open(os.path.join(HERE, "hybrid.json"), "w") as f:
        json.dump(dict(cases=cases, splits=splits, split_seed=5, orders=orders), f, indent=1)
    print("hybrid:", len(cases), "cases")


def harness_inputs():
    """Seeded synthetic result rows, answer strings and an attack CSV for the evaluation harness (shared with the test)."""
    rnd = random.Random(23)
    facts = ["Paris", "the Eiffel Tower", "1889", "Gustave Eiffel", "an iron lattice", "Île-de-France", "A. G. Eiffel & Co."]
    fakes = ["London", "1901", "a steel arch", "Rome"]
    filler = ["It rains often.", "The river is long!", "Nobody knows, really.", "An apple a day.", "Towers are tall; bridges are long."]
    rows = []
    for q in range(40):
        gts = rnd.sample(facts, rnd.randint(1, 2))
        fas = rnd.sample(fakes, rnd.randint(0, 2))
        docs = []
        for _ in range(rnd.randint(0, 7)):
            d = rnd.choice(filler)
            if rnd.random() < 0.35:
                d += " " + rnd.choice(gts).upper() + "."
            docs.append(d)
        rows.append({
            "query_id": f"q{q}", "question": f"question {q}?", "short_answers": gts if q % 9 else gts[0],
            "false_answer": fas if q % 4 else (fas[0] if fas else ""), "malicious_doc": "bad doc" if q % 3 else "",
            "retrieved_docs": docs if q % 13 else None, "retrieved_doc_ids": [f"d{i}" for i in range(len(docs))],
            "rag_answer_iso": rnd.choice(gts + fas + ["unknown"]), "rag_answer_noiso": rnd.choice(gts + fas + ["unknown"]),
            "ground_truth_match_iso": rnd.random() < 0.5, "ground_truth_match_noiso": rnd.random() < 0.4,
            "false_match_iso": rnd.random() < 0.2, "false_match_noiso": rnd.random() < 0.5,
        })
    texts = ["The Eiffel Tower!", "  an  APPLE, a day... ", "Île-de-France", "<think>hmm the answer</think>Paris", None,
             "Answer: Paris\nbecause", "- Final Answer:  1889  ", "\n\n  \nsecond line\nthird", "<THINK>x</THINK> final answer: Rome",
             "", "a an the", "naïve café — déjà vu", "A.G. Eiffel & Co."]
    csv_rows = [
        ("who built it", "7", '["Gustave Eiffel", "Eiffel"]', "Rome", "doc one"),
        ("who built it", "7", '["ignored"]', "Rome", "doc two"),
        ("who built it", "7", "", "London", ""),
        ("when", "8", "['1889', '1887']", "", "doc three"),
        ("where", "9", "Paris ||| France", "Berlin", "doc four"),
        ("what", "10", "iron, steel", " ", "doc five"),
        ("why", "11", "because", "x", "doc six"),
        ("when", "8", "1889", "1900", "doc seven"),
    ]
    return rows, texts, csv_rows


def gen_harness():
    """Reference evaluation utilities (utils/normalization.py, metrics.py, parsing.py, save_results.py) on seeded synthetic rows."""
    import contextlib
    import csv
    import io
    import tempfile

    from src.pipeline.utils import metrics as ref_metrics
    from src.pipeline.utils import normalization as ref_norm
    from src.pipeline.utils import parsing as ref_parsing
    from src.pipeline.utils import save_results as ref_save

    rows, texts, csv_rows = harness_inputs()
    out = {}
    out["normalize"] = [ref_norm.normalize_answer(t) for t in texts if t is not None]
    out["extract"] = [ref_norm.extract_final_answer(t) for t in texts]
    out["exact"] = [[bool(ref_metrics.exact_match(a, b)) for b in texts] for a in texts]
    out["exact_norm_module"] = [[bool(ref_norm.exact_match(a, b)) for b in texts if b is not None] for a in texts if a is not None]
    out["parse_list"] = [ref_parsing.parse_list_field(c) for c in [r[2] for r in csv_rows] + [None, "  ", "[1, 2]", "{'a': 1}"]]
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "attack.csv")
        with open(path, "w", encoding="utf-8", newline="") as f:
            w = csv.writer(f)
            w.writerow(["query", "query_id", "ground_truth_answers", "false_answer", "malicious_document"])
            w.writerows(csv_rows)
        out["load_from_csv"] = [list(x) for x in ref_parsing.load_from_csv(path)]
        out["load_from_csv_by_id"] = [list(x) for x in ref_parsing.load_from_csv(path, match_field_for_groups="query_id")]
        full = [dict(r, short_answers=ref_metrics.ensure_list(r["short_answers"]),
                     false_answer=r["false_answer"] if isinstance(r["false_answer"], str) else " / ".join(r["false_answer"]),
                     retrieved_docs=r["retrieved_docs"] or []) for r in rows]
        csv_path = os.path.join(td, "sub", "res.csv")
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref_save.save_results(full, csv_path)
        out["save_results_csv"] = open(csv_path, encoding="utf-8", newline="").read()
        out["save_results_stdout"] = buf.getvalue()
    out["retrieval_stats"] = {str(k): ref_metrics.compute_retrieval_ground_truth_stats(rows, k) for k in (1, 3, 5)}
    out["retrieval_stats_empty"] = ref_metrics.compute_retrieval_ground_truth_stats([], 4)
    out["overlap"] = ref_metrics.compute_answer_overlap_and_attack_stats(rows)
    out["overlap_empty"] = ref_metrics.compute_answer_overlap_and_attack_stats([])
    out["pair_metrics"] = ref_metrics.build_pair_metrics(rows, 5, 2)
    out["pair_metrics_empty"] = ref_metrics.build_pair_metrics([], 3, 0)
    out["false_answer_stats"] = ref_metrics.compute_false_answer_stats_for_results(rows)
    with open(os.path.join(HERE, "harness.json"), "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False)
    print("harness:", len(rows), "rows")


if __name__ == "__main__":
    if len(sys.argv) > 1:  # regenerate selected fixtures only: make_golden.py hybrid ...
        torch.manual_seed(0)
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    torch.manual_seed(0)
    gen_masks()
    gen_attention()
    gen_generation()
    gen_knn()
    gen_selector()
    gen_retrieval()
    gen_ranked_list()
    gen_hybrid()
    gen_harness()
    gen_encoder()
    gen_carg()
    gen_pipeline()
    print("golden fixtures written to", HERE)
