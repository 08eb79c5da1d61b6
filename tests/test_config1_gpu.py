"""BASELINE configs[0] on the GPU against goldens the REFERENCE code produced on the CPU (tests/golden/config1.json,
config1_emb.npz — made by tests/golden/make_config1_golden.py): 50 sampled NQ queries, a 1 k-document slice of the synthetic
corpus, e5-large-v2-geometry encoder, TOP_K = 5, one adversarial document per query (closest_to_centroid) at rank 1, SDAG
strict, Qwen2.5-0.5B-geometry generator, greedy, 16 new tokens.

Both models are created on the CPU generator exactly as the golden script creates them (tests/synth.py), then moved to the
GPU, so the weights are the reference run's weights."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle.flat_ip import FlatIPOracle

import synth
from test_generate_gpu import GAP_EXCUSE, tokens_match_or_excused

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "config1.json")) as f:
        g = json.load(f)
    emb = np.load(os.path.join(golden_dir, "config1_emb.npz"))
    return g, emb


@pytest.fixture(scope="module")
def corpus_and_queries(gold):
    from tools.workloads import load_nq_fixture, make_corpus

    g, _ = gold
    queries = load_nq_fixture()
    picked = random.Random(42).sample(range(len(queries)), len(queries))[:50]
    assert picked == g["picked"]
    docs = make_corpus(100_000)[: g["n_docs"]]
    return docs, queries, picked


@pytest.fixture(scope="module")
def ranker():
    from sdag_b200.encoder import B200Ranker

    return B200Ranker(synth.config1_encoder(), synth.make_tokenizer(), device="cuda")


def test_config1_encoder_embeddings_match_cpu_oracle(gold, corpus_and_queries, ranker):
    """bf16 encoder on the GPU vs the fp32 CPU oracle embeddings of the golden run: 50 queries + the first 64 documents.
    Tolerance (bf16 compute through 24 layers, stated in DESIGN §4): cosine >= 0.999, max-abs 3e-2 on unit vectors."""
    g, emb = gold
    docs, queries, picked = corpus_and_queries
    q_texts = ["query: " + queries[i]["question"] for i in picked]
    d_texts = ["passage: " + d for d in docs[: emb["docs"].shape[0]]]
    got_q = ranker.encode(q_texts, convert_to_tensor=True, normalize_embeddings=True).cpu().numpy()
    got_d = ranker.encode(d_texts, convert_to_tensor=True, normalize_embeddings=True).cpu().numpy()
    for got, ref in ((got_q, emb["queries"].astype(np.float32)), (got_d, emb["docs"].astype(np.float32))):
        cos = (got * ref).sum(1) / np.linalg.norm(ref, axis=1)
        assert cos.min() >= 0.999, cos.min()
        assert np.abs(got - ref).max() <= 3e-2


def test_config1_retrieval_top5_exact_on_gpu_embeddings_and_close_to_reference_run(gold, corpus_and_queries, ranker):
    """encode corpus -> FlatIPIndex -> DenseRetriever.retrieve_batch (TOP_K = 5): ids / scores equal the independent fp32
    oracle over the SAME embeddings (the search itself is exact); against the reference run (fp32 CPU embeddings) the top-5
    sets are reported — a random-init encoder packs all scores into a narrow band, so bf16-vs-fp32 embedding noise may
    reorder them, and the scores must agree to the embedding tolerance."""
    from sdag_b200.encoder import encode_corpus
    from sdag_b200.retrieval import DenseRetriever

    g, _ = gold
    docs, queries, picked = corpus_and_queries
    meta = [{"id": f"d{i}", "text": d} for i, d in enumerate(docs)]
    index = encode_corpus(ranker, docs)
    assert index.ntotal == len(docs)
    retriever = DenseRetriever(ranker, index, meta)
    q_texts = [queries[i]["question"] for i in picked]
    batch = retriever.retrieve_batch(q_texts, max_k_needed=g["top_k"], embed_batch_size=32)
    oracle = FlatIPOracle(index.d)
    oracle.add(index.reconstruct_rows(torch.arange(index.ntotal)).cpu().numpy())
    D, I = oracle.search(np.stack(batch.q_embs), g["top_k"])
    ids = np.array([[int(s[1:]) for s in row] for row in batch.ids_full])
    scores = np.array(batch.scores_full, dtype=np.float32)
    np.testing.assert_allclose(scores, D, rtol=1e-5, atol=1e-6)
    assert (ids == I).mean() >= 0.99 and np.allclose(scores[ids != I], D[ids != I], rtol=2e-6)  # near-ties only
    overlap = np.mean([len(set(r["retrieved_ids"]) & set(row)) / g["top_k"] for r, row in zip(g["records"], batch.ids_full)])
    ref_scores = np.array([r["retrieved_scores"] for r in g["records"]], dtype=np.float32)
    print(f"top-5 overlap with the reference run's retrieval (fp32 CPU embeddings): {overlap:.2f}")
    assert np.abs(scores[:, 0] - ref_scores[:, 0]).max() <= 3e-2  # best score per query within the embedding tolerance


def test_config1_sdag_answers_match_reference_run(gold, corpus_and_queries, monkeypatch):
    """run_rag_with_doc_isolation with the reference run's retrieved documents and chosen adversarial document, on the
    reference run's fp32 weights (Hugging Face forward + this library's attention, whose arithmetic is bf16): every query's
    greedy tokens equal the reference's; a deviation is accepted only where the reference's own fp32 top-2 logit gap is at
    most GAP_EXCUSE.

    A random-init model over a 151 936-word vocabulary is a hard case for greedy parity: the reference's own decisions
    have a median top-2 gap of 0.18 and a third of them are below 0.1 (tests/golden/config1.json), so many 17-token
    continuations meet a near-tie somewhere; what is asserted is that NO deviation happens at a decision the reference
    made with a clear margin (observed on a B200: every flip sits at a reference gap <= 0.04), and that the first token
    (the prefill's decision, SDAG.py:181) agrees on >= 80 % of the queries.  The counts are printed."""
    import sdag_b200.sdag as sdag

    g, _ = gold
    docs, queries, picked = corpus_and_queries
    doc_text = {f"d{i}": d for i, d in enumerate(docs)}
    tok = synth.make_tokenizer()
    model = synth.config1_generator().to("cuda")
    monkeypatch.setattr(sdag, "TEMPERATURE", 0)
    recs = g["records"]
    equal = flips = first_ok = 0
    for lo in range(0, len(recs), 10):
        chunk = recs[lo : lo + 10]
        seen = []
        orig = tok.decode
        tok.decode = lambda i, **kw: (seen.append([int(x) for x in i]), orig(i, **kw))[1]
        try:
            answers = sdag.run_rag_with_doc_isolation_batch(
                model, tok, [r["question"] for r in chunk],
                [[queries[r["query_index"]]["malicious_docs"][c] for c in r["chosen_malicious"]] for r in chunk],
                [[doc_text[d] for d in r["retrieved_ids"]] for r in chunk], max_new_tokens=g["new_tokens"], device="cuda",
                add_attack_in_rank=1, ranker=None, doc_neighbors_k=0)
        finally:
            tok.decode = orig
        assert len(seen) == len(chunk)
        for r, mine, ans in zip(chunk, seen, answers):
            assert len(mine) == g["new_tokens"] + 1 or tok.eos_token_id in mine[1:]
            first_ok += int(mine[0] == r["new_token_ids"][0])
            if tokens_match_or_excused(mine, r["new_token_ids"], r["top2_gaps"], f"query {r['query_index']}"):
                assert ans == r["answer"]
                equal += 1
            else:
                flips += 1
    print(f"config 1 greedy parity: {equal} of {len(recs)} answers token-for-token equal, {flips} with an excused near-tie flip, "
          f"first token equal on {first_ok}")
    assert first_ok >= int(0.8 * len(recs))


def test_config1_bf16_engine_matches_hf_bf16_forward(gold, corpus_and_queries):
    """The fused engine on the config-1 generator geometry (Qwen2.5-0.5B: q/k/v biases, 14 q / 2 kv heads of 64 — the
    unfused q/k/v + RoPE path —, tied embeddings) in bf16 against the Hugging Face forward of the SAME bf16 model with this
    library's attention hook: last-position prefill logits within bf16 GEMM-order noise, the first token equal unless the
    Hugging Face run's own top-2 gap is below 0.1.  (Against the fp32 reference run a bf16 MODEL flips at gaps up to 0.55:
    model-precision noise, not kernel error — the fp32-weight test above is the parity check.)"""
    import sdag_b200.sdag as sdag
    from sdag_b200.llama_runner import get_runner
    from sdag_b200.prompts import SYSTEM_PROMPT_RAG, USER_RAG_PROMPT

    g, _ = gold
    docs, queries, picked = corpus_and_queries
    doc_text = {f"d{i}": d for i, d in enumerate(docs)}
    tok = synth.make_tokenizer()
    model = sdag.prepare_model(synth.config1_generator().to("cuda", dtype=torch.bfloat16))
    runner = get_runner(model)
    assert runner is not None and "llm_gemm" in runner.path
    agree = 0
    recs = g["records"][:12]
    for r in recs:
        mal = [queries[r["query_index"]]["malicious_docs"][c] for c in r["chosen_malicious"]]
        chat, S, spans, qa, _ = sdag.build_rag_chat_and_spans(tok, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT, r["question"], mal,
                                                              [doc_text[d] for d in r["retrieved_ids"]], 1)
        ids = tok(chat, return_tensors="pt")["input_ids"].cuda()
        assert ids.shape[1] == r["L"] and S == r["S"] and [list(s) for s in spans] == r["spans"]  # the reference run's layout
        bm = sdag.build_blockmap([sdag.SdagLayout(ids.shape[1], S, spans, qa)])
        fused, _ = runner.prefill(ids, bm, want_cache=False)
        with torch.no_grad():
            hf = model(input_ids=ids, use_cache=False, sdag_meta=bm, logits_to_keep=1).logits[:, -1]
        torch.testing.assert_close(fused.float(), hf.float(), atol=0.25, rtol=0.05)  # bf16 through 24 layers, two GEMM orders; observed max 0.15
        top2 = hf.float().topk(2, dim=-1).values[0]
        same = int(fused.argmax(-1)) == int(hf.argmax(-1))
        assert same or float(top2[0] - top2[1]) < 0.1
        agree += int(same)
    print(f"bf16 engine vs HF bf16 forward: first token equal on {agree} of {len(recs)} prompts")
    assert agree >= len(recs) - 3
