"""End-to-end drop-in check (BASELINE config 1 / 4 in miniature): the three replaced entry points chained the way the
reference's main.py chains them — retrieve_batch -> select_malicious_docs_for_batch -> run_rag_with_doc_isolation —
against tests/golden/pipeline.json, produced by chaining the reference's own functions on the same inputs."""
import json
import os

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


def test_chained_pipeline_matches_reference_golden(golden_dir):
    import sdag_b200.sdag as sdag
    from sdag_b200.retrieval import DenseRetriever, FlatIPIndex
    from sdag_b200.selection import select_malicious_docs_for_batch

    with open(os.path.join(golden_dir, "pipeline.json")) as f:
        g = json.load(f)
    sdag.TEMPERATURE = 0
    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok)).cuda()
    ranker = synth.HashRanker()
    meta = [dict(id=f"d{i}", text=t) for i, t in enumerate(g["corpus"])]
    emb = np.stack([synth.text_embedding("passage: " + t) for t in g["corpus"]])
    index = FlatIPIndex(emb.shape[1])
    index.add(emb)
    retr = DenseRetriever(ranker, index, meta)
    k = g["top_k"]
    rb = retr.retrieve_batch(g["queries"], max_k_needed=k + 1, embed_batch_size=32)
    assert rb.ids_full == g["ids_full"]
    chosen = select_malicious_docs_for_batch(ranker, rb.docs_texts_full, g["candidates"], g["strategy"], g["max_docs"])
    assert chosen == g["chosen"]
    exact = 0
    orig = tok.decode
    for q, docs, mal, want_ids, want_ans in zip(g["queries"], rb.docs_texts_full, chosen, g["token_ids"], g["answers"]):
        seen = {}
        tok.decode = lambda ids, _s=seen, **kw: (_s.__setitem__("ids", [int(x) for x in ids]), orig(ids, **kw))[1]
        ans = sdag.run_rag_with_doc_isolation(model=model, tokenizer=tok, query=q, malicious_docs=mal, retrieved_docs=docs[: k + 1],
                                              max_new_tokens=5, device="cuda", add_attack_in_rank=g["rank"], ranker=ranker,
                                              doc_neighbors_k=g["doc_neighbors_k"])
        tok.decode = orig
        assert seen["ids"][0] == want_ids[0]  # the prefill token must always agree
        if seen["ids"] == want_ids:
            assert ans == want_ans
            exact += 1
    # fp32 reference on CPU vs bf16 attention on the GPU: a chaotic random-init model may flip a late near-tie
    assert exact >= len(g["queries"]) - 1, exact


def test_qwen2_geometry_runs_through_fused_engine():
    """Qwen2.5-0.5B head geometry (14 q heads / 2 kv heads, head_dim 64, qkv biases) in miniature: fused runner vs
    the HF forward with the attention hook."""
    from transformers import Qwen2Config, Qwen2ForCausalLM

    import sdag_b200.sdag as sdag
    from sdag_b200.llama_runner import get_runner

    cfg = Qwen2Config(vocab_size=300, hidden_size=896, intermediate_size=512, num_hidden_layers=2, num_attention_heads=14,
                      num_key_value_heads=2, max_position_embeddings=4096, tie_word_embeddings=True)
    torch.manual_seed(1)
    model = Qwen2ForCausalLM(cfg).to(torch.bfloat16).cuda().eval()
    sdag.prepare_model(model)
    assert get_runner(model) is not None
    from oracle.sdag_mask import synthetic_layout

    L, S, spans, qa = synthetic_layout(20, 3, 70, 2, 11)
    ids = torch.randint(0, 300, (2, L), generator=torch.Generator().manual_seed(3)).cuda()
    bm = sdag.build_blockmap([sdag.SdagLayout(L, S, spans, qa, [[1], [2, 0], []])])
    sdag.USE_FUSED_PREFILL = False
    try:
        ref, _ = sdag.prefill_forward(model, ids, bm)
    finally:
        sdag.USE_FUSED_PREFILL = True
    got, _ = sdag.prefill_forward(model, ids, bm)
    torch.testing.assert_close(got.float(), ref.float(), atol=0.06, rtol=0.05)
    assert torch.equal(got.argmax(-1), ref.argmax(-1))
    out = get_runner(model).generate(ids, bm, 12, None)
    assert out.shape == (2, 13)


def test_main_shaped_harness_writes_reference_csv_and_json(golden_dir, tmp_path):
    """sdag_b200.harness.main over the miniature pipeline: two (TOP_K, ADD_ATTACK_IN_RANK) pairs, the per-pair CSV / JSON
    files of the reference's main.py; the ISO answers equal the per-query run_rag_with_doc_isolation answers and the
    metrics equal build_pair_metrics over the CSV's own rows."""
    import csv

    import sdag_b200.sdag as sdag
    from sdag_b200 import config, harness
    from sdag_b200.retrieval import DenseRetriever, FlatIPIndex

    with open(os.path.join(golden_dir, "pipeline.json")) as f:
        g = json.load(f)
    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok)).cuda()
    ranker = synth.HashRanker()
    meta = [dict(id=f"d{i}", text=t) for i, t in enumerate(g["corpus"])]
    emb = np.stack([synth.text_embedding("passage: " + t) for t in g["corpus"]])
    index = FlatIPIndex(emb.shape[1])
    index.add(emb)
    retr = DenseRetriever(ranker, index, meta)
    nq = len(g["queries"])
    saved = {k: getattr(config, k) for k in ("TOP_K", "ADD_ATTACK_IN_RANK", "MAX_GEN_TOKENS_RAG", "TEMPERATURE", "BATCH_SIZE_EMBED_Q",
                                              "MALICIOUS_DOC_SELECTION_STRATEGY", "MAX_MALICIOUS_DOCS_PER_QUERY", "DOC_NEIGHBORS_K")}
    old_temp = sdag.TEMPERATURE
    try:
        config.TOP_K, config.ADD_ATTACK_IN_RANK = [g["top_k"], 2], [g["rank"], 0]
        config.MAX_GEN_TOKENS_RAG, config.TEMPERATURE, config.BATCH_SIZE_EMBED_Q = 5, 0.0, 3
        config.MALICIOUS_DOC_SELECTION_STRATEGY, config.MAX_MALICIOUS_DOCS_PER_QUERY = g["strategy"], g["max_docs"]
        config.DOC_NEIGHBORS_K = g["doc_neighbors_k"]
        sdag.TEMPERATURE = 0
        sdag.MAX_GEN_TOKENS_RAG = 5
        gts = [[w] for w in ("tower", "honey", "paris", "river", "moon", "stone", "apple", "cloud")][:nq]
        out = harness.main(retriever=retr, llm_model=model, tokenizer=tok, ranker=ranker, queries=g["queries"], short_answers=gts,
                           query_ids=[f"q{i}" for i in range(nq)], false_answer_groups=[["rome"]] * nq,
                           malicious_doc_groups=g["candidates"], output_base=str(tmp_path / "run"))
    finally:
        for k, v in saved.items():
            setattr(config, k, v)
        sdag.TEMPERATURE = old_temp
    assert set(out) == {(g["top_k"], g["rank"]), (2, 0)}
    for (k, pos), metrics in out.items():
        stem = tmp_path / f"run_top_k={k}_attacker_pos={pos}"
        rows = list(csv.DictReader(open(str(stem) + ".csv", encoding="utf-8", newline="")))
        assert [r["query_id"] for r in rows] == [f"q{i}" for i in range(nq)] and list(rows[0]) == list(harness.RESULT_FIELDS)
        # the pool is the first k + 1 documents of a retrieval done once at the deepest k (main.py:153-169, 433-437)
        assert all(len(r["retrieved_doc_ids"].split("\n")) == min(k + 1, g["top_k"]) for r in rows)
        on_disk = json.load(open(str(stem) + ".json", encoding="utf-8"))
        assert on_disk["num_queries"] == nq and on_disk["top_k"] == k and on_disk["attack_position_in_rank"] == pos
        assert on_disk["run_config"]["TOP_K"] == [g["top_k"], 2]
        gt_iso = sum(int(r["ground_truth_match_iso"]) for r in rows)
        assert on_disk["answer_match_stats"]["iso"]["ground_truth_match_count"] == gt_iso
        assert (pos != 0) == any(r["malicious_doc"] for r in rows)
    # the ISO column is what the per-query entry point answers for the same pool (first line of the answer)
    from sdag_b200.selection import select_malicious_docs_for_batch

    rows = list(csv.DictReader(open(str(tmp_path / f"run_top_k={g['top_k']}_attacker_pos={g['rank']}") + ".csv", encoding="utf-8", newline="")))
    rb = retr.retrieve_batch(g["queries"], max_k_needed=g["top_k"], embed_batch_size=3)
    chosen = select_malicious_docs_for_batch(ranker, rb.docs_texts_full, g["candidates"], g["strategy"], g["max_docs"])
    old_temp, sdag.TEMPERATURE = sdag.TEMPERATURE, 0
    try:
        want = [sdag.run_rag_with_doc_isolation(model=model, tokenizer=tok, query=q, malicious_docs=mal, retrieved_docs=docs[: g["top_k"] + 1],
                                                max_new_tokens=5, device="cuda", add_attack_in_rank=g["rank"], ranker=ranker,
                                                doc_neighbors_k=g["doc_neighbors_k"])
                for q, docs, mal in zip(g["queries"], rb.docs_texts_full, chosen)]
    finally:
        sdag.TEMPERATURE = old_temp
    assert [harness.extract_final_answer(w) for w in want] == [r["rag_answer_iso"] for r in rows]
    assert [r["malicious_doc"] for r in rows] == [" ||| ".join(c) for c in chosen]
