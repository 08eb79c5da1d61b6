"""Edge cases of the drop-in surface on the GPU path: empty inputs, limits, degenerate prompts (SURVEY.md §8c asks for the
cases the domain has: empty and ragged inputs, maximum sizes, missing hits)."""
import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


def test_retrieval_edges():
    from sdag_b200._lib import SdagB200Error
    from sdag_b200.retrieval import DenseRetriever, FlatIPIndex

    x = torch.randn(1000, 64, device="cuda")
    x /= x.norm(dim=1, keepdim=True)
    idx = FlatIPIndex.from_device(x)
    D, I = idx.search(np.zeros((0, 64), np.float32), 5)
    assert D.shape == (0, 5) and I.shape == (0, 5)
    D, I = FlatIPIndex.from_device(x[:3].contiguous()).search(x[:2].cpu().numpy(), 5)
    assert (I[:, 3:] == -1).all() and (I[:, 0] == [0, 1]).all() and (D[:, 3:] == -np.finfo(np.float32).max).all()
    assert idx.search(x[:2].cpu().numpy(), 64)[1].shape == (2, 64)
    with pytest.raises(SdagB200Error):
        idx.search(x[:2].cpu().numpy(), 65)
    D, I = FlatIPIndex(64).search(x[:2].cpu().numpy(), 3)  # empty index: faiss returns -1 ids
    assert (I == -1).all()
    meta = [{"id": str(i), "text": f"doc {i}"} for i in range(1000)]
    with pytest.raises(ValueError):  # np.vstack([]) in the reference's embed_queries (dense.py:65) raises the same
        DenseRetriever(synth.HashRanker(), idx, meta).retrieve_batch([], 5, 32)


def test_selector_and_encoder_edges():
    from sdag_b200.encoder import B200Ranker, encode_corpus
    from sdag_b200.selection import select_malicious_docs_for_batch

    ranker = synth.HashRanker()
    assert select_malicious_docs_for_batch(ranker, [], [], "closest_to_centroid", 1) == []
    assert select_malicious_docs_for_batch(ranker, [["a", "b"]], [[]], "closest_to_centroid", 1) == [[]]
    tok = synth.make_tokenizer()
    br = B200Ranker(synth.tiny_bert(len(tok)), tok)
    assert br.encode([], convert_to_tensor=True, normalize_embeddings=True).shape == (0, 128)
    assert bool(br.encode([""], convert_to_tensor=True, normalize_embeddings=True).isfinite().all())  # zero tokens: zero vector
    assert encode_corpus(br, []).ntotal == 0
    long = br.encode(["x" * 5000, "x" * 512], convert_to_tensor=True, normalize_embeddings=True)  # truncated at 512 tokens
    torch.testing.assert_close(long[0], long[1], atol=1e-6, rtol=0)


def test_generation_edges():
    import sdag_b200.sdag as sdag
    from sdag_b200.generation import generate_batch_seq2seq

    sdag.TEMPERATURE = 0
    tok = synth.make_tokenizer()
    tok.padding_side = "left"
    model = synth.tiny_llama(len(tok), dtype=torch.bfloat16, hidden=256, heads=4, kv_heads=2, head_dim=64, inter=512,
                             init_scale=3.0).cuda()
    ranker = synth.HashRanker()
    assert generate_batch_seq2seq(tok, model, [], 4, 0, 1.0) == []
    one = generate_batch_seq2seq(tok, model, ["ab", "c"], 1, 0, 1.0)
    assert len(one) == 2 and all(len(tok(t)["input_ids"]) <= 1 for t in one)
    # no documents at all / only blank documents (dropped by the prompt builder): the prompt is system + question only
    a = sdag.run_rag_with_doc_isolation(model, tok, "q?", [], [], 4, "cuda", 1, ranker, 0)
    b = sdag.run_rag_with_doc_isolation(model, tok, "q?", [], ["", "  "], 4, "cuda", 1, ranker, 2)
    assert a == b
    assert sdag.run_rag_with_doc_isolation_batch(model, tok, [], [], [], 4, "cuda", 1, ranker, 0) == []
    # max_new_tokens = 0 still yields the prefill token (SDAG.py:181-189 picks it before the loop)
    seen = {}
    orig = tok.decode
    tok.decode = lambda i, **kw: (seen.__setitem__("ids", [int(t) for t in i]), orig(i, **kw))[1]
    try:
        sdag.run_rag_with_doc_isolation(model, tok, "q?", ["m"], ["d1", "d2"], 0, "cuda", 1, ranker, 0)
    finally:
        tok.decode = orig
    assert len(seen["ids"]) == 1
