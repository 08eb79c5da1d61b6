"""GPU parity of the sentence-encoder path (B200Ranker: BERT forward on the prefill attention kernel under a
bidirectional block map, fused LayerNorm / GELU, mean-pool + normalise into the index shard) against the fp32 CPU
oracle and its golden fixture.  Tolerance: the product path computes in bf16 (fp32 accumulation / statistics), the
oracle in fp32 — cosine >= 0.999 and max-abs 3e-2 on unit vectors, stated here, not bit parity."""
import os

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu


def test_layernorm_gelu_pool_kernels_match_torch():
    import sdag_b200._lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    rows, dim = 77, 1024
    x = torch.randn(rows, dim, generator=g, device="cuda").to(torch.bfloat16)
    r = torch.randn(rows, dim, generator=g, device="cuda").to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(dim, generator=g, device="cuda")).to(torch.bfloat16)
    b = (0.1 * torch.randn(dim, generator=g, device="cuda")).to(torch.bfloat16)
    out = torch.empty_like(x)
    L.check(lib.enc_layernorm(x.data_ptr(), r.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), rows, dim, 1e-12,
                              L.stream_ptr(x.device)), "ln")
    ref = torch.nn.functional.layer_norm((x + r).float(), (dim,), w.float(), b.float(), 1e-12)
    torch.testing.assert_close(out.float(), ref, atol=2e-2, rtol=2e-2)
    assert (out == ref.to(torch.bfloat16)).float().mean() > 0.98
    L.check(lib.enc_layernorm(x.data_ptr(), 0, w.data_ptr(), b.data_ptr(), out.data_ptr(), rows, dim, 1e-12,
                              L.stream_ptr(x.device)), "ln no residual")
    torch.testing.assert_close(out.float(), torch.nn.functional.layer_norm(x.float(), (dim,), w.float(), b.float(), 1e-12),
                               atol=2e-2, rtol=2e-2)
    y = torch.empty_like(x)
    L.check(lib.enc_gelu(x.data_ptr(), y.data_ptr(), x.numel(), L.stream_ptr(x.device)), "gelu")
    ref = torch.nn.functional.gelu(x.float())  # exact erf form; the kernel's erf is accurate to 1.5e-7
    assert (y == ref.to(torch.bfloat16)).float().mean() > 0.98
    torch.testing.assert_close(y.float(), ref, atol=1e-6, rtol=2 ** -7)  # never more than one bf16 ulp away
    wide = torch.linspace(-12, 12, 4096, device="cuda").to(torch.bfloat16)
    yw = torch.empty_like(wide)
    L.check(lib.enc_gelu(wide.data_ptr(), yw.data_ptr(), wide.numel(), L.stream_ptr(wide.device)), "gelu tails")
    torch.testing.assert_close(yw.float(), torch.nn.functional.gelu(wide.float()), atol=1e-6, rtol=2 ** -7)
    # pooling: ragged lengths (0 included), scattered destination rows, fp32 + bf16 outputs
    B, Lm = 5, 40
    h = torch.randn(B, Lm, dim, generator=g, device="cuda").to(torch.bfloat16)
    lens = torch.tensor([40, 1, 17, 0, 33], dtype=torch.int32, device="cuda")
    rows_dst = torch.tensor([4, 0, 2, 1, 3], dtype=torch.int64, device="cuda")
    o32 = torch.zeros(B, dim, dtype=torch.float32, device="cuda")
    o16 = torch.zeros(B, dim, dtype=torch.bfloat16, device="cuda")
    L.check(lib.enc_mean_pool_normalize(h.data_ptr(), h.stride(0), h.stride(1), lens.data_ptr(), B, Lm, dim, rows_dst.data_ptr(), 0,
                                        o32.data_ptr(), o16.data_ptr(), 1, L.stream_ptr(h.device)), "pool")
    for bi in range(B):
        n = int(lens[bi])
        e = h[bi, :n].float().sum(0) / max(n, 1e-9)
        e = e / max(float(e.norm()), 1e-12)
        torch.testing.assert_close(o32[int(rows_dst[bi])], e, atol=1e-5, rtol=1e-4)
        assert torch.equal(o16[int(rows_dst[bi])], o32[int(rows_dst[bi])].to(torch.bfloat16))


def test_full_blockmap_is_bidirectional_attention():
    """Prefill kernel under sdag_build_blockmap_full == softmax(QK^T)V over each sequence's own tokens (no causal mask),
    D = 64 (BERT geometry), ragged batch, strided q/k/v views of a fused projection."""
    import sdag_b200.sdag as sdag
    from sdag_b200.encoder import full_blockmap

    B, L, H, D = 3, 300, 4, 64
    lens = [300, 129, 5]
    g = torch.Generator(device="cuda").manual_seed(3)
    qkv = torch.randn(B, L, 3, H, D, generator=g, device="cuda").to(torch.bfloat16)
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    bm = full_blockmap(torch.tensor(lens, dtype=torch.int32, device="cuda"), L)
    assert [bm.header(i)["pairs"] for i in range(B)] == [n * n for n in lens]
    out = sdag.sdag_prefill_attention(q, k, v, bm, D ** -0.5, ragged=True)  # [B, L, H, D]
    for b, n in enumerate(lens):
        qf, kf, vf = (t[b, :, :n].double() for t in (q, k, v))
        ref = torch.softmax(qf @ kf.transpose(1, 2) * D ** -0.5, dim=-1) @ vf  # [H, n, D]
        torch.testing.assert_close(out[b, :n].permute(1, 0, 2).double(), ref, atol=2e-2, rtol=2e-2)
        assert (out[b, n:] == 0).all()


def _ranker(dtype=torch.float32):
    from sdag_b200.encoder import B200Ranker

    tok = synth.make_tokenizer()
    return B200Ranker(synth.tiny_bert(len(tok), dtype=dtype), tok), tok


def test_ranker_embeddings_match_oracle_golden(golden_dir):
    gold = np.load(os.path.join(golden_dir, "encoder.npz"))
    ranker, tok = _ranker()
    assert [len(tok(t)["input_ids"][:512]) for t in synth.ENCODER_TEXTS] == gold["lengths"].tolist()
    emb = ranker.encode(synth.ENCODER_TEXTS, convert_to_tensor=True, normalize_embeddings=True)
    assert emb.shape == gold["emb"].shape and emb.dtype == torch.float32 and emb.is_cuda
    e = emb.cpu().numpy()
    np.testing.assert_allclose(np.linalg.norm(e, axis=1), 1.0, atol=1e-5)
    cos = (e * gold["emb"]).sum(1)
    assert cos.min() > 0.999, cos
    assert np.abs(e - gold["emb"]).max() < 3e-2
    raw = ranker.encode(synth.ENCODER_TEXTS, convert_to_tensor=False, normalize_embeddings=False)
    assert isinstance(raw, np.ndarray)
    np.testing.assert_allclose(raw, gold["raw"], atol=6e-2, rtol=6e-2)
    # batching is invisible: tiny token budget (one text per batch) gives the same rows, in input order
    ranker.max_batch_tokens = 1
    e1 = ranker.encode(synth.ENCODER_TEXTS, convert_to_tensor=True, normalize_embeddings=True).cpu().numpy()
    np.testing.assert_allclose(e1, e, atol=2e-3)
    assert ranker.encode("query: hello", convert_to_tensor=True, normalize_embeddings=True).shape == (ranker.get_sentence_embedding_dimension(),)


def test_ranker_plugs_into_reference_shaped_functions(golden_dir):
    """The same object serves embed_queries / encode_texts_with_ranker / compute_doc_knn_for_docs, and encode_corpus fills
    an index whose search returns each passage for its own text as the query."""
    import sdag_b200.sdag as sdag
    from sdag_b200.encoder import encode_corpus
    from sdag_b200.retrieval import embed_queries, search_index
    from sdag_b200.selection import encode_texts_with_ranker

    ranker, _ = _ranker()
    docs = [d for d in synth.SAMPLE_DOCS if d.strip()] + synth.SAMPLE_MAL
    index = encode_corpus(ranker, docs, prefix="")
    assert index.ntotal == len(docs) and index.d == ranker.get_sentence_embedding_dimension()
    xq = embed_queries(ranker, docs, batch_size=4)
    assert xq.shape == (len(docs), index.d) and xq.dtype == np.float32
    # the same text as query finds its own row first, with inner product 1 (unit vectors, fp32 re-scoring)
    same = ranker.encode(docs, convert_to_tensor=False, normalize_embeddings=True)
    ids, scores = search_index(index, np.ascontiguousarray(same), 3)
    assert (ids[:, 0] == np.arange(len(docs))).all()
    np.testing.assert_allclose(scores[:, 0], 1.0, atol=1e-4)
    ids_q, _ = search_index(index, np.ascontiguousarray(xq), 3)
    assert ids_q.shape == (len(docs), 3)
    sel = encode_texts_with_ranker(ranker, docs[:3])
    assert sel.shape == (3, index.d)
    nbrs = sdag.compute_doc_knn_for_docs(ranker, docs[:5], 2)
    assert len(nbrs) == 5 and all(len(n) == 2 and i not in n for i, n in enumerate(nbrs))


@pytest.mark.parametrize("M,N,K", [(600, 512, 128), (1030, 264, 72)])
def test_gemm_gelu_epilogue_equals_linear_then_gelu(M, N, K):
    """llm_gemm with the GELU epilogue (BertIntermediate in one launch) vs llm_gemm -> enc_gelu: bit-identical; vs torch's
    erf GELU of the fp32 linear output: bf16 tolerance."""
    import sdag_b200._lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M)
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16)
    st = L.stream_ptr(x.device)
    fused = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    L.check(lib.llm_gemm(x.data_ptr(), K, W.data_ptr(), b.data_ptr(), fused.data_ptr(), N, M, N, K, 2, st), "gelu epilogue")
    two = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    L.check(lib.llm_gemm(x.data_ptr(), K, W.data_ptr(), b.data_ptr(), two.data_ptr(), N, M, N, K, 0, st), "plain")
    L.check(lib.enc_gelu(two.data_ptr(), two.data_ptr(), two.numel(), st), "enc_gelu")
    assert torch.equal(fused, two)
    ref = torch.nn.functional.gelu(torch.nn.functional.linear(x.float(), W.float(), b.float()).to(torch.bfloat16).float())
    torch.testing.assert_close(fused.float(), ref, atol=2e-2, rtol=2e-2)
