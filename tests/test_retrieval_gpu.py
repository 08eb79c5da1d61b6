"""GPU parity: exact inner-product top-k, merge, selector and doc-kNN through the C ABI vs the oracle."""
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle.flat_ip import FlatIPOracle, merge_topk_oracle
from oracle.selector import centroid_sims_oracle, doc_knn_oracle, select_order_oracle

import synth

pytestmark = pytest.mark.gpu


def _unit_rows(n, d, seed, device="cuda"):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n, d, generator=g, device=device, dtype=torch.float32)
    return x / x.norm(dim=1, keepdim=True)


def _check_against_oracle(x, q, k, ids, scores, rtol=1e-5):
    """ids bit-exact except where the oracle's own fp32 scores are tied / within the fp32 summation envelope."""
    oracle = FlatIPOracle(x.shape[1])
    oracle.add(x)
    D, I = oracle.search(q, k)
    np.testing.assert_allclose(scores, D, rtol=rtol, atol=1e-6)
    mism = ids != I
    if mism.any():
        # documented exception class: near-ties (|ds| <= 2e-6 relative) may swap
        exact = (q.astype(np.float64) @ x.astype(np.float64).T)
        for r, c in zip(*np.nonzero(mism)):
            a, b = ids[r, c], I[r, c]
            assert a >= 0 and b >= 0
            assert abs(exact[r, a] - exact[r, b]) <= 2e-6 * max(1.0, abs(exact[r, b])), (r, c, a, b)
    return int(mism.sum())


@pytest.mark.parametrize("N,d,Q,k", [(5000, 128, 37, 10), (1300, 64, 130, 11), (70000, 256, 64, 51)])
def test_flat_ip_search_matches_oracle(N, d, Q, k):
    from sdag_b200.retrieval import FlatIPIndex

    x = _unit_rows(N, d, 42)
    q = _unit_rows(Q, d, 7)
    idx = FlatIPIndex(d)
    idx.add(x.cpu().numpy())
    assert idx.ntotal == N and idx.d == d
    D, I = idx.search(q.cpu().numpy(), k)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.shape == (Q, k)
    assert (np.diff(D, axis=1) <= 0).all()
    _check_against_oracle(x.cpu().numpy(), q.cpu().numpy(), k, I, D)
    assert int(idx.last_stats[0]) == 0  # no overflow fallback on well-spread data


@pytest.mark.parametrize("k", [10, 11, 51])
def test_flat_ip_search_cta_pair_kernel_d1024_with_tie_set(k):
    """The configuration bench.py times, in small: d = 1024 (16 k-blocks through the 6-stage ring of
    ip_gemm_filter_kernel_2cta), a 640-query batch (five query tiles: three CTA-pair tiles, one of them half empty) over
    200 k rows, 1 % of the rows duplicated at known indices (SURVEY.md §8d tie stress set), against the independent fp32
    oracle.  k = TOP_K, TOP_K + 1 and the config-5 size."""
    from sdag_b200.retrieval import FlatIPIndex

    N, d, Q = 200_000, 1024, 640
    x = _unit_rows(N, d, 42)
    rng = np.random.default_rng(1)
    src = rng.choice(N, N // 100, replace=False)
    dst = rng.choice(np.setdiff1d(np.arange(N), src), N // 100, replace=False)
    x[torch.from_numpy(dst).cuda()] = x[torch.from_numpy(src).cuda()]
    q = _unit_rows(Q, d, 7)
    q[:64] = x[torch.from_numpy(src[:64]).cuda()]  # these queries' two best hits are an exactly tied pair
    idx = FlatIPIndex.from_device(x)
    s, i = idx.search_device(q, k)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    assert int(idx.last_stats[0]) == 0  # candidate lists did not overflow: the tensor-core path produced the result
    xh, qh = x.cpu().numpy(), q.cpu().numpy()
    oracle = FlatIPOracle(d)
    oracle.add(xh)
    D, I = oracle.search(qh, k)
    np.testing.assert_allclose(s, D, rtol=1e-5, atol=1e-6)
    assert (np.diff(s, axis=1) <= 0).all()
    for r, c in zip(*np.nonzero(i != I)):  # documented exception: fp32 near-ties (<= 2e-6 relative) may swap
        a, b = i[r, c], I[r, c]
        ea, eb = float(qh[r].astype(np.float64) @ xh[a].astype(np.float64)), float(qh[r].astype(np.float64) @ xh[b].astype(np.float64))
        assert a >= 0 and b >= 0 and abs(ea - eb) <= 2e-6 * max(1.0, abs(eb)), (r, c, a, b, ea, eb)
    tied = 0
    for r in range(64):  # exact ties: lower id first (faiss heap rule)
        assert s[r, 0] == s[r, 1] and i[r, 0] < i[r, 1] and {int(i[r, 0]), int(i[r, 1])} == {int(src[r]), int(dst[r])}
        tied += 1
    assert tied == 64


def test_flat_ip_search_unnormalised_rows_and_incremental_add():
    from sdag_b200.retrieval import FlatIPIndex

    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(9000, 128, generator=g, device="cuda") * torch.rand(9000, 1, generator=g, device="cuda") * 3
    q = torch.randn(20, 128, generator=g, device="cuda")
    idx = FlatIPIndex(128)
    idx.add(x[:4000].cpu().numpy())
    idx.add(x[4000:])
    D, I = idx.search(q.cpu().numpy(), 10)
    _check_against_oracle(x.cpu().numpy(), q.cpu().numpy(), 10, I, D)


def test_flat_ip_search_ties_lower_id_first():
    """1 % duplicated rows at known indices: exactly equal scores keep the lower id first (faiss heap rule)."""
    from sdag_b200.retrieval import FlatIPIndex

    x = _unit_rows(20000, 128, 42)
    rng = np.random.default_rng(0)
    src = rng.choice(20000, 200, replace=False)
    dst = rng.choice(20000, 200, replace=False)
    x[torch.from_numpy(dst).cuda()] = x[torch.from_numpy(src).cuda()]
    q = x[torch.from_numpy(src[:50]).cuda()].clone()  # each query's best hits are a duplicated pair
    idx = FlatIPIndex.from_device(x)
    s, i = idx.search_device(q, 10)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    exact_dupes = 0
    for r in range(s.shape[0]):
        for c in range(9):
            if s[r, c] == s[r, c + 1]:
                assert i[r, c] < i[r, c + 1]
                exact_dupes += 1
    assert exact_dupes >= 40
    _check_against_oracle(x.cpu().numpy(), q.cpu().numpy(), 10, i, s)


def test_flat_ip_search_fewer_rows_than_k_pads():
    from sdag_b200.retrieval import FlatIPIndex, NEG_FLT_MAX

    x = _unit_rows(3, 64, 1)
    q = _unit_rows(2, 64, 2)
    idx = FlatIPIndex(64)
    idx.add(x.cpu().numpy())
    D, I = idx.search(q.cpu().numpy(), 5)
    assert (I[:, 3:] == -1).all() and (D[:, 3:] == np.float32(NEG_FLT_MAX)).all()
    oracle = FlatIPOracle(64)
    oracle.add(x.cpu().numpy())
    Do, Io = oracle.search(q.cpu().numpy(), 5)
    assert np.array_equal(I, Io)
    np.testing.assert_allclose(D[:, :3], Do[:, :3], rtol=1e-5)
    empty = FlatIPIndex(64)
    De, Ie = empty.search(q.cpu().numpy(), 4)
    assert (Ie == -1).all()


def test_flat_ip_search_bf16_corpus_is_exact_over_rounded_rows():
    from sdag_b200.retrieval import FlatIPIndex

    x = _unit_rows(30000, 128, 5).to(torch.bfloat16)
    q = _unit_rows(33, 128, 6)
    idx = FlatIPIndex.from_device(x, corpus_dtype="bf16")
    s, i = idx.search_device(q, 10)
    _check_against_oracle(x.float().cpu().numpy(), q.cpu().numpy(), 10, i.cpu().numpy(), s.cpu().numpy())


def test_flat_ip_search_clustered_corpus_overflow_fallback_still_exact():
    """Thousands of near-identical rows defeat the threshold filter: the counted exact-scan fallback must kick in
    and the answer must still be the exact one."""
    from sdag_b200.retrieval import FlatIPIndex

    g = torch.Generator(device="cuda").manual_seed(9)
    base = _unit_rows(1, 64, 11)
    x = base + 1e-4 * torch.randn(6000, 64, generator=g, device="cuda")
    x = torch.cat([x, _unit_rows(4000, 64, 12)])
    q = torch.cat([base, _unit_rows(3, 64, 13)])
    idx = FlatIPIndex.from_device(x.contiguous(), cand_cap=512)
    s, i = idx.search_device(q, 10)
    assert int(idx.last_stats[0]) >= 1
    _check_against_oracle(x.cpu().numpy(), q.cpu().numpy(), 10, i.cpu().numpy(), s.cpu().numpy())


def test_flat_ip_round_trip_property_at_scale():
    """Size-independent property at a size the oracle cannot finish quickly: every corpus row queried against the
    index must return itself first with score |x|^2 (unit rows -> 1), and results must be sorted."""
    from sdag_b200.retrieval import FlatIPIndex

    N, d = 1_000_000, 256
    x = _unit_rows(N, d, 42)
    idx = FlatIPIndex.from_device(x)
    rows = torch.arange(0, N, N // 1024, device="cuda")[:1024]
    s, i = idx.search_device(x[rows], 10)
    assert torch.equal(i[:, 0], rows)
    torch.testing.assert_close(s[:, 0], torch.ones_like(s[:, 0]), atol=1e-5, rtol=1e-5)
    assert (s[:, 1:] <= s[:, :-1]).all()
    # spot-check 16 queries fully against a brute-force fp32 matmul
    ref = x[rows[:16]] @ x.T
    rs, ri = ref.topk(10, dim=1)
    torch.testing.assert_close(s[:16], rs, atol=2e-6, rtol=1e-5)
    assert (ri == i[:16]).float().mean() > 0.98


def test_topk_merge_matches_oracle_and_sharded_equals_unsharded():
    from sdag_b200.retrieval import FlatIPIndex, merge_topk_device

    N, d, Q, k, G = 40000, 128, 50, 10, 4
    x = _unit_rows(N, d, 42)
    x[30000] = x[5]  # a cross-shard exact tie
    q = torch.cat([_unit_rows(Q - 1, d, 7), x[5:6]])
    full_s, full_i = FlatIPIndex.from_device(x).search_device(q, k)
    per = N // G
    ss, ii = [], []
    for g in range(G):
        shard = FlatIPIndex.from_device(x[g * per : (g + 1) * per].contiguous(), id_offset=g * per)
        s, i = shard.search_device(q, k)
        ss.append(s)
        ii.append(i)
    gs, gi = torch.stack(ss), torch.stack(ii)
    ms, mi = merge_topk_device(gs, gi, k)
    assert torch.equal(mi, full_i) and torch.equal(ms, full_s)
    os_, oi = merge_topk_oracle(gs.cpu().numpy(), gi.cpu().numpy(), k)
    assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(ms.cpu().numpy(), os_)
    # padding (-1) sorts last
    gi2 = gi.clone()
    gs2 = gs.clone()
    gi2[0, :, 5:] = -1
    gs2[0, :, 5:] = -3.4028234663852886e38
    ms2, mi2 = merge_topk_device(gs2, gi2, k)
    os2, oi2 = merge_topk_oracle(gs2.cpu().numpy(), gi2.cpu().numpy(), k)
    assert np.array_equal(mi2.cpu().numpy(), oi2)


def test_dense_retriever_reproduces_reference_call_site_golden(golden_dir):
    """DenseRetriever.retrieve_batch over FlatIPIndex == the reference's dense.py run on the oracle index."""
    from sdag_b200.retrieval import DenseRetriever, FlatIPIndex

    with open(os.path.join(golden_dir, "retrieval.json")) as f:
        blob = json.load(f)
    emb = np.stack([synth.text_embedding("passage: " + t) for t in blob["corpus"]])
    for rec in blob["records"]:
        n = rec["n_rows"]
        idx = FlatIPIndex(emb.shape[1])
        idx.add(emb[:n])
        retr = DenseRetriever(synth.HashRanker(), idx, blob["meta"][:n])
        rb = retr.retrieve_batch(rec["queries"], max_k_needed=rec["k"], embed_batch_size=2)
        assert rb.ids_full == rec["ids_full"]
        assert rb.docs_texts_full == rec["docs_texts_full"]
        np.testing.assert_allclose(np.array(rb.scores_full), np.array(rec["scores_full"]), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.stack(rb.q_embs), np.array(rec["q_embs"]), rtol=0, atol=0)


def test_selector_kernel_matches_oracle_and_golden(golden_dir):
    from sdag_b200.selection import centroid_order_device, select_malicious_docs_for_batch

    rng = np.random.default_rng(0)
    nq, R, C, d = 9, 11, 5, 1024
    r = rng.standard_normal((nq, R, d)).astype(np.float32)
    c = rng.standard_normal((nq, C, d)).astype(np.float32)
    r /= np.linalg.norm(r, axis=2, keepdims=True)
    c /= np.linalg.norm(c, axis=2, keepdims=True)
    r_cnt = np.array([11, 1, 5, 11, 7, 2, 11, 3, 9], dtype=np.int32)
    c_cnt = np.array([5, 5, 3, 1, 5, 2, 4, 5, 5], dtype=np.int32)
    c[0, 3] = c[0, 1]  # exact tie -> lower index first
    for furthest in (False, True):
        sims, order = centroid_order_device(torch.from_numpy(r).cuda(), torch.from_numpy(r_cnt).cuda(),
                                            torch.from_numpy(c).cuda(), torch.from_numpy(c_cnt).cuda(), furthest)
        sims, order = sims.cpu().numpy(), order.cpu().numpy()
        for i in range(nq):
            ref = centroid_sims_oracle(r[i, : r_cnt[i]], c[i, : c_cnt[i]])
            np.testing.assert_allclose(sims[i, : c_cnt[i]], ref, rtol=1e-5, atol=1e-6)
            key = ref if furthest else -ref
            expect = np.argsort(key, kind="stable")
            assert list(order[i, : c_cnt[i]]) == list(expect), (i, furthest)
            assert (order[i, c_cnt[i]:] == -1).all()

    with open(os.path.join(golden_dir, "selector.json")) as f:
        blob = json.load(f)
    for rec in blob["records"]:
        random.seed(rec.get("seed", 1234))
        got = select_malicious_docs_for_batch(synth.HashRanker(), blob["retrieved"], blob["candidates"], rec["strategy"],
                                              rec["max_docs"])
        assert got == rec["chosen"], rec


def test_doc_knn_kernel_matches_oracle_and_reference_golden(golden_dir):
    import sdag_b200.sdag as sdag

    rng = np.random.default_rng(42)
    emb = rng.standard_normal((50, 64)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    for k in (1, 2, 7):
        got = sdag.doc_knn_device(torch.from_numpy(emb).cuda(), k).cpu().numpy()
        assert [list(r) for r in got] == doc_knn_oracle(emb, k)
    with open(os.path.join(golden_dir, "knn.json")) as f:
        recs = json.load(f)
    for r in recs:
        assert sdag.compute_doc_knn_for_docs(synth.HashRanker(), r["docs"], r["k"]) == r["neighbors"]


def test_selector_reuses_index_embeddings_by_row():
    """select_malicious_docs_for_batch with RetrievalBatch.rows_full + the index reads the retrieved documents'
    embeddings from the index (no second encode of those texts) and chooses exactly what the re-encoding path chooses."""
    import synth
    from sdag_b200.retrieval import DenseRetriever, FlatIPIndex
    from sdag_b200.selection import select_malicious_docs_for_batch

    ranker = synth.HashRanker()
    docs = [f"document number {i} about topic {i % 7}" for i in range(200)]
    emb = ranker.encode(["passage: " + d for d in docs], convert_to_tensor=True, normalize_embeddings=True)
    index = FlatIPIndex.from_device(emb.cuda().float())
    meta = [{"id": f"d{i}", "text": d} for i, d in enumerate(docs)]
    rb = DenseRetriever(ranker, index, meta).retrieve_batch(["what about topic 3?", "and topic 5?", "x"], 6, 32)
    assert rb.rows_full is not None and [len(r) for r in rb.rows_full] == [6, 6, 6]
    assert all(meta[r]["text"] == t for rows, texts in zip(rb.rows_full, rb.docs_texts_full) for r, t in zip(rows, texts))
    cands = [[f"fake {q} {j}" for j in range(5)] for q in range(3)]
    for strategy in ("closest_to_centroid", "furthest_from_centroid"):
        ranker.calls.clear()
        a = select_malicious_docs_for_batch(ranker, rb.docs_texts_full, cands, strategy, 2)
        calls_encode = len([c for c in ranker.calls if any(t.startswith("passage: document") for t in c)])
        ranker.calls.clear()
        b = select_malicious_docs_for_batch(ranker, rb.docs_texts_full, cands, strategy, 2,
                                            retrieved_rows_batch_full=rb.rows_full, index=index)
        calls_reuse = len([c for c in ranker.calls if any(t.startswith("passage: document") for t in c)])
        assert a == b
        assert calls_encode == 1 and calls_reuse == 0


def test_load_faiss_index_and_meta_on_hand_assembled_fixture(golden_dir):
    """The byte-documented IndexFlatIP fixture (tests/golden/make_ixfi_fixture.py) through the reference's loader name
    (dense.py:15-41) and a search: ids / scores of the three stored vectors."""
    from sdag_b200.retrieval import DenseRetriever, load_faiss_index_and_meta, materialize_dense_hits, search_index

    index, meta = load_faiss_index_and_meta(os.path.join(golden_dir, "faiss_flat_ip_3x64.index"),
                                            os.path.join(golden_dir, "faiss_flat_ip_3x64.docs_meta.jsonl"))
    assert index.ntotal == 3 and index.d == 64 and len(meta) == 3
    r = np.arange(3, dtype=np.float32)[:, None]
    c = np.arange(64, dtype=np.float32)[None, :]
    x = ((r + 1) * np.float32(0.5) - c / np.float32(64)).astype(np.float32)
    q = np.stack([x[2], -x[0]]).astype(np.float32)
    idx, sc = search_index(index, q, 4)
    want = q.astype(np.float64) @ x.astype(np.float64).T
    order = np.argsort(-want, axis=1, kind="stable")
    assert np.array_equal(idx[:, :3], order) and (idx[:, 3] == -1).all()
    np.testing.assert_allclose(sc[:, :3], np.take_along_axis(want, order, axis=1), rtol=1e-5)
    texts, ids, _ = materialize_dense_hits(idx, sc, meta)
    assert ids[0][0] == "doc-2" and ids[0][3] == "NA" and texts[0][3] == ""
    assert isinstance(DenseRetriever(None, index, meta), DenseRetriever)


def test_flat_ip_search_with_early_rescoring_knob_in_subprocess():
    """SDAG_IP_EARLY_RESCORE=1 (read once per process; off by default because it measured slower): the candidates found
    before the last epoch are re-scored on a second stream under the last epoch's GEMM and tagged in place — ids and scores
    must still equal the oracle's, exact ties included."""
    import subprocess
    import sys

    code = r'''
import numpy as np, torch, sys
sys.path.insert(0, "tests")
from oracle.flat_ip import FlatIPOracle
from sdag_b200.retrieval import FlatIPIndex
g = torch.Generator(device="cuda").manual_seed(5)
x = torch.randn(300000, 256, generator=g, device="cuda"); x /= x.norm(dim=1, keepdim=True)
x[150000:150100] = x[:100]
q = torch.randn(300, 256, generator=g, device="cuda"); q /= q.norm(dim=1, keepdim=True)
q[:50] = x[:50]
idx = FlatIPIndex.from_device(x)
for k in (10, 51):
    s, i = idx.search_device(q, k)
    o = FlatIPOracle(256); o.add(x.cpu().numpy())
    D, I = o.search(q.cpu().numpy(), k)
    np.testing.assert_allclose(s.cpu().numpy(), D, rtol=1e-5, atol=1e-6)
    got = i.cpu().numpy()
    x64, q64 = x.double(), q.double()
    for r, c in zip(*np.nonzero(got != I)):  # the documented exception class: fp32 near-ties (<= 2e-6) may swap
        a, b = int(got[r, c]), int(I[r, c])
        assert a >= 0 and b >= 0 and abs(float(q64[r] @ x64[a]) - float(q64[r] @ x64[b])) <= 2e-6, (k, r, c, a, b)
    assert (got[:50, 0] == np.arange(50)).all()  # the exactly tied pairs: lower id first
    print("k", k, "near-tie swaps", int((got != I).sum()))
    assert int(idx.last_stats[0]) == 0
print("early rescore ok")
'''
    env = dict(os.environ, SDAG_IP_EARLY_RESCORE="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "early rescore ok" in r.stdout, r.stdout + r.stderr
