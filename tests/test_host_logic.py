"""CPU tests: host-side logic of the package against the reference goldens, ABI symbol coverage, loud failure
without a GPU.  No kernel is launched here."""
import ctypes
import json
import os
import random
import re

import numpy as np
import pytest


def test_library_exports_every_declared_symbol():
    import sdag_b200._lib as L

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "sdag_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|size_t|const char\*)\s+(\w+)\s*\(", header, flags=re.M))
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    # the ctypes table and the header agree on the number of parameters of every entry point
    for name, params in re.findall(r"^\s*(?:int|size_t|const char\*)\s+(\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.M | re.S):
        n_params = 0 if params.strip() in ("", "void") else params.count(",") + 1
        assert n_params == len(L.SIGNATURES[name][1]), (name, n_params, len(L.SIGNATURES[name][1]))
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.sdag_b200_version() >= 100
    assert lib.sdag_b200_error_string(1001) == b"invalid argument"
    # pure size queries work without a GPU
    assert lib.sdag_layout_ints(10, 2) == 8 + 20 + 20
    assert lib.sdag_blockmap_bytes(2642, 10, 0) % 16 == 0 and lib.sdag_blockmap_bytes(2642, 10, 0) > 0
    assert lib.sdag_prefill_workspace_bytes(1, 8, 128) == 4 * 8 * 33 * 128 + 16
    assert lib.ip_search_workspace_bytes(4096, 1024, 10, 2048) > 4096 * 2048 * 8


def test_ranked_list_injection_matches_reference_golden(golden_dir):
    from sdag_b200.ranked_list import inject_malicious_docs_into_ranked_list

    with open(os.path.join(golden_dir, "ranked_list.json")) as f:
        recs = json.load(f)
    for r in recs:
        random.seed(r["seed"])
        assert inject_malicious_docs_into_ranked_list(r["base"], r["malicious"], r["pos"]) == r["out"], r


def test_prompt_text_is_the_reference_text(golden_dir):
    """chat_sha256 in the golden file was computed by the reference over its own prompt constants."""
    import hashlib

    import synth
    from sdag_b200.prompts import SYSTEM_PROMPT_RAG, USER_RAG_PROMPT
    from sdag_b200.sdag import build_rag_chat_and_spans

    tok = synth.make_tokenizer()
    with open(os.path.join(golden_dir, "generation.json")) as f:
        recs = json.load(f)
    for r in recs:
        chat, S, spans, qa, ranked = build_rag_chat_and_spans(tok, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT, synth.SAMPLE_QUERY,
                                                              r["malicious"], r["retrieved"], r["rank"])
        assert hashlib.sha256(chat.encode()).hexdigest() == r["chat_sha256"]
        assert [list(s) for s in spans] == r["spans"] and S == r["S"] and qa == r["qa_start"]
        assert ranked == r["ranked_docs"]


def test_materialize_and_search_index_contract(golden_dir):
    """search_index swaps (scores, ids) -> (ids, scores); materialisation of out-of-range ids (dense.py:118-125)."""
    from oracle.flat_ip import FlatIPOracle
    from sdag_b200.retrieval import DenseRetriever, materialize_dense_hits, search_index

    import synth

    with open(os.path.join(golden_dir, "retrieval.json")) as f:
        blob = json.load(f)
    emb = np.stack([synth.text_embedding("passage: " + t) for t in blob["corpus"]])
    for rec in blob["records"]:
        n = rec["n_rows"]
        idx = FlatIPOracle(emb.shape[1])  # faiss-shaped stand-in: the host logic is index-agnostic
        idx.add(emb[:n])
        rb = DenseRetriever(synth.HashRanker(), idx, blob["meta"][:n]).retrieve_batch(rec["queries"], rec["k"], 2)
        assert rb.ids_full == rec["ids_full"] and rb.docs_texts_full == rec["docs_texts_full"]
        np.testing.assert_allclose(np.array(rb.scores_full), np.array(rec["scores_full"]), rtol=1e-6)
    ids, sc = search_index(idx, emb[:2], 3)
    assert ids.dtype == np.int64 and sc.dtype == np.float32
    t, i, s = materialize_dense_hits([[0, -1, 99]], [[1.0, 2.0, 3.0]], [{"id": 7, "text": "x"}])
    assert t == [["x", "", ""]] and i == [["7", "NA", "NA"]] and s == [[1.0, 2.0, 3.0]]


def test_selector_random_paths_use_global_rng_like_reference(golden_dir):
    from sdag_b200.selection import select_malicious_docs_for_batch

    import synth

    with open(os.path.join(golden_dir, "selector.json")) as f:
        blob = json.load(f)
    for rec in blob["records"]:
        if rec["strategy"] != "random":
            continue  # non-"random" strategies need the GPU when retrieved docs are non-empty
        random.seed(rec["seed"])
        got = select_malicious_docs_for_batch(synth.HashRanker(), blob["retrieved"], blob["candidates"], "random", rec["max_docs"])
        assert got == rec["chosen"]


def test_faiss_flat_file_round_trip(tmp_path):
    import struct

    from sdag_b200.retrieval import load_faiss_index_and_meta, write_flat_ip_file

    x = np.random.default_rng(0).standard_normal((5, 8)).astype(np.float32)
    p = tmp_path / "faiss.index"
    write_flat_ip_file(str(p), x)
    raw = p.read_bytes()
    assert raw[:4] == b"IxFI" and struct.unpack("<i", raw[4:8])[0] == 8 and len(raw) == 4 + 4 + 24 + 1 + 4 + 8 + 5 * 8 * 4
    with pytest.raises(FileNotFoundError):
        load_faiss_index_and_meta(str(p), str(tmp_path / "missing.jsonl"))
    with pytest.raises(FileNotFoundError):
        load_faiss_index_and_meta(str(tmp_path / "missing.index"), str(p))


def test_faiss_flat_ip_reader_on_hand_assembled_fixture(golden_dir, tmp_path):
    """tests/golden/faiss_flat_ip_3x64.index was assembled field by field from faiss' serializer (make_ixfi_fixture.py);
    the header below is that layout spelled out in hex, independent of both the fixture script and the reader."""
    import struct

    from sdag_b200.retrieval import parse_flat_ip_file, write_flat_ip_file

    path = os.path.join(golden_dir, "faiss_flat_ip_3x64.index")
    raw = open(path, "rb").read()
    header = bytes.fromhex(
        "49784649"          # 0   fourcc "IxFI"  (IndexFlat, METRIC_INNER_PRODUCT)
        "40000000"          # 4   d = 64                        int32
        "0300000000000000"  # 8   ntotal = 3                    int64 (idx_t)
        "0000100000000000"  # 16  dummy = 1 << 20               int64
        "0000100000000000"  # 24  dummy = 1 << 20               int64
        "01"                # 32  is_trained = true             1 byte
        "00000000"          # 33  metric_type = 0 (inner product) int32
        "c000000000000000"  # 37  size = ntotal * d = 192 float32 values   uint64 (WRITEXBVECTOR)
    )
    assert len(header) == 45 and raw[:45] == header and len(raw) == 45 + 4 * 192
    r = np.arange(3, dtype=np.float32)[:, None]
    c = np.arange(64, dtype=np.float32)[None, :]
    want = ((r + 1) * np.float32(0.5) - c / np.float32(64)).astype(np.float32)
    x = parse_flat_ip_file(path)
    assert x.shape == (3, 64) and x.dtype == np.float32 and np.array_equal(np.asarray(x), want)
    # this repo's writer produces the same bytes
    mine = tmp_path / "mine.index"
    write_flat_ip_file(str(mine), want)
    assert mine.read_bytes() == raw
    # the metadata file next to it: one {"id", "text"} object per line (dense.py:38-41)
    meta = [json.loads(line) for line in open(os.path.join(golden_dir, "faiss_flat_ip_3x64.docs_meta.jsonl"), encoding="utf-8")]
    assert [m["id"] for m in meta] == ["doc-0", "doc-1", "doc-2"] and all("text" in m for m in meta)
    # other index types, metrics and damaged files are refused with a reason
    for fourcc, word in ((b"IxF2", "IndexFlatL2"), (b"IwFl", "fourcc")):
        bad = tmp_path / "bad.index"
        bad.write_bytes(fourcc + raw[4:])
        with pytest.raises(ValueError, match=word):
            parse_flat_ip_file(str(bad))
    bad = tmp_path / "metric.index"
    bad.write_bytes(raw[:33] + struct.pack("<i", 1) + raw[37:])
    with pytest.raises(ValueError, match="metric_type"):
        parse_flat_ip_file(str(bad))
    bad.write_bytes(raw[:-8])
    with pytest.raises(ValueError, match="truncated"):
        parse_flat_ip_file(str(bad))
    bad.write_bytes(raw[:37] + struct.pack("<Q", 191) + raw[45:])
    with pytest.raises(ValueError, match="corrupt"):
        parse_flat_ip_file(str(bad))


def test_product_path_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sdag_b200.sdag as sdag
    from sdag_b200._lib import SdagB200Error

    with pytest.raises(SdagB200Error):
        sdag.run_rag_with_doc_isolation(model=None, tokenizer=None, query="q", malicious_docs=[], retrieved_docs=[],
                                        max_new_tokens=1, device="cpu", add_attack_in_rank=0, ranker=None, doc_neighbors_k=0)
    with pytest.raises(SdagB200Error):
        sdag.generate_with_custom_mask(None, None, torch.zeros(1, 3, dtype=torch.long), torch.ones(3, 3, dtype=torch.bool))


def test_package_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "sparse-document-attention-rag-_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_hybrid_rrf_matches_reference_golden(golden_dir):
    """RRF fusion vs the reference's retrieval/hybrid.py run on the same lists.  The reference iterates a ``set`` of keys,
    so the order among exactly equal fused scores is hash-dependent there: scores must match bit for bit, and every
    group of equal scores must hold the same documents."""
    import random

    from sdag_b200.hybrid import fuse_sparse_and_dense_batch, split_k_between_sparse_and_dense

    with open(os.path.join(golden_dir, "hybrid.json")) as f:
        gold = json.load(f)
    for c in gold["cases"]:
        t, i, s = fuse_sparse_and_dense_batch(c["sparse_texts"], c["sparse_ids"], c["dense_texts"], c["dense_ids"],
                                              top_k=c["top_k"], seed=c["seed"], k0=c["k0"])
        assert s == c["fused_scores"]
        for q in range(len(s)):
            full = sorted(set(s[q]))
            # a tie group cut by top_k may keep different members: compare only groups strictly above the last score
            for val in full:
                mine = sorted((t[q][j], i[q][j]) for j in range(len(s[q])) if s[q][j] == val)
                ref = sorted((c["fused_texts"][q][j], c["fused_ids"][q][j]) for j in range(len(s[q])) if s[q][j] == val)
                if val != s[q][-1]:
                    assert mine == ref
                else:
                    assert len(mine) == len(ref)
    rng = random.Random(gold["split_seed"])
    for rec in gold["splits"]:
        assert list(split_k_between_sparse_and_dense(rec["top_k"], rng)) == rec["split"]


def test_hybrid_retriever_plugs_two_arms_together():
    from sdag_b200.datamodels import RetrievalBatch
    from sdag_b200.hybrid import HybridRetriever
    from sdag_b200.retrieval import Retriever

    class Arm(Retriever):
        def __init__(self, tag, docs):
            self.tag, self.docs = tag, docs

        def retrieve_batch(self, queries, max_k_needed, embed_batch_size):
            n = len(queries)
            return RetrievalBatch(q_embs=[self.tag] * n, docs_texts_full=[list(self.docs[:max_k_needed])] * n,
                                  ids_full=[[f"{d}-id" for d in self.docs[:max_k_needed]]] * n,
                                  scores_full=[[1.0] * min(max_k_needed, len(self.docs))] * n)

    hy = HybridRetriever(Arm("dense", ["a", "b", "c", "d"]), Arm("sparse", ["c", "x", "a", "y"]), seed=42)
    rb = hy.retrieve_batch(["q1", "q2"], 4, 32)
    assert rb.q_embs == ["dense", "dense"]
    assert rb.docs_texts_full[0][0] in ("a", "c") and len(rb.docs_texts_full[0]) <= 4
    assert rb.ids_full[0][0].endswith("-id") and rb.scores_full[0] == sorted(rb.scores_full[0], reverse=True)


def test_apply_ranked_list_order_matches_reference_golden(golden_dir):
    import random

    from sdag_b200.ranked_list import apply_ranked_list_order

    with open(os.path.join(golden_dir, "hybrid.json")) as f:
        gold = json.load(f)
    for rec in gold["orders"]:
        random.seed(rec["seed"])
        docs = list(rec["docs"])
        out = apply_ranked_list_order(docs, rec["mode"])
        assert out == rec["out"]
        if rec["mode"] in ("top_down", "other"):
            assert out is docs


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU port timed on the host cores) must print ONE JSON line with the contract keys."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--no-secondary"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "tok/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
