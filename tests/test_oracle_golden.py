"""The oracle is only trusted once it reproduces the reference's own outputs (tests/golden/, made by
tests/golden/make_golden.py importing /root/reference).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.attention import masked_attention_oracle
from oracle.sdag_mask import allowed_pairs, build_mask_oracle, gap_rows
from oracle.selector import doc_knn_oracle, select_order_oracle

import synth


def _mask_meta(golden_dir):
    with open(os.path.join(golden_dir, "masks.json")) as f:
        return json.load(f)


def test_mask_oracle_matches_reference_bits(golden_dir):
    packed = np.load(os.path.join(golden_dir, "masks.npz"))
    for e in _mask_meta(golden_dir):
        if e["L"] > 4096:
            continue
        m = build_mask_oracle(e["L"], e["S"], e["spans"], e["qa_start"], e["neighbors"])
        assert allowed_pairs(m) == e["pairs"], e["name"]
        assert len(gap_rows(m)) == e["gap_rows"], e["name"]
        bits = np.packbits(m, axis=1)
        assert hashlib.sha256(bits.tobytes()).hexdigest() == e["sha256"], e["name"]
        if e["name"] in packed.files:
            assert np.array_equal(bits, packed[e["name"]]), e["name"]


def test_cpu_reference_mask_equals_reference_builder(golden_dir):
    """The mask builder of the CPU reference arm (oracle/cpu_reference.py): both its restatement and — where
    /root/reference exists — the imported reference function give the reference's bits on every strict golden layout."""
    from oracle.cpu_reference import reference_mask_builder, reference_style_mask

    builders = [reference_style_mask, reference_mask_builder()[0]]
    checked = 0
    for e in _mask_meta(golden_dir):
        if e["L"] > 4096 or e["neighbors"] is not None:
            continue
        canonical = e["S"] >= 0 and all(a < b for a, b in e["spans"]) and e["qa_start"] >= 0
        if not canonical:
            continue
        for build in builders:
            m = build(e["L"], e["S"], [tuple(s) for s in e["spans"]], e["qa_start"]).numpy()
            if build is reference_style_mask and hashlib.sha256(np.packbits(m, axis=1).tobytes()).hexdigest() != e["sha256"]:
                # the restatement covers the layouts the prompt builder produces (disjoint ascending spans)
                spans = e["spans"]
                ordered = all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
                assert not ordered or not spans or spans[0][0] < e["S"] or spans[-1][1] > e["qa_start"], e["name"]
                continue
            assert hashlib.sha256(np.packbits(m, axis=1).tobytes()).hexdigest() == e["sha256"], e["name"]
            checked += 1
    assert checked >= 4
    by = {e["name"]: e for e in _mask_meta(golden_dir)}
    e = by["cfg2_k10x256"]  # the benchmark layout itself
    m = reference_style_mask(e["L"], e["S"], [tuple(s) for s in e["spans"]], e["qa_start"]).numpy()
    assert int(m.sum()) == 516148 and hashlib.sha256(np.packbits(m, axis=1).tobytes()).hexdigest() == e["sha256"]


def test_mask_oracle_pair_counts_match_survey(golden_dir):
    by = {e["name"]: e for e in _mask_meta(golden_dir)}
    assert by["cfg1_k5x256"]["pairs"] == 259388
    assert by["cfg2_k10x256"]["pairs"] == 516148
    assert by["cfg5_k50x640_strict"]["pairs"] == 12563828
    assert by["cfg5_k50x640_K2"]["pairs"] == 53523828


def test_attention_oracle_matches_hf_eager(golden_dir):
    g = np.load(os.path.join(golden_dir, "attention.npz"))
    by = {e["name"]: e for e in _mask_meta(golden_dir)}
    names = sorted({f.split("/")[0] for f in g.files})
    assert len(names) >= 8
    for name in names:
        e = by[name]
        m = build_mask_oracle(e["L"], e["S"], e["spans"], e["qa_start"], e["neighbors"])
        q, k, v, ref = g[name + "/q"], g[name + "/k"], g[name + "/v"], g[name + "/out"]
        out = masked_attention_oracle(q, k, v, m, q.shape[-1] ** -0.5)
        np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-5, err_msg=name)
        # gap rows are the mean of V over every key (Appendix A2)
        for r in gap_rows(m):
            np.testing.assert_allclose(out[r, 0], v[0].mean(axis=0), rtol=1e-6, atol=1e-6)


def test_knn_oracle_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "knn.json")) as f:
        recs = json.load(f)
    for r in recs:
        docs, k = r["docs"], r["k"]
        nonempty = [i for i, d in enumerate(docs) if d and d.strip()]
        expect = [[] for _ in docs]
        if k > 0 and len(nonempty) > 1:
            emb = np.stack([synth.text_embedding("passage: " + docs[i]) for i in nonempty])
            local = doc_knn_oracle(emb, k)
            for row, i in enumerate(nonempty):
                expect[i] = [nonempty[j] for j in local[row]]
        assert expect == r["neighbors"]


def test_selector_oracle_matches_reference(golden_dir):
    with open(os.path.join(golden_dir, "selector.json")) as f:
        blob = json.load(f)
    for rec in blob["records"]:
        if rec["strategy"] not in ("closest_to_centroid", "furthest_from_centroid"):
            continue
        for retrieved, cands, chosen in zip(blob["retrieved"], blob["candidates"], rec["chosen"]):
            nonempty = [d for d in retrieved if d and d.strip()]
            if not cands:
                assert chosen == []
                continue
            if not nonempty:
                continue  # reference falls back to the global RNG (S6)
            md = rec["max_docs"]
            target = len(cands) if (md is None or md < 0 or md >= len(cands)) else md
            r_emb = np.stack([synth.text_embedding("passage: " + d) for d in nonempty])
            c_emb = np.stack([synth.text_embedding("passage: " + d) for d in cands])
            order = select_order_oracle(r_emb, c_emb, rec["strategy"], target)
            assert [cands[i] for i in order] == chosen


def test_encoder_oracle_reproduces_golden(golden_dir):
    """oracle/encoder.py against the committed HF-BertModel fixture (same seeded tiny BERT, fp32, eager)."""
    import os

    import numpy as np

    import synth
    from oracle.encoder import encode_oracle

    gold = np.load(os.path.join(golden_dir, "encoder.npz"))
    tok = synth.make_tokenizer()
    bert = synth.tiny_bert(len(tok))
    toks = [tok(t)["input_ids"][:512] for t in synth.ENCODER_TEXTS]
    assert [len(t) for t in toks] == gold["lengths"].tolist()
    np.testing.assert_allclose(encode_oracle(bert, toks, pad_id=tok.pad_token_id), gold["emb"], atol=1e-5)
    np.testing.assert_allclose(encode_oracle(bert, toks[::-1], pad_id=tok.pad_token_id, batch=2)[::-1], gold["emb"], atol=1e-5)
