"""Evaluation harness (sdag_b200.harness) against outputs of the reference's own utilities on seeded synthetic rows
(tests/golden/harness.json, written by tests/golden/make_golden.py: utils/normalization.py, metrics.py, parsing.py,
save_results.py).  Host logic only."""
import contextlib
import csv
import io
import json
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "harness.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def inputs():
    # the input generator lives next to the golden script; importing make_golden itself would need /root/reference
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"), encoding="utf-8").read()
    start = src.index("def harness_inputs():")
    end = src.index("def gen_harness():")
    ns = {"random": __import__("random")}
    exec(src[start:end], ns)
    return ns["harness_inputs"]()


def test_answer_scoring_matches_reference(gold, inputs):
    from sdag_b200 import harness as h

    _, texts, _ = inputs
    assert [h.normalize_answer(t) for t in texts if t is not None] == gold["normalize"]
    assert [h.extract_final_answer(t) for t in texts] == gold["extract"]
    assert [[h.exact_match(a, b) for b in texts] for a in texts] == gold["exact"]
    assert [[h.exact_match(a, b) for b in texts if b is not None] for a in texts if a is not None] == gold["exact_norm_module"]


def test_attack_csv_loader_matches_reference(gold, inputs, tmp_path):
    from sdag_b200 import harness as h

    _, _, csv_rows = inputs
    cells = [r[2] for r in csv_rows] + [None, "  ", "[1, 2]", "{'a': 1}"]
    assert [h.parse_list_field(c) for c in cells] == gold["parse_list"]
    path = tmp_path / "attack.csv"
    with open(path, "w", encoding="utf-8", newline="") as f:
        w = csv.writer(f)
        w.writerow(["query", "query_id", "ground_truth_answers", "false_answer", "malicious_document"])
        w.writerows(csv_rows)
    assert [list(x) for x in h.load_from_csv(str(path))] == gold["load_from_csv"]
    assert [list(x) for x in h.load_from_csv(str(path), match_field_for_groups="query_id")] == gold["load_from_csv_by_id"]
    bad = tmp_path / "bad.csv"
    bad.write_text("query,query_id\nq,1\n")
    with pytest.raises(ValueError, match="missing required columns"):
        h.load_from_csv(str(bad))


def test_result_csv_is_byte_identical_to_reference(gold, inputs, tmp_path):
    from sdag_b200 import harness as h

    rows, _, _ = inputs
    full = [dict(r, short_answers=h.ensure_list(r["short_answers"]),
                 false_answer=r["false_answer"] if isinstance(r["false_answer"], str) else " / ".join(r["false_answer"]),
                 retrieved_docs=r["retrieved_docs"] or []) for r in rows]
    out = tmp_path / "sub" / "res.csv"
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        h.save_results(full, str(out))
    assert open(out, encoding="utf-8", newline="").read() == gold["save_results_csv"]
    assert buf.getvalue() == gold["save_results_stdout"]


def test_pair_metrics_json_matches_reference(gold, inputs):
    from sdag_b200 import harness as h

    rows, _, _ = inputs

    def same(a, b):  # same JSON text: keys, key order, values
        return json.dumps(a, ensure_ascii=False) == json.dumps(b, ensure_ascii=False)

    for k in (1, 3, 5):
        assert same(h.compute_retrieval_ground_truth_stats(rows, k), gold["retrieval_stats"][str(k)]), k
    assert same(h.compute_retrieval_ground_truth_stats([], 4), gold["retrieval_stats_empty"])
    assert same(h.compute_answer_overlap_and_attack_stats(rows), gold["overlap"])
    assert same(h.compute_answer_overlap_and_attack_stats([]), gold["overlap_empty"])
    assert same(h.build_pair_metrics(rows, 5, 2), gold["pair_metrics"])
    assert same(h.build_pair_metrics([], 3, 0), gold["pair_metrics_empty"])
    assert same(h.compute_false_answer_stats_for_results(rows), gold["false_answer_stats"])


def test_pair_specs_and_row_scoring():
    from sdag_b200 import harness as h

    assert h.build_pair_specs([5, 10, 3], [1, -1]) == [h.PairSpec(5, 1), h.PairSpec(10, -1)]
    assert h.compute_max_k_needed([h.PairSpec(5, 1), h.PairSpec(10, 0)]) == 10
    assert h.compute_max_k_needed([h.PairSpec(5, 1)], "doc_corruption") == 6
    assert h.attack_config_requests_docs(0) is False and h.attack_config_requests_docs([0, 2]) and not h.attack_config_requests_docs("x")
    row = h.score_answers("q1", "who?", ["Gustave Eiffel"], ["Rome"], ["d a", "d b"], ["1", "2"], ["evil"], 1,
                          "<think>..</think>Answer: gustave EIFFEL.\nmore", "It was Rome")
    assert row["rag_answer_iso"] == "gustave EIFFEL." and row["ground_truth_match_iso"] and not row["false_match_iso"]
    assert row["false_match_noiso"] and not row["ground_truth_match_noiso"] and row["malicious_doc"] == "evil"
    assert h.score_answers("q", "w", ["x"], [], [], [], ["evil"], 0, "x", "y")["malicious_doc"] == ""


def test_main_loop_partitions_queries_across_ranks(tmp_path, monkeypatch):
    """harness.main with ``query_slice`` (the query-parallel multi-GPU split, no collective): two ranks' CSV rows together are
    the single-rank run's rows; retrieval happens once per batch at the deepest k; every pair gets its own files."""
    import csv as _csv

    from sdag_b200 import config, harness as h
    from sdag_b200.datamodels import RetrievalBatch
    import sdag_b200.selection as selection

    nq = 11
    queries = [f"what is item {i}?" for i in range(nq)]
    calls = []

    class FakeRetriever:
        def retrieve_batch(self, qs, max_k_needed, embed_batch_size):
            calls.append((tuple(qs), max_k_needed, embed_batch_size))
            docs = [[f"{q} doc{j}" + (" answer" + q.split()[3] if j == 1 else "") for j in range(max_k_needed)] for q in qs]
            ids = [[f"{q}-{j}" for j in range(max_k_needed)] for q in qs]
            return RetrievalBatch(q_embs=[None] * len(qs), docs_texts_full=docs, ids_full=ids,
                                  scores_full=[[1.0 / (j + 1) for j in range(max_k_needed)] for _ in qs])

    def fake_pair(*, pair, batch_qids, batch_qs, batch_gts, retrieval_docs_full, retrieval_ids_full, false_groups_batch,
                  chosen_mals_batch, **kw):
        return [h.score_answers(qid, q, gts, fas, docs[: pair.top_k + 1], ids[: pair.top_k + 1], mals, pair.attacker_pos,
                                f"Answer: {gts[0]}" if int(qid[1:]) % 2 else "unknown", f"{fas[0]} it is")
                for qid, q, gts, fas, docs, ids, mals in zip(batch_qids, batch_qs, batch_gts, false_groups_batch, retrieval_docs_full,
                                                             retrieval_ids_full, chosen_mals_batch)]

    monkeypatch.setattr(h, "run_pair_malicious_doc_attack_for_batch", fake_pair)
    monkeypatch.setattr(selection, "select_malicious_docs_for_batch",
                        lambda ranker, retrieved_docs_batch_full, malicious_doc_groups_batch, strategy, max_docs: [g[:max_docs] for g in malicious_doc_groups_batch])
    monkeypatch.setattr(config, "TOP_K", [3, 5])
    monkeypatch.setattr(config, "ADD_ATTACK_IN_RANK", [1, 0])
    monkeypatch.setattr(config, "BATCH_SIZE_EMBED_Q", 4)
    common = dict(retriever=FakeRetriever(), llm_model=None, tokenizer=None, ranker=None, queries=queries,
                  short_answers=[[f"answer{i}?"] for i in range(nq)], query_ids=[f"q{i}" for i in range(nq)],
                  false_answer_groups=[["wrong"]] * nq, malicious_doc_groups=[[f"mal{i}a", f"mal{i}b"] for i in range(nq)])
    full = h.main(output_base=str(tmp_path / "all"), **common)
    assert set(full) == {(3, 1), (5, 0)} and all(c[1] == 5 and len(c[0]) <= 4 for c in calls) and len(calls) == 3
    parts = [h.main(output_base=str(tmp_path / f"rank{r}"), query_slice=range(r, nq, 2), **common) for r in range(2)]

    def rows(stem, k, pos):
        return list(_csv.DictReader(open(f"{stem}_top_k={k}_attacker_pos={pos}.csv", encoding="utf-8", newline="")))

    for k, pos in full:
        whole = {r["query_id"]: r for r in rows(tmp_path / "all", k, pos)}
        split = {r["query_id"]: r for rank in range(2) for r in rows(tmp_path / f"rank{rank}", k, pos)}
        assert split == whole and len(whole) == nq
        assert sum(p[(k, pos)]["num_queries"] for p in parts) == nq
        iso = full[(k, pos)]["answer_match_stats"]["iso"]
        assert iso["ground_truth_match_count"] == nq // 2 and full[(k, pos)]["answer_match_stats"]["no_iso"]["false_answer_match_count"] == nq
        assert sum(p[(k, pos)]["answer_match_stats"]["iso"]["ground_truth_match_count"] for p in parts) == nq // 2
        # the second retrieved document holds the ground truth -> every query sits in the "exactly 1 relevant document" bucket at rank 2
        bucket = full[(k, pos)]["retrieval_ground_truth_stats"]["per_k_exact_match_buckets"]["1"]
        assert bucket["queries_with_exactly_k_ground_truth_docs_count"] == nq
        assert list(bucket["single_relevant_doc_rank_distribution"]) == ["2"]
    assert h.main(output_base=str(tmp_path / "none"), query_slice=[], **common) == {}


def test_json_config_and_query_set(tmp_path, monkeypatch, inputs):
    import csv as _csv

    from sdag_b200 import config, harness as h

    monkeypatch.setattr(config, "TOP_K", [5])
    monkeypatch.setattr(config, "DOC_NEIGHBORS_K", 0)
    cfg = tmp_path / "run.json"
    cfg.write_text(json.dumps({"TOP_K": [3, 10], "DOC_NEIGHBORS_K": 2, "NOT_A_KNOB": 1}))
    h.apply_json_config(h.load_json_config(str(cfg)))
    assert config.TOP_K == [3, 10] and config.DOC_NEIGHBORS_K == 2 and not hasattr(config, "NOT_A_KNOB")
    assert h.load_json_config(str(tmp_path / "missing.json")) == {} and h.load_json_config(None) == {}
    (tmp_path / "broken.json").write_text("{not json")
    assert h.load_json_config(str(tmp_path / "broken.json")) == {}
    h.apply_json_config({})
    _, _, csv_rows = inputs
    path = tmp_path / "attack.csv"
    with open(path, "w", encoding="utf-8", newline="") as f:
        w = _csv.writer(f)
        w.writerow(["query", "query_id", "ground_truth_answers", "false_answer", "malicious_document"])
        w.writerows(csv_rows)
    qd = h.load_queries_unified(str(path))
    assert qd.questions == ["who built it", "when", "where", "what", "why"] and qd.query_ids == ["7", "8", "9", "10", "11"]
    assert qd.false_answer_groups[0] == ["Rome", "London"] and qd.malicious_doc_groups[1] == ["doc three", "doc seven"]
    assert h.export_config_dict()["TOP_K"] == [3, 10]
