#!/usr/bin/env python
"""Input fixture for BASELINE configs 1 and 4: the reference's 1000 sampled NQ queries joined with the 5 adversarial
candidate documents per query of its attack CSV, as one compact gzip'd JSON (the GPU box has no /root/reference).

    python tests/golden/make_nq_fixture.py        # needs /root/reference (this container)

Sources (data, not code): /root/reference/data/sampled_nq_1000_queries.json (id, question, short_answers) and
/root/reference/data/poisonedRAG_attack_results_GPT4_NQ_5_mal_docs_per_query.csv, grouped exactly as the reference's
loader does (utils/parsing.py:9-97: rows grouped by `query` in first-seen order; id / query / ground truths of the first row,
distinct non-empty false answers, every non-empty malicious document) — through this repo's mirror of that loader, whose
parity with the reference function is pinned by tests/golden/harness.json.
Output: tests/golden/nq_attack_1000.json.gz = {"source": ..., "queries": [{"qid", "question", "short_answers",
"false_answers", "malicious_docs"} x 1000]} in CSV order.
"""
import gzip
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_DATA = "/root/reference/data"


def main():
    from sdag_b200.harness import load_from_csv

    csv_path = os.path.join(REF_DATA, "poisonedRAG_attack_results_GPT4_NQ_5_mal_docs_per_query.csv")
    questions, gts, false_groups, mal_groups, qids = load_from_csv(csv_path)
    with open(os.path.join(REF_DATA, "sampled_nq_1000_queries.json"), encoding="utf-8") as f:
        sampled = json.load(f)
    by_q = {s["question"].strip(): s for s in sampled}
    out = []
    for q, gt, fa, mal, qid in zip(questions, gts, false_groups, mal_groups, qids):
        s = by_q.get(q)
        out.append({"qid": s["id"] if s else qid, "question": q, "short_answers": s["short_answers"] if s else gt,
                    "false_answers": fa, "malicious_docs": mal})
    assert len(out) == 1000 and all(len(o["malicious_docs"]) == 5 for o in out), (len(out),)
    assert sum(1 for o in out if o["question"] in by_q) == 1000
    path = os.path.join(HERE, "nq_attack_1000.json.gz")
    blob = json.dumps({"source": "sampled_nq_1000_queries.json + poisonedRAG_attack_results_GPT4_NQ_5_mal_docs_per_query.csv",
                       "queries": out}, ensure_ascii=False, separators=(",", ":")).encode("utf-8")
    with open(path, "wb") as raw, gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, compresslevel=9) as f:
        f.write(blob)
    print(f"wrote {path}: {len(out)} queries, {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    main()
