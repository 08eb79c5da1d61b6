"""Hybrid retrieval: reciprocal-rank fusion of a dense and a sparse ranked list (host logic).

Mirrors ``src/pipeline/retrieval/hybrid.py``: ``split_k_between_sparse_and_dense`` (:10-27), ``rrf_fuse_one_query``
(:30-107), ``fuse_sparse_and_dense_batch`` (:110-165) and ``HybridRetriever`` (:168-225).  The dense arm is this
package's ``DenseRetriever`` (CUDA search); the sparse arm is whatever ``Retriever`` the caller plugs in (the
reference's BM25 arm needs a JVM, SURVEY.md §2 — out of scope here).

One documented difference: the reference collects the fused keys in a Python ``set``, so documents whose fused
scores are exactly equal come out in hash order, which changes from process to process.  Here ties keep first-seen
order (sparse list first, then dense) — one of the orders the reference can produce.
"""
from __future__ import annotations

import random
from typing import Dict, List, Sequence, Tuple

from .datamodels import RetrievalBatch
from .retrieval import Retriever


def split_k_between_sparse_and_dense(top_k: int, rng: random.Random) -> Tuple[int, int]:
    """(k_sparse, k_dense) summing to ``top_k``; an odd ``top_k`` gives the extra slot to one arm by one ``rng.random()``
    draw (< 0.5: sparse) — even values consume no randomness."""
    half, odd = divmod(top_k, 2)
    if not odd:
        return half, half
    return (half + 1, half) if rng.random() < 0.5 else (half, half + 1)


def _fusion_key(doc_id, text: str) -> str:
    return doc_id if doc_id not in (None, "", "NA") else text


def rrf_fuse_one_query(sparse_docs: List[str], sparse_ids: List[str], dense_docs: List[str], dense_ids: List[str],
                       k0: int = 60) -> Tuple[List[str], List[str], List[float]]:
    """score(doc) = sum over the lists holding it of 1 / (k0 + rank), rank 1-indexed; documents are identified by id,
    or by text when the id is missing / "" / "NA".  A key seen twice in one list keeps its LAST rank there (dict
    overwrite in the reference) while text and id come from its first appearance, sparse list first."""
    score_parts: Dict[str, List[float]] = {}
    first_seen: Dict[str, Tuple[str, str]] = {}
    for arm, (docs, ids) in enumerate(((sparse_docs, sparse_ids), (dense_docs, dense_ids))):
        for rank, (text, doc_id) in enumerate(zip(docs, ids), start=1):
            key = _fusion_key(doc_id, text)
            first_seen.setdefault(key, (text, doc_id))
            score_parts.setdefault(key, [0.0, 0.0])[arm] = 1.0 / float(k0 + rank)
    # sparse term added first, like the reference's accumulation order (bitwise-equal float sums)
    fused = [(0.0 + parts[0] + parts[1], key) for key, parts in score_parts.items()]
    fused.sort(key=lambda sk: sk[0], reverse=True)  # stable: ties keep first-seen order
    return ([first_seen[k][0] for _, k in fused], [first_seen[k][1] for _, k in fused], [s for s, _ in fused])


def fuse_sparse_and_dense_batch(sparse_texts: List[List[str]], sparse_ids: List[List[str]], dense_texts: List[List[str]],
                                dense_ids: List[List[str]], top_k: int, seed: int, k0: int = 60):
    """Per query: split the budget, cut both lists to their share, fuse, keep ``top_k``.  One ``random.Random(seed)`` is
    shared by the batch, consumed in query order."""
    rng = random.Random(seed)
    out_texts, out_ids, out_scores = [], [], []
    for s_t, s_i, d_t, d_i in zip(sparse_texts, sparse_ids, dense_texts, dense_ids):
        k_s, k_d = split_k_between_sparse_and_dense(top_k, rng)
        texts, ids, scores = rrf_fuse_one_query(s_t[:k_s], s_i[:k_s], d_t[:k_d], d_i[:k_d], k0=k0)
        out_texts.append(texts[:top_k])
        out_ids.append(ids[:top_k])
        out_scores.append(scores[:top_k])
    return out_texts, out_ids, out_scores


class HybridRetriever(Retriever):
    """``RETRIEVER_BACKEND = "sparse_and_dense"``: both arms retrieve ``max_k_needed``, the fused list replaces the
    texts / ids / scores, the query embeddings stay the dense arm's."""

    def __init__(self, dense_retriever: Retriever, sparse_retriever: Retriever, seed: int, k0: int = 60) -> None:
        self.dense, self.sparse, self.seed, self.k0 = dense_retriever, sparse_retriever, seed, k0

    def retrieve_batch(self, queries: Sequence[str], max_k_needed: int, embed_batch_size: int) -> RetrievalBatch:
        dense = self.dense.retrieve_batch(queries, max_k_needed, embed_batch_size)
        sparse = self.sparse.retrieve_batch(queries, max_k_needed, embed_batch_size)
        texts, ids, scores = fuse_sparse_and_dense_batch(sparse.docs_texts_full, sparse.ids_full, dense.docs_texts_full,
                                                         dense.ids_full, top_k=max_k_needed, seed=self.seed, k0=self.k0)
        return RetrievalBatch(q_embs=dense.q_embs, docs_texts_full=texts, ids_full=ids, scores_full=scores)
