"""Oracle: centroid-distance selector and doc-kNN on embeddings (TEST INFRASTRUCTURE).

Restates, on already-computed embeddings, the numeric core of

* ``_select_malicious_docs_for_query`` — reference ``src/pipeline/attack/malicious_selection.py:37-137``:
  centroid = plain mean of the retrieved-doc embeddings, NOT re-normalised (:89, :117);
  ``sims = cand @ centroid.T`` (:95, :123); near = ``argmax`` / ``argsort(-sims)``, far = ``argmin`` /
  ``argsort(sims)`` (:97-100, :125-128); ``np.argmax/argmin`` return the first occurrence on ties.
* ``compute_doc_knn_for_docs`` — reference ``src/pipeline/sparse_attention_RAG/SDAG.py:14-65``:
  ``sims = emb @ emb.T`` fp32, ``argsort(-row)``, self removed by index, first k kept.

The encoder forward (sentence-transformers) is outside these functions: tests feed embeddings directly.
"""
from __future__ import annotations

from typing import List

import numpy as np


def centroid_sims_oracle(retrieved_emb: np.ndarray, cand_emb: np.ndarray) -> np.ndarray:
    centroid = np.asarray(retrieved_emb, dtype=np.float32).mean(axis=0, keepdims=True)
    return np.matmul(np.asarray(cand_emb, dtype=np.float32), centroid.T).reshape(-1)


def select_order_oracle(retrieved_emb: np.ndarray, cand_emb: np.ndarray, strategy: str, target_n: int) -> List[int]:
    """Indices of the chosen candidates, in the order the reference returns them."""
    sims = centroid_sims_oracle(retrieved_emb, cand_emb)
    if target_n == 1:
        if strategy == "closest_to_centroid":
            return [int(np.argmax(sims))]
        if strategy == "furthest_from_centroid":
            return [int(np.argmin(sims))]
        raise ValueError("random strategies are not numeric")
    if strategy == "closest_to_centroid":
        order = np.argsort(-sims)
    elif strategy == "furthest_from_centroid":
        order = np.argsort(sims)
    else:
        raise ValueError("random strategies are not numeric")
    return [int(i) for i in order[:target_n]]


def doc_knn_oracle(emb: np.ndarray, k_neighbors: int) -> List[List[int]]:
    """kNN lists among n unit vectors (all docs non-empty)."""
    emb = np.asarray(emb, dtype=np.float32)
    n = emb.shape[0]
    if k_neighbors <= 0 or n <= 1:
        return [[] for _ in range(n)]
    sims = np.matmul(emb, emb.T)
    out: List[List[int]] = []
    for r in range(n):
        order = np.argsort(-sims[r])
        knn = [int(j) for j in order if j != r][:k_neighbors]
        out.append(knn)
    return out
