"""Types crossing the drop-in boundary (shape-compatible with src/pipeline/models/datamodels.py:30-41)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Optional


@dataclass
class RetrievalBatch:
    """Per query, ranked lists of length ``max_k_needed``: texts, ids (str), scores (float)."""

    q_embs: List[Any]
    docs_texts_full: List[List[str]]
    ids_full: List[List[str]]
    scores_full: List[List[float]]
    # extension (not in the reference type): index row of every hit (-1 when outside the metadata), so that later stages
    # can read the stored embeddings instead of re-encoding the texts (selection.select_malicious_docs_for_batch)
    rows_full: Optional[List[List[int]]] = None
