"""Types crossing the drop-in boundary (shape-compatible with src/pipeline/models/datamodels.py:30-41)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List


@dataclass
class RetrievalBatch:
    """Per query, ranked lists of length ``max_k_needed``: texts, ids (str), scores (float)."""

    q_embs: List[Any]
    docs_texts_full: List[List[str]]
    ids_full: List[List[str]]
    scores_full: List[List[float]]
