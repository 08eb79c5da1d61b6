"""Oracle: the SDAG allow-mask (TEST INFRASTRUCTURE).

Restates ``build_blocked_causal_mask_full`` — reference ``src/pipeline/sparse_attention_RAG/SDAG.py:68-127``
— as closed-form numpy instead of the reference's per-row slice assignments.  Regions only ever set
entries to True, so they union:

    i <  S            : j <= i                                         (SDAG.py:107-108)
    s_d <= i < e_d    : j < S  or  s_d <= j <= i  or  j in span(n), n in N(d), 0 <= n < num_docs   (113-122)
    i >= qa_start     : j <= i                                         (124-125)
    anything else     : nothing ("gap row")

Neighbours are honoured only when ``len(doc_neighbors) == len(doc_token_spans)`` (SDAG.py:111).  Slices
follow Python semantics (clipped to ``[0, L)``), which is what makes odd inputs — spans past the end,
overlapping spans, qa_start < e_d — behave exactly like the reference.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def _clip(a: int, n: int) -> int:
    """Python slice-bound normalisation for a non-negative-length axis of size n."""
    if a < 0:
        a += n
        if a < 0:
            a = 0
    return min(a, n)


def build_mask_oracle(
    base_seq_len: int,
    sys_user_len: int,
    doc_token_spans: Sequence[Tuple[int, int]],
    qa_start: int,
    doc_neighbors: Optional[List[List[int]]] = None,
) -> np.ndarray:
    """bool[L, L]; mask[i, j] True <=> token i may attend to token j."""
    L = int(base_seq_len)
    mask = np.zeros((L, L), dtype=bool)
    rows = np.arange(L)[:, None]
    cols = np.arange(L)[None, :]
    causal = cols <= rows

    # system / user block: ``for i in range(sys_user_len): mask[i, :i+1] = True``  (an IndexError in the
    # reference when sys_user_len > L; callers never do that — we clip)
    s_rows = min(max(int(sys_user_len), 0), L)
    mask[:s_rows] |= causal[:s_rows]

    num_docs = len(doc_token_spans)
    use_nbrs = doc_neighbors is not None and len(doc_neighbors) == num_docs
    S_cols = _clip(int(sys_user_len), L)  # ``mask[i, :sys_user_len]`` — slice semantics (negative wraps)
    for d, (ds, de) in enumerate(doc_token_spans):
        ds, de = int(ds), int(de)
        for i in range(max(ds, 0), min(de, L)):  # range(d_start, d_end); i must index a real row
            mask[i, :S_cols] = True
            lo = _clip(ds, L)
            mask[i, lo : i + 1] = True
            if use_nbrs:
                for n in doc_neighbors[d]:
                    if n < 0 or n >= num_docs:
                        continue
                    ns, ne = doc_token_spans[n]
                    mask[i, _clip(int(ns), L) : _clip(int(ne), L)] = True
    q0 = max(int(qa_start), 0)
    if q0 < L:
        mask[q0:] |= causal[q0:]
    return mask


def allowed_pairs(mask: np.ndarray) -> int:
    """P of SURVEY.md §8d: number of (i, j) pairs the mask allows (gap rows contribute 0)."""
    return int(mask.sum())


def gap_rows(mask: np.ndarray) -> np.ndarray:
    """Indices of rows with no allowed key (softmax over them is uniform — Appendix A2)."""
    return np.nonzero(~mask.any(axis=1))[0]


def synthetic_layout(sys_user_len: int, n_docs: int, doc_len: int, gap: int, qa_len: int):
    """The fixed-shape prompt layouts of SURVEY.md §8d (config 2: 48, 10, 256, 1, 24 -> L = 2642)."""
    spans = []
    pos = sys_user_len
    for _ in range(n_docs):
        spans.append((pos, pos + doc_len))
        pos += doc_len + gap
    qa_start = pos
    L = qa_start + qa_len
    return L, sys_user_len, spans, qa_start
