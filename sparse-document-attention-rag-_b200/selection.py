"""Near / far centroid-distance selection of adversarial documents (hot path B, selector) — host side.

Mirrors ``src/pipeline/attack/malicious_selection.py``: ``encode_texts_with_ranker`` (:12-34),
``_select_malicious_docs_for_query`` (:37-137) and ``select_malicious_docs_for_batch`` (:140-180), with the
same strategies (``"random"``, ``"closest_to_centroid"``, ``"furthest_from_centroid"``, anything else ->
random), the same ``max_docs`` rule (None / negative / >= len -> all candidates) and the same fall-backs to
the global ``random`` module (no retrieved docs, empty candidate list -> ``[]``).

What changes is where the arithmetic runs: the reference loops over queries and per query encodes, copies the
embeddings to the host and does a numpy matvec + argsort.  Here every text of the batch is encoded with two
``ranker.encode`` calls (one for retrieved docs, one for candidates), the embeddings stay on the GPU and one
``centroid_select`` launch scores and orders all queries.
"""
from __future__ import annotations

import random
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import config as _config

RANKER_MODEL_NAME = _config.RANKER_MODEL_NAME  # patchable (malicious_selection.py:9)


def _prefixed(texts: Sequence[str]) -> List[str]:
    use_prefix = "e5" in RANKER_MODEL_NAME.lower()
    return [("passage: " + t) if use_prefix else t for t in texts]


def encode_texts_with_ranker(ranker, texts: List[str]) -> np.ndarray:
    """Unit-norm float32 embeddings ``[n, dim]`` (``[0, dim]`` for an empty list) — malicious_selection.py:12-34."""
    if not texts:
        return np.zeros((0, ranker.get_sentence_embedding_dimension()), dtype=np.float32)
    emb = ranker.encode(_prefixed(texts), convert_to_tensor=True, normalize_embeddings=True)
    return emb.cpu().numpy().astype(np.float32)


def _encode_device(ranker, texts: List[str], device) -> torch.Tensor:
    emb = ranker.encode(_prefixed(texts), convert_to_tensor=True, normalize_embeddings=True)
    return torch.as_tensor(emb).to(device=device, dtype=torch.float32)


def centroid_order_device(retrieved: torch.Tensor, r_count: torch.Tensor, cand: torch.Tensor, c_count: torch.Tensor,
                          furthest: bool):
    """``retrieved [nq,R,d]``, ``cand [nq,C,d]`` (fp32, device) + valid counts -> (sims ``[nq,C]``, order ``[nq,C]``)."""
    lib = _lib.load()
    nq, R, d = retrieved.shape
    C = cand.shape[1]
    sims = torch.empty((nq, C), dtype=torch.float32, device=retrieved.device)
    order = torch.empty((nq, C), dtype=torch.int32, device=retrieved.device)
    _lib.check(
        lib.centroid_select(retrieved.contiguous().data_ptr(), r_count.data_ptr(), R, cand.contiguous().data_ptr(),
                            c_count.data_ptr(), C, nq, d, 1 if furthest else 0, sims.data_ptr(), order.data_ptr(),
                            _lib.stream_ptr(retrieved.device)),
        "centroid_select",
    )
    return sims, order


def _target_n(n_candidates: int, max_docs: Optional[int]) -> int:
    if max_docs is None or max_docs < 0 or max_docs >= n_candidates:
        return n_candidates
    return max_docs


def select_malicious_docs_for_batch(
    ranker,
    retrieved_docs_batch_full: List[List[str]],
    malicious_doc_groups_batch: List[List[str]],
    strategy: str,
    max_docs: int,
    device="cuda",
    retrieved_rows_batch_full: Optional[List[List[int]]] = None,
    index=None,
) -> List[List[str]]:
    """Same signature and results as the reference function.  Extension: with ``retrieved_rows_batch_full`` (the
    ``RetrievalBatch.rows_full`` of this package's ``DenseRetriever``) and the ``index`` they refer to, the retrieved
    documents' embeddings are read from the index instead of being encoded again — the reference re-encodes them for
    every query (malicious_selection.py:85,113) although they were embedded with the same ``"passage: "`` prefix when the
    index was built.  Rows that are missing (-1) fall back to encoding."""
    numeric = strategy in ("closest_to_centroid", "furthest_from_centroid")
    pairs = list(zip(retrieved_docs_batch_full, malicious_doc_groups_batch))

    # which queries take the numeric path (same tests, in the same order, as the reference's per-query function)
    plan = []
    for retrieved, cands in pairs:
        if not cands:
            plan.append(("empty", None))
            continue
        n = _target_n(len(cands), max_docs)
        nonempty = [d for d in retrieved if d and d.strip()]
        if not numeric or not nonempty:
            plan.append(("random", n))
        else:
            plan.append(("numeric", (n, nonempty)))
    reuse = retrieved_rows_batch_full is not None and index is not None and hasattr(index, "reconstruct_rows")

    order_rows = {}
    numeric_ids = [i for i, (kind, _) in enumerate(plan) if kind == "numeric"]
    if numeric_ids:
        r_lists = [plan[i][1][1] for i in numeric_ids]
        c_lists = [list(pairs[i][1]) for i in numeric_ids]
        r_flat = [t for lst in r_lists for t in lst]
        c_flat = [t for lst in c_lists for t in lst]
        r_emb = None
        if reuse:
            rows_flat = []
            for i in numeric_ids:
                keep = [j for j, t in enumerate(pairs[i][0]) if t and t.strip()]
                rows_i = list(retrieved_rows_batch_full[i])
                rows_flat += [rows_i[j] if j < len(rows_i) else -1 for j in keep]
            if rows_flat and min(rows_flat) >= 0:
                r_emb = index.reconstruct_rows(torch.tensor(rows_flat, dtype=torch.int64)).to(device)
        if r_emb is None:
            r_emb = _encode_device(ranker, r_flat, device)
        c_emb = _encode_device(ranker, c_flat, device)
        d = r_emb.shape[1]
        R = max(len(l) for l in r_lists)
        C = max(len(l) for l in c_lists)
        nq = len(numeric_ids)
        r_pad = torch.zeros((nq, R, d), dtype=torch.float32, device=device)
        c_pad = torch.zeros((nq, C, d), dtype=torch.float32, device=device)
        ro = co = 0
        for j, (rl, cl) in enumerate(zip(r_lists, c_lists)):
            r_pad[j, : len(rl)] = r_emb[ro : ro + len(rl)]
            c_pad[j, : len(cl)] = c_emb[co : co + len(cl)]
            ro += len(rl)
            co += len(cl)
        r_cnt = torch.tensor([len(l) for l in r_lists], dtype=torch.int32, device=device)
        c_cnt = torch.tensor([len(l) for l in c_lists], dtype=torch.int32, device=device)
        _, order = centroid_order_device(r_pad, r_cnt, c_pad, c_cnt, furthest=(strategy == "furthest_from_centroid"))
        order_host = order.cpu().numpy()
        for j, i in enumerate(numeric_ids):
            order_rows[i] = order_host[j]

    chosen_batch: List[List[str]] = []
    for i, ((retrieved, cands), (kind, info)) in enumerate(zip(pairs, plan)):
        if kind == "empty":
            chosen_batch.append([])
        elif kind == "random":
            n = info
            # same draws from the global RNG as the reference (malicious_selection.py:78-79, 84, 106-107, 111)
            chosen_batch.append([random.choice(cands)] if n == 1 else random.sample(cands, n))
        else:
            n = info[0]
            chosen_batch.append([cands[int(c)] for c in order_rows[i][:n]])
    return chosen_batch
