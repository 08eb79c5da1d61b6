"""Oracle: exact max-inner-product search, ``faiss.IndexFlatIP`` restated (TEST INFRASTRUCTURE).

The reference calls ``index.search(q.astype(float32), top_k)`` and swaps the result to
``(indices, scores)`` — ``src/pipeline/retrieval/dense.py:68-84``.  The arithmetic is in ``faiss-cpu``
(un-vendored, version not pinned: README.md:78; not installable here), so this file restates the
published behaviour of ``IndexFlat`` with ``METRIC_INNER_PRODUCT``:

* scores are exact fp32 inner products (blocked ``sgemm``: 4096-query x 1024-row blocks);
* per query a size-k min-heap; a candidate replaces the root only when strictly greater, so among equal
  scores at the cut the lower index survives; results come back sorted by descending score (ties: ascending
  id);
* fewer than k vectors: the tail is padded with id -1 and score ``-FLT_MAX`` (what ``dense.py:118-125``
  then maps to ``""`` / ``"NA"``).

**parity unpinned** for the faiss arithmetic itself (no faiss here, no golden vectors in the reference);
pinned at the reference's call sites by running ``dense.py`` unmodified on top of this class
(``tests/golden/make_golden.py``).  fp32 summation order differs between BLAS builds, so scores are
compared to 1e-5 relative and ids exactly except for ties / near-ties inside that envelope.
"""
from __future__ import annotations

import numpy as np
import torch

NEG_FLT_MAX = -3.4028234663852886e38


class FlatIPOracle:
    """faiss-shaped: ``.d``, ``.ntotal``, ``.add(x)``, ``.search(x, k) -> (D float32[Q,k], I int64[Q,k])``."""

    def __init__(self, d: int):
        self.d = int(d)
        self._chunks = []
        self.ntotal = 0
        self.tie_pad = 16

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d
        self._chunks.append(torch.from_numpy(x))
        self.ntotal += x.shape[0]

    def _corpus(self) -> torch.Tensor:
        if len(self._chunks) > 1:
            self._chunks = [torch.cat(self._chunks, dim=0)]
        return self._chunks[0] if self._chunks else torch.zeros((0, self.d), dtype=torch.float32)

    def search(self, x, k: int, block_rows: int = 65536):
        xq = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        Q = xq.shape[0]
        k = int(k)
        xb = self._corpus()
        N = xb.shape[0]
        D = np.full((Q, k), NEG_FLT_MAX, dtype=np.float32)
        I = np.full((Q, k), -1, dtype=np.int64)
        if N == 0 or Q == 0 or k == 0:
            return D, I
        cand_s, cand_i = [], []
        for r0 in range(0, N, block_rows):
            blk = xb[r0 : r0 + block_rows]
            s = xq @ blk.T  # fp32 sgemm
            # keep a few extra per block so that a group of equal scores straddling the cut is still fully
            # present when the (score desc, id asc) order is applied below (tie groups > tie_pad: not handled)
            kk = min(k + self.tie_pad, blk.shape[0])
            vals, idx = torch.topk(s, kk, dim=1, sorted=True)
            cand_s.append(vals.numpy())
            cand_i.append((idx + r0).numpy())
        cs = np.concatenate(cand_s, axis=1)
        ci = np.concatenate(cand_i, axis=1)
        # order: score descending, id ascending
        order = np.lexsort((ci, -cs.astype(np.float64)), axis=1)
        take = min(k, cs.shape[1], N)
        rows = np.arange(Q)[:, None]
        D[:, :take] = cs[rows, order[:, :take]]
        I[:, :take] = ci[rows, order[:, :take]]
        bad = ~np.isfinite(D)
        D[bad] = NEG_FLT_MAX
        I[bad] = -1
        return D, I


def merge_topk_oracle(scores: np.ndarray, ids: np.ndarray, k: int):
    """Merge per-shard results ``[G, Q, k]`` into the global top-k with the same order rule
    (score descending, id ascending; id -1 padding last)."""
    G, Q, kk = scores.shape
    s = np.transpose(scores, (1, 0, 2)).reshape(Q, G * kk)
    i = np.transpose(ids, (1, 0, 2)).reshape(Q, G * kk)
    key_id = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((key_id, -s.astype(np.float64)), axis=1)[:, :k]
    rows = np.arange(Q)[:, None]
    return s[rows, order], i[rows, order]
