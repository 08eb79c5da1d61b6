"""Dense retrieval (hot path B) — host side.

Mirrors ``src/pipeline/retrieval/dense.py`` / ``retriever.py`` of the reference:

========================================  ==========================================================
reference                                 here
========================================  ==========================================================
``Retriever`` ABC        retriever.py:9   ``Retriever`` (same abstract method)
``load_faiss_index_and_meta``  dense:15   same name; reads a faiss ``IndexFlatIP`` file into HBM
``embed_queries``              dense:44   same name / semantics ("query: " prefix iff e5)
``search_index``               dense:68   same name; returns ``(indices, scores)``
``materialize_dense_hits``     dense:87   same name; out-of-range index -> ``""`` / ``"NA"``
``DenseRetriever``             dense:133  same name (``B200DenseRetriever`` is an alias)
``faiss.IndexFlatIP``          (3rd pty)  ``FlatIPIndex`` — ``.d .ntotal .add(x) .search(x, k)``
========================================  ==========================================================

``FlatIPIndex`` keeps the corpus resident in HBM (fp32 rows for exact re-scoring + a bf16 shadow for the
tensor-core pass) and runs ``ip_search`` from ``libsdag_b200.so``; ``ShardedFlatIPIndex`` partitions the
rows over the ranks of a ``torch.distributed`` group and merges local results after one all-gather.
"""
from __future__ import annotations

import json
import os
import struct
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import config as _config
from .datamodels import RetrievalBatch

RANKER_MODEL_NAME = _config.RANKER_MODEL_NAME  # patchable, like the reference's import-time copy (dense.py:10)
NEG_FLT_MAX = -3.4028234663852886e38
DEFAULT_CAND_CAP = 2048
MAX_QUERY_BATCH = 4096


class Retriever(ABC):
    """Retriever plug point (reference ``retrieval/retriever.py:9-19``)."""

    @abstractmethod
    def retrieve_batch(self, queries: Sequence[str], max_k_needed: int, embed_batch_size: int) -> RetrievalBatch:
        raise NotImplementedError


_FAST_GROWTH_ROWS = int(os.environ.get("SDAG_IP_FAST_GROWTH_ROWS", "0")) or 65536  # mirrors ip_search.cu (launch accounting only)


def _n_epochs(n_rows: int, cand_cap: int) -> int:
    """Number of (GEMM + tighten) epochs ip_search runs: the first cap/2 rows are dumped, then the swept prefix grows x4
    per epoch below 64 k rows and x2 above (ip_search.cu, IP_FAST_GROWTH_ROWS)."""
    span = max(256, cand_cap // 2 // 256 * 256)
    done, n = 0, 0
    while done < n_rows:
        done = min(n_rows, done + span)
        span = 3 * done if done < _FAST_GROWTH_ROWS else done
        n += 1
    return n


class FlatIPIndex:
    """Exact inner-product index with the faiss ``IndexFlatIP`` surface the reference touches
    (``dense.py:36, 83``): ``d``, ``ntotal``, ``add``, ``search``.

    ``corpus_dtype="bf16"`` stores bf16 rows only; search is then exact over ``float32(bf16(corpus))``.
    """

    MAX_ROWS = 0x7FFFFF00  # rows per shard (int32 row numbers inside ip_search); k <= cand_cap / 32; d % 64 == 0, d <= 8192

    def __init__(self, d: int, device="cuda", corpus_dtype: str = "fp32", id_offset: int = 0,
                 cand_cap: int = DEFAULT_CAND_CAP):
        if corpus_dtype not in ("fp32", "bf16"):
            raise ValueError("corpus_dtype must be 'fp32' or 'bf16'")
        self.d = int(d)
        self.device = torch.device(device)
        self.corpus_dtype = corpus_dtype
        self.id_offset = int(id_offset)
        self.cand_cap = int(cand_cap)
        self.ntotal = 0
        self._f32: Optional[torch.Tensor] = None   # [cap_rows, d] fp32 (None for bf16 corpora)
        self._bf16: Optional[torch.Tensor] = None  # [cap_rows, d] bf16 shadow
        self._max_norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._ws: Optional[torch.Tensor] = None
        self.last_stats: Optional[torch.Tensor] = None

    # -- construction ---------------------------------------------------------------------------------
    @classmethod
    def from_device(cls, corpus: torch.Tensor, **kw) -> "FlatIPIndex":
        """Adopt a device-resident ``[N, d]`` tensor without copying it (fp32, or bf16 with corpus_dtype='bf16')."""
        idx = cls(corpus.shape[1], device=corpus.device, **kw)
        idx._adopt(corpus)
        return idx

    @classmethod
    def from_encoded(cls, rows_f32: torch.Tensor, rows_bf16: torch.Tensor, **kw) -> "FlatIPIndex":
        """Adopt fp32 rows together with the bf16 shadow the encoder's pooling kernel already wrote
        (``encoder.encode_corpus``): no cast pass; only the largest row norm is computed."""
        if rows_f32.dtype != torch.float32 or rows_bf16.dtype != torch.bfloat16 or rows_f32.shape != rows_bf16.shape:
            raise TypeError("fp32 rows and a bf16 shadow of the same shape expected")
        idx = cls(rows_f32.shape[1], device=rows_f32.device, **kw)
        idx._f32, idx._bf16, idx.ntotal = rows_f32.contiguous(), rows_bf16.contiguous(), rows_f32.shape[0]
        if idx.ntotal:
            nr = torch.linalg.vector_norm(idx._f32, dim=1).max() * 1.000001
            idx._max_norm = torch.maximum(idx._max_norm, nr.reshape(1).float())
        return idx

    def reconstruct_rows(self, rows: torch.Tensor) -> torch.Tensor:
        """Stored vectors of the given local rows as fp32 ``[n, d]`` on the device (faiss ``reconstruct_batch``)."""
        rows = rows.to(self.device, dtype=torch.int64) - self.id_offset
        src = self._f32 if self._f32 is not None else self._bf16
        return src[: self.ntotal].index_select(0, rows).float()

    def _adopt(self, x: torch.Tensor) -> None:
        lib = _lib.load()
        x = x.contiguous()
        n = x.shape[0]
        if self.corpus_dtype == "fp32":
            if x.dtype != torch.float32:
                raise TypeError("fp32 corpus expected")
            self._f32 = x
            self._bf16 = torch.empty((n, self.d), dtype=torch.bfloat16, device=self.device)
            _lib.check(lib.ip_cast_rows_bf16(x.data_ptr(), self._bf16.data_ptr(), self._max_norm.data_ptr(), n, self.d,
                                             _lib.stream_ptr(self.device)), "ip_cast_rows_bf16")
        else:
            self._bf16 = x.to(torch.bfloat16)
            # row norms of the stored (bf16) rows, conservatively rounded up
            nr = torch.linalg.vector_norm(self._bf16.float(), dim=1).max() * 1.000001 if n else torch.zeros((), device=self.device)
            self._max_norm = torch.maximum(self._max_norm, nr.reshape(1).float())
        self.ntotal = n

    def reserve(self, n_rows: int) -> None:
        """Size the storage for ``n_rows`` rows up front (empty index only): later ``add`` calls then never re-allocate."""
        if self.ntotal != 0 or n_rows <= 0:
            return
        if self.corpus_dtype == "fp32":
            self._f32 = torch.empty((n_rows, self.d), dtype=torch.float32, device=self.device)
        self._bf16 = torch.empty((n_rows, self.d), dtype=torch.bfloat16, device=self.device)

    def add(self, x) -> None:
        """Append rows (numpy float32 ``[n, d]`` like faiss, or a torch tensor on any device).  Storage grows
        geometrically, so a sequence of appends costs O(total rows) copies; only the new rows are cast / measured."""
        lib = _lib.load()
        t = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)) if not torch.is_tensor(x) else x
        if t.dim() != 2 or t.shape[1] != self.d:
            raise ValueError(f"expected [n, {self.d}] rows")
        n = t.shape[0]
        if n == 0:
            return
        if self.ntotal + n > self.MAX_ROWS:
            raise _lib.SdagB200Error(f"an index shard holds at most {self.MAX_ROWS} rows (int32 row numbers in the candidate lists); "
                                     "partition larger corpora with ShardedFlatIPIndex")
        fp32 = self.corpus_dtype == "fp32"
        t = t.to(self.device, dtype=torch.float32 if fp32 else torch.bfloat16)
        if self.ntotal == 0 and (self._bf16 is None or self._bf16.shape[0] < n):
            self._adopt(t)
            return
        store = self._f32 if fp32 else self._bf16
        need = self.ntotal + n
        if store.shape[0] < need:  # amortised growth: double (at least), copy the live rows once
            cap = max(need, 2 * store.shape[0])
            if fp32:
                grown = torch.empty((cap, self.d), dtype=torch.float32, device=self.device)
                grown[: self.ntotal] = self._f32[: self.ntotal]
                self._f32 = grown
            grown16 = torch.empty((cap, self.d), dtype=torch.bfloat16, device=self.device)
            grown16[: self.ntotal] = self._bf16[: self.ntotal]
            self._bf16 = grown16
        lo = self.ntotal
        if fp32:
            self._f32[lo:need] = t
            # bf16 shadow + row norms of the new rows only; the kernel folds their largest norm into max_norm (atomic max)
            _lib.check(lib.ip_cast_rows_bf16(self._f32[lo:need].data_ptr(), self._bf16[lo:need].data_ptr(), self._max_norm.data_ptr(),
                                             n, self.d, _lib.stream_ptr(self.device)), "ip_cast_rows_bf16")
        else:
            self._bf16[lo:need] = t
            nr = torch.linalg.vector_norm(t.float(), dim=1).max() * 1.000001
            self._max_norm = torch.maximum(self._max_norm, nr.reshape(1).float())
        self.ntotal = need

    # -- search ---------------------------------------------------------------------------------------
    def search_device(self, xq: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """xq fp32 ``[Q, d]`` on the index's device -> (scores fp32 ``[Q,k]``, ids int64 ``[Q,k]``) on device.
        Rows sorted by descending score, ties by ascending id; ``-FLT_MAX`` / ``-1`` padding when ntotal < k."""
        lib = _lib.load()
        k = int(k)
        if xq.dim() != 2 or xq.shape[1] != self.d:
            raise ValueError(f"expected [Q, {self.d}] queries")
        if k <= 0:
            raise ValueError("k must be positive")
        if k > self.cand_cap // 32:
            raise _lib.SdagB200Error(f"k={k} exceeds the supported maximum {self.cand_cap // 32} (cand_cap / 32)")
        if self.d % 64 or self.d > 8192:
            raise _lib.SdagB200Error(f"d={self.d}: the tensor-core pass needs d % 64 == 0 and d <= 8192")
        xq = xq.to(self.device, dtype=torch.float32).contiguous()
        Q = xq.shape[0]
        scores = torch.empty((Q, k), dtype=torch.float32, device=self.device)
        ids = torch.empty((Q, k), dtype=torch.int64, device=self.device)
        if Q == 0:
            return scores, ids
        if self._bf16 is None:  # empty index
            self._bf16 = torch.zeros((1, self.d), dtype=torch.bfloat16, device=self.device)
        stats = torch.zeros(4, dtype=torch.int32, device=self.device)
        for q0 in range(0, Q, MAX_QUERY_BATCH):
            q1 = min(Q, q0 + MAX_QUERY_BATCH)
            nq = q1 - q0
            need = int(lib.ip_search_workspace_bytes(nq, self.d, k, self.cand_cap))
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            part_stats = torch.zeros(4, dtype=torch.int32, device=self.device)
            _lib.count_launches(2 + 2 * _n_epochs(self.ntotal, self.cand_cap) + 2)  # init+cast, (gemm+tighten)*, scan+rescore
            _lib.check(
                lib.ip_search(
                    xq[q0:q1].data_ptr(), nq,
                    _lib.ptr(self._f32) if self.corpus_dtype == "fp32" else 0,
                    self._bf16.data_ptr(), self.ntotal, self.d, self.id_offset, self._max_norm.data_ptr(), k,
                    scores[q0:q1].data_ptr(), ids[q0:q1].data_ptr(), part_stats.data_ptr(),
                    self._ws.data_ptr(), self._ws.numel(), self.cand_cap, _lib.stream_ptr(self.device)),
                "ip_search",
            )
            stats += part_stats
        self.last_stats = stats
        return scores, ids

    def search(self, x, k: int):
        """faiss call shape: numpy float32 ``[Q, d]`` in, ``(D float32 [Q,k], I int64 [Q,k])`` numpy out."""
        xq = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32))
        if torch.cuda.is_available():
            xq = xq.pin_memory()
        scores, ids = self.search_device(xq.to(self.device, non_blocking=True), k)
        return scores.cpu().numpy(), ids.cpu().numpy()


class ShardedFlatIPIndex:
    """Corpus rows block-partitioned over the ranks of a process group: every rank searches its
    shard for all queries, one ``all_gather`` moves the local ``[Q,k]`` results, every rank merges ``G*k -> k``
    with the same order rule.  Global id = shard row offset + local row."""

    def __init__(self, local: FlatIPIndex, group=None, merge_fn=None, exchange: Optional[str] = None):
        """``exchange``: ``"p2p"`` = one kernel over peer memory (``topk_exchange_merge``: remote stores over NVLink into a
        symmetric allocation, per-block flags, merge — no collective library call), ``"nccl"`` = ``topk_pack`` + one
        ``all_gather_into_tensor`` + ``topk_merge_packed``; default: the environment variable ``SDAG_EXCHANGE``, else
        ``"p2p"`` on CUDA (falling back to ``"nccl"``, with a ``note_fallback``, when symmetric memory cannot be set up)."""
        import torch.distributed as dist

        self.local = local
        self.group = group
        self.exchange = exchange or os.environ.get("SDAG_EXCHANGE") or "p2p"
        self._xchg = None
        # merge of the gathered [G,Q,k] lists; the CUDA kernel unless a test injects a checker for CPU (gloo) runs
        self.merge_fn = merge_fn or merge_topk_device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.d = local.d
        if self.world > 1:
            n = torch.tensor([local.ntotal], dtype=torch.int64, device=local.device)
            dist.all_reduce(n, group=group)
            self.ntotal = int(n.item())
        else:
            self.ntotal = local.ntotal

    def search_device(self, xq: torch.Tensor, k: int):
        s, i = self.local.search_device(xq, k)
        return self._finish(s, i, k)

    # -- pipelined form: the exchange of batch n runs on a side stream under the local search of batch n + 1 -------------------
    def submit(self, xq: torch.Tensor, k: int):
        """Start a search and return a handle for ``result``.  The local search is enqueued on the current stream; the
        exchange + merge (which has to wait for the slowest shard of THIS batch) goes to a side stream, so the next
        ``submit`` is not held back by it: a node's GPUs settle at different clocks under the power cap and a lock-step loop
        pays the slowest shard of every batch, a pipelined one the mean.  The two receive areas of the peer-memory exchange
        (alternating by call) are exactly what one batch of look-ahead needs; calls must be made in the same order on every
        rank, and at most one handle should be outstanding when the next ``submit`` is issued."""
        s, i = self.local.search_device(xq, k)
        if self.world == 1 or self.exchange != "p2p" or self.merge_fn is not merge_topk_device:
            return self._finish(s, i, k), None
        if self._xchg is None:  # first call: set the symmetric buffer up synchronously (a collective on the host side)
            out = self._search_exchange_p2p(s, i, k)
            return (out if out is not None else self._finish(s, i, k)), None
        dev = s.device
        if getattr(self, "_xstream", None) is None:
            self._xstream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._xstream):
            self._xstream.wait_event(ready)
            out = self._search_exchange_p2p(s, i, k)
            s.record_stream(self._xstream)
            i.record_stream(self._xstream)
            done = torch.cuda.Event()
            done.record(self._xstream)
        return out, done

    def result(self, handle):
        """``(scores, ids)`` of a ``submit``: the current stream waits for that batch's exchange."""
        out, done = handle
        if done is not None:
            cur = torch.cuda.current_stream(out[0].device)
            cur.wait_event(done)
            out[0].record_stream(cur)
            out[1].record_stream(cur)
        return out

    def _finish(self, s, i, k):
        """Exchange + merge of local results on the current stream (the tail of ``search_device``)."""
        import torch.distributed as dist

        if self.world == 1:
            return s, i
        Q = s.shape[0]
        if self.merge_fn is not merge_topk_device:
            gs = torch.empty((self.world * Q, k), dtype=s.dtype, device=s.device)
            gi = torch.empty((self.world * Q, k), dtype=i.dtype, device=i.device)
            dist.all_gather_into_tensor(gs, s.contiguous(), group=self.group)
            dist.all_gather_into_tensor(gi, i.contiguous(), group=self.group)
            return self.merge_fn(gs.view(self.world, Q, k), gi.view(self.world, Q, k), k)
        if self.exchange == "p2p":
            out = self._search_exchange_p2p(s, i, k)
            if out is not None:
                return out
        lib = _lib.load()
        rec = torch.empty((Q, k, 3), dtype=torch.int32, device=s.device)
        _lib.count_launches(2)
        _lib.check(lib.topk_pack(s.data_ptr(), i.data_ptr(), Q * k, rec.data_ptr(), _lib.stream_ptr(s.device)), "topk_pack")
        gathered = torch.empty((self.world, Q, k, 3), dtype=torch.int32, device=s.device)
        dist.all_gather_into_tensor(gathered, rec, group=self.group)
        out_s = torch.empty((Q, k), dtype=torch.float32, device=s.device)
        out_i = torch.empty((Q, k), dtype=torch.int64, device=s.device)
        _lib.check(lib.topk_merge_packed(gathered.data_ptr(), self.world, Q, k, out_s.data_ptr(), out_i.data_ptr(),
                                         _lib.stream_ptr(s.device)), "topk_merge_packed")
        return out_s, out_i

    def _search_exchange_p2p(self, s: torch.Tensor, i: torch.Tensor, k: int):
        """Exchange + merge in one kernel over peer memory; None when the symmetric buffer cannot be set up (every rank
        takes the same decision: the setup is a collective and fails or succeeds node-wide)."""
        import torch.distributed as dist

        lib = _lib.load()
        dev = s.device
        if self._xchg is None or self._xchg["k"] < k:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                group = self.group if self.group is not None else dist.group.WORLD
                nbytes = int(lib.topk_exchange_buffer_bytes(self.world, MAX_QUERY_BATCH, k))
                buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
                buf.zero_()
                hdl = symm_mem.rendezvous(buf, group)
                hdl.barrier()  # every rank's flags are zero before anybody raises one
                self._xchg = dict(buf=buf, hdl=hdl, k=k, epoch=0, peers=int(hdl.buffer_ptrs_dev), rank=dist.get_rank(group))
            except Exception as exc:  # no symmetric memory on this system / build: the NCCL path computes the same result
                _lib.note_fallback("ShardedFlatIPIndex", "peer-memory exchange unavailable: NCCL all-gather + merge", exc,
                                   fatal_in_strict=False)  # ~25 us per search slower, same result
                self.exchange = "nccl"
                return None
        x = self._xchg
        Q = s.shape[0]
        out_s = torch.empty((Q, k), dtype=torch.float32, device=dev)
        out_i = torch.empty((Q, k), dtype=torch.int64, device=dev)
        s, i = s.contiguous(), i.contiguous()
        for q0 in range(0, Q, MAX_QUERY_BATCH):
            q1 = min(Q, q0 + MAX_QUERY_BATCH)
            x["epoch"] += 1
            _lib.count_launches(1)
            _lib.check(lib.topk_exchange_merge(s[q0:q1].data_ptr(), i[q0:q1].data_ptr(), q1 - q0, MAX_QUERY_BATCH, k,
                                               x["rank"], self.world, x["peers"], x["epoch"], out_s[q0:q1].data_ptr(),
                                               out_i[q0:q1].data_ptr(), _lib.stream_ptr(dev)), "topk_exchange_merge")
        return out_s, out_i

    def search(self, x, k: int):
        xq = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(self.local.device)
        s, i = self.search_device(xq, k)
        return s.cpu().numpy(), i.cpu().numpy()


def merge_topk_device(scores: torch.Tensor, ids: torch.Tensor, k: int):
    """``[G,Q,k]`` per-shard results -> global ``[Q,k]`` (``topk_merge`` kernel)."""
    lib = _lib.load()
    G, Q, kk = scores.shape
    out_s = torch.empty((Q, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=scores.device)
    if kk != k:
        raise ValueError("shard results must hold k entries per query")
    _lib.check(lib.topk_merge(scores.contiguous().data_ptr(), ids.contiguous().data_ptr(), G, Q, k, out_s.data_ptr(),
                              out_i.data_ptr(), _lib.stream_ptr(scores.device)), "topk_merge")
    return out_s, out_i


# ------------------------------------------------------------------------------------------------------
# faiss IndexFlatIP file format — what faiss.write_index produces for an IndexFlatIP and what the
# reference hands to faiss.read_index (dense.py:36).  Restated from faiss' impl/index_write.cpp (write_index ->
# write_index_header + WRITEXBVECTOR) and impl/io_macros.h; no faiss in this image, so the layout is pinned by the
# hand-assembled fixture tests/golden/faiss_flat_ip_3x64.index (tests/golden/make_ixfi_fixture.py documents every byte).
# All fields little-endian, no padding:
#
#   offset  size  field         value / source
#   0       4     fourcc        "IxFI" (IndexFlat, METRIC_INNER_PRODUCT; "IxF2" = L2, "IxFl" = other metrics)   index_write.cpp
#   4       4     d             int32 vector dimension                                                  write_index_header
#   8       8     ntotal        int64 number of vectors                                                 write_index_header
#   16      8     dummy         int64 1 << 20 (legacy field, ignored on read)                           write_index_header
#   24      8     dummy         int64 1 << 20                                                           write_index_header
#   32      1     is_trained    uint8 bool (1 for flat indexes)                                         write_index_header
#   33      4     metric_type   int32: 0 = METRIC_INNER_PRODUCT, 1 = METRIC_L2                          write_index_header
#   [37     4     metric_arg    float32, only when metric_type > 1]                                     write_index_header
#   37      8     size          uint64 = ntotal * d: number of float32 values (the byte count / 4)      WRITEXBVECTOR
#   45      4*size  xb / codes  the vectors, row-major float32                                          WRITEXBVECTOR
# ------------------------------------------------------------------------------------------------------
_IXFI_HEADER = struct.Struct("<iqqqBi")  # d, ntotal, dummy, dummy, is_trained, metric_type (after the 4-byte fourcc)


def write_flat_ip_file(path: str, x: np.ndarray) -> None:
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    with open(path, "wb") as f:
        f.write(b"IxFI")
        f.write(_IXFI_HEADER.pack(d, n, 1 << 20, 1 << 20, 1, 0))
        f.write(struct.pack("<Q", n * d))
        f.write(x.tobytes())


def parse_flat_ip_file(path: str) -> np.ndarray:
    """The ``[ntotal, d]`` float32 vectors of a faiss ``IndexFlatIP`` file (host side of ``read_flat_ip_file``; the array
    is a read-only memory map).  ``ValueError`` for other index types / metrics and for truncated or inconsistent files."""
    size_on_disk = os.path.getsize(path)
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc != b"IxFI":
            kind = {b"IxF2": "an IndexFlatL2", b"IxFl": "an IndexFlat with another metric"}.get(fourcc, f"fourcc {fourcc!r}")
            raise ValueError(f"{path}: not a faiss IndexFlatIP file ({kind}); only flat inner-product indexes are supported")
        head = f.read(_IXFI_HEADER.size)
        if len(head) != _IXFI_HEADER.size:
            raise ValueError(f"{path}: truncated index header")
        d, n, _, _, _trained, metric = _IXFI_HEADER.unpack(head)
        if metric != 0:
            raise ValueError(f"{path}: metric_type {metric} is not METRIC_INNER_PRODUCT (0)")
        raw = f.read(8)
        if len(raw) != 8:
            raise ValueError(f"{path}: truncated index header")
        (count,) = struct.unpack("<Q", raw)
        offset = f.tell()
    if d <= 0 or n < 0 or count != n * d:
        raise ValueError(f"{path}: corrupt index file: {count} stored values for ntotal={n}, d={d}")
    if size_on_disk < offset + 4 * count:
        raise ValueError(f"{path}: truncated index file: {size_on_disk - offset} vector bytes, {4 * count} expected")
    if count == 0:
        return np.zeros((0, d), dtype=np.float32)
    return np.memmap(path, dtype="<f4", mode="r", offset=offset, shape=(n, d))


def read_flat_ip_file(path: str, device="cuda", **kw) -> FlatIPIndex:
    x = parse_flat_ip_file(path)
    idx = FlatIPIndex(x.shape[1], device=device, **kw)
    for r0 in range(0, x.shape[0], 1 << 20):  # 1M-row pieces: bounded pinned staging, storage sized once
        if r0 == 0 and x.shape[0] > (1 << 20):
            idx.reserve(x.shape[0])
        idx.add(np.array(x[r0 : r0 + (1 << 20)], dtype=np.float32))  # a writable copy of the mapped piece
    return idx


def load_faiss_index_and_meta(index_path: str, meta_path: str, device="cuda") -> Tuple[Any, List[Dict[str, Any]]]:
    """Same contract as the reference (``dense.py:15-41``): ``FileNotFoundError`` for a missing file, metadata
    is one JSON object per line."""
    if not os.path.exists(index_path):
        raise FileNotFoundError(f"FAISS index not found at {index_path}")
    if not os.path.exists(meta_path):
        raise FileNotFoundError(f"Metadata jsonl not found at {meta_path}")
    index = read_flat_ip_file(index_path, device=device)
    with open(meta_path, "r", encoding="utf-8") as f:
        meta = [json.loads(line) for line in f]
    return index, meta


def embed_queries(ranker, queries: List[str], batch_size: int) -> np.ndarray:
    """Unit-norm query embeddings, ``"query: "`` prefix iff the ranker is an e5 model (``dense.py:44-65``)."""
    use_prefix = "e5" in RANKER_MODEL_NAME.lower()
    texts = [("query: " + q) if use_prefix else q for q in queries]
    chunks = []
    for i in range(0, len(texts), batch_size):
        emb = ranker.encode(texts[i : i + batch_size], convert_to_tensor=True, batch_size=batch_size,
                            normalize_embeddings=True)
        chunks.append(emb.cpu().numpy())
    return np.vstack(chunks)


def search_index(index: Any, query_embeddings: np.ndarray, top_k: int) -> Tuple[np.ndarray, np.ndarray]:
    """``(indices, scores)`` — note the swap relative to ``index.search`` (``dense.py:82-84``)."""
    scores, indices = index.search(query_embeddings.astype(np.float32), top_k)
    return indices, scores


def materialize_dense_hits(indices, scores, meta):
    """ids -> (texts, ids, scores) lists; an index outside ``[0, len(meta))`` becomes ``""`` / ``"NA"`` with its
    score kept (``dense.py:87-130``)."""
    texts_b: List[List[str]] = []
    ids_b: List[List[str]] = []
    scores_b: List[List[float]] = []
    n_meta = len(meta)
    for row_idx, row_sc in zip(indices, scores):
        texts, ids, scs = [], [], []
        for idx, sc in zip(row_idx, row_sc):
            idx = int(idx)
            if 0 <= idx < n_meta:
                m = meta[idx]
                texts.append(m.get("text", ""))
                ids.append(str(m.get("id", "NA")))
            else:
                texts.append("")
                ids.append("NA")
            scs.append(float(sc))
        texts_b.append(texts)
        ids_b.append(ids)
        scores_b.append(scs)
    return texts_b, ids_b, scores_b


class DenseRetriever(Retriever):
    """Drop-in for the reference's ``DenseRetriever`` (``dense.py:133-178``); ``index`` is any faiss-shaped object,
    normally a ``FlatIPIndex`` / ``ShardedFlatIPIndex``."""

    def __init__(self, ranker, index: Any, meta: List[Dict[str, Any]]) -> None:
        self.ranker = ranker
        self.index = index
        self.meta = meta

    def retrieve_batch(self, queries: Sequence[str], max_k_needed: int, embed_batch_size: int) -> RetrievalBatch:
        q_embs = embed_queries(self.ranker, list(queries), batch_size=embed_batch_size)
        indices, scores = search_index(self.index, q_embs, top_k=max_k_needed)
        texts, ids, scs = materialize_dense_hits(indices, scores, self.meta)
        n_meta = len(self.meta)
        rows = [[int(i) if 0 <= int(i) < n_meta else -1 for i in row] for row in indices]
        return RetrievalBatch(q_embs=list(q_embs), docs_texts_full=texts, ids_full=ids, scores_full=scs, rows_full=rows)


B200DenseRetriever = DenseRetriever
