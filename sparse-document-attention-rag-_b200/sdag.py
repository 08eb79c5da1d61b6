"""SDAG generation path (hot path A) — host side.

Mirrors the public functions of the reference's ``src/pipeline/sparse_attention_RAG/SDAG.py`` (same names,
argument meaning and return values) on top of the CUDA library:

=============================================  ====================================================
reference (SDAG.py)                            here
=============================================  ====================================================
``compute_doc_knn_for_docs``          :14-65   same signature; k x k similarity + top-K on the GPU
``build_blocked_causal_mask_full``    :68-127  same signature; one kernel instead of L slice writes
``generate_with_custom_mask``         :130-213 same signature; mask -> block map, no [L,L] bias
``build_rag_chat_and_spans``          :216-304 same signature (host string / tokenizer logic)
``run_rag_with_doc_isolation``        :307-375 same signature; never builds the dense mask
=============================================  ====================================================

The attention itself is plugged into Hugging Face through ``AttentionInterface`` under the name
``"sdag_b200"`` (``register_attention()``); span metadata reaches the kernel as the ``sdag_meta`` keyword
argument of ``model(...)`` instead of a dense ``attention_mask``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import config as _config
from .prompts import QUESTION_MARKER, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT
from .ranked_list import inject_malicious_docs_into_ranked_list

# Patchable like the reference's module-level copies (SDAG.py:9).
MAX_GEN_TOKENS_RAG = _config.MAX_GEN_TOKENS_RAG
TEMPERATURE = _config.TEMPERATURE

ATTN_NAME = "sdag_b200"


# ------------------------------------------------------------------------------------------------------
# Layout -> device records
# ------------------------------------------------------------------------------------------------------
@dataclass
class SdagLayout:
    """Arguments of ``build_blocked_causal_mask_full`` for one prompt (SDAG.py:68-76)."""

    L: int
    S: int
    spans: Sequence[Tuple[int, int]]
    qa_start: int
    neighbors: Optional[List[List[int]]] = None

    @property
    def use_neighbors(self) -> bool:  # SDAG.py:111
        return self.neighbors is not None and len(self.neighbors) == len(self.spans)

    @property
    def nbr_k(self) -> int:
        return max((len(n) for n in self.neighbors), default=0) if self.use_neighbors else 0


def _pack_layouts(layouts: Sequence[SdagLayout], max_docs: int, nbr_k: int) -> np.ndarray:
    stride = _lib.LAYOUT_HEADER + 2 * max_docs + max_docs * nbr_k
    rec = np.zeros((len(layouts), stride), dtype=np.int32)
    for i, lay in enumerate(layouts):
        n = len(lay.spans)
        rec[i, 0:5] = (lay.L, lay.S, lay.qa_start, n, 1 if lay.use_neighbors else 0)
        if n:
            rec[i, _lib.LAYOUT_HEADER : _lib.LAYOUT_HEADER + 2 * n] = np.asarray(lay.spans, dtype=np.int64).reshape(-1).clip(
                -(2**31), 2**31 - 1
            )
        nb = rec[i, _lib.LAYOUT_HEADER + 2 * max_docs :].reshape(max_docs, nbr_k) if nbr_k else None
        if nb is not None:
            nb[:] = -1
            if lay.use_neighbors:
                for d, lst in enumerate(lay.neighbors):
                    if lst:
                        # negative ids are skipped by the reference (SDAG.py:119); keep them distinct from padding
                        nb[d, : len(lst)] = [v if v >= 0 else -2 for v in lst]
    return rec


def _to_device_i32(arr: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(arr).pin_memory().to(device, non_blocking=True) if torch.cuda.is_available() else torch.from_numpy(arr)


class BlockMap:
    """Device-resident tile lists for a batch of prompts (see include/sdag_b200.h for the record layout)."""

    def __init__(self, buf: torch.Tensor, stride_bytes: int, cap_qtiles: int, n_prompts: int, max_L: int, kind: str):
        self.buf = buf
        self.stride_bytes = stride_bytes
        self.cap_qtiles = cap_qtiles
        self.n_prompts = n_prompts
        self.max_L = max_L
        self.kind = kind  # "spans" | "mask"

    def header(self, i: int = 0) -> dict:
        """Host copy of prompt i's header (synchronises; tests / benchmarks only)."""
        stride = self.stride_bytes // 4
        h = self.buf.view(-1)[i * stride : i * stride + _lib.BLOCKMAP_HEADER].cpu().numpy()
        return dict(n_qtiles=int(h[0]), n_kvtiles=int(h[1]), status=int(h[2]), n_gap_rows=int(h[3]), L=int(h[4]),
                    cap_qtiles=int(h[5]), cap_kvtiles=int(h[6]), words_per_row=int(h[7]),
                    pairs=(int(h[8]) & 0xFFFFFFFF) | (int(h[9]) << 32), bits_off=int(h[10]))

    def tiles(self, i: int = 0):
        """(qtiles[n,4], kvtiles[m,4], raw int32 buffer) of prompt i on the host (tests only)."""
        stride = self.stride_bytes // 4
        raw = self.buf.view(-1)[i * stride : (i + 1) * stride].cpu().numpy()
        h = self.header(i)
        q = raw[_lib.BLOCKMAP_HEADER : _lib.BLOCKMAP_HEADER + 4 * h["cap_qtiles"]].reshape(-1, 4)[: h["n_qtiles"]]
        kv0 = _lib.BLOCKMAP_HEADER + 4 * h["cap_qtiles"]
        kv = raw[kv0 : kv0 + 4 * h["cap_kvtiles"]].reshape(-1, 4)[: h["n_kvtiles"]]
        return q, kv, raw


def build_blockmap(layouts: Sequence[SdagLayout], device="cuda") -> BlockMap:
    """Block maps of a batch of prompts straight from their token offsets (``sdag_build_blockmap``).

    Layouts the reference's prompt builder cannot produce (overlapping / unsorted spans, more than 16
    neighbours) come back with ``status != 0``; ``build_blockmap_checked`` falls back to the dense-mask path.
    """
    lib = _lib.load()
    layouts = list(layouts)
    max_docs = max(1, max(len(l.spans) for l in layouts))
    nbr_k = max(l.nbr_k for l in layouts)
    max_L = max(l.L for l in layouts)
    if max_docs > _lib.MAX_DOCS or nbr_k > _lib.MAX_NBR_K:
        raise _lib.SdagB200Error("layout exceeds the span-path limits (docs <= 512, neighbours <= 16)")
    rec = _to_device_i32(_pack_layouts(layouts, max_docs, nbr_k), device)
    stride_bytes = int(lib.sdag_blockmap_bytes(max_L, max_docs, nbr_k))
    buf = torch.empty((len(layouts), stride_bytes // 4), dtype=torch.int32, device=device)
    _lib.check(
        lib.sdag_build_blockmap(rec.data_ptr(), len(layouts), max_docs, nbr_k, max_L, buf.data_ptr(), stride_bytes,
                                _lib.stream_ptr(buf.device)),
        "sdag_build_blockmap",
    )
    cap_q = (max_L + _lib.SDAG_TILE - 1) // _lib.SDAG_TILE + 2 * max_docs + 3
    bm = BlockMap(buf, stride_bytes, cap_q, len(layouts), max_L, "spans")
    bm._layout_rec = rec  # keep alive until the kernel has run
    return bm


def build_blockmap_from_mask(mask: torch.Tensor) -> BlockMap:
    """Block map of an arbitrary bool ``[L, L]`` allow-mask (the general ``generate_with_custom_mask`` input)."""
    lib = _lib.load()
    if mask.dim() != 2 or mask.shape[0] != mask.shape[1]:
        raise ValueError("prompt mask must be [L, L]")
    if mask.dtype != torch.bool:
        # additive float bias (SDAG.py:170-171): 0 keeps a key; the reference's own fill value (finfo.min) or
        # -inf drops it.  Arbitrary finite biases are not representable by a tile list.
        allowed = mask == 0
        dropped = (mask <= torch.finfo(mask.dtype).min) | torch.isinf(mask)
        if not bool((allowed | dropped).all()):
            raise NotImplementedError("only {0, finfo.min / -inf} additive masks are supported")
        mask = allowed
    L = mask.shape[0]
    m8 = mask.contiguous().view(torch.uint8)
    nbytes = int(lib.sdag_maskmap_bytes(L))
    buf = torch.empty((1, nbytes // 4), dtype=torch.int32, device=mask.device)
    _lib.check(lib.sdag_blockmap_from_mask(m8.data_ptr(), L, L, buf.data_ptr(), nbytes, _lib.stream_ptr(mask.device)),
               "sdag_blockmap_from_mask")
    T = (L + _lib.SDAG_TILE - 1) // _lib.SDAG_TILE
    bm = BlockMap(buf, nbytes, T, 1, L, "mask")
    bm._mask = m8
    return bm


def build_blockmap_checked(layout: SdagLayout, device="cuda") -> BlockMap:
    """Span path when the layout is canonical, dense-mask path otherwise (one host sync to read the status)."""
    if len(layout.spans) <= _lib.MAX_DOCS and layout.nbr_k <= _lib.MAX_NBR_K:
        bm = build_blockmap([layout], device)
        if bm.header(0)["status"] == 0:
            return bm
    mask = build_blocked_causal_mask_full(layout.L, 0, layout.S, list(layout.spans), layout.qa_start, device,
                                          layout.neighbors)
    return build_blockmap_from_mask(mask)


_causal_cache = {}


def causal_blockmap(L: int, device) -> BlockMap:
    """Plain causal attention (the CARG baseline) expressed as a layout whose every row is a system row."""
    key = (L, str(device))
    bm = _causal_cache.get(key)
    if bm is None:
        bm = build_blockmap([SdagLayout(L=L, S=L, spans=[], qa_start=L)], device)
        if len(_causal_cache) > 64:
            _causal_cache.clear()
        _causal_cache[key] = bm
    return bm


# ------------------------------------------------------------------------------------------------------
# Attention calls
# ------------------------------------------------------------------------------------------------------
_workspaces = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = str(device)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def sdag_prefill_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, blockmap: BlockMap,
                           scaling: Optional[float] = None) -> torch.Tensor:
    """q ``[B,H,L,D]`` (any strides with unit innermost), k/v ``[B,Hkv,S,D]`` bf16 -> out ``[B,L,H,D]`` bf16."""
    lib = _lib.load()
    if q.dtype != torch.bfloat16 or k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
        raise TypeError("sdag_b200 attention computes in bf16 (fp32 softmax / accumulation)")
    B, H, L, D = q.shape
    Hkv, S = k.shape[1], k.shape[2]
    if blockmap.n_prompts not in (1, B):
        raise ValueError("one block map per batch entry (or a single shared one) expected")
    if blockmap.max_L != L or S != L:
        raise ValueError(f"block map was built for L={blockmap.max_L}, got Lq={L}, Skv={S}")
    q, k, v = (_unit_inner(t) for t in (q, k, v))
    out = torch.empty((B, L, H, D), dtype=torch.bfloat16, device=q.device)
    ws_bytes = int(lib.sdag_prefill_workspace_bytes(B, Hkv, D))
    ws = _workspace(q.device, ws_bytes)
    scaling = float(D ** -0.5 if scaling is None else scaling)
    _lib.check(
        lib.sdag_prefill_fwd(
            q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
            k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
            v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
            out.data_ptr(), out.stride(0), out.stride(1), out.stride(2),
            B, H, Hkv, L, S, D, scaling,
            blockmap.buf.data_ptr(), blockmap.stride_bytes if blockmap.n_prompts > 1 else 0, blockmap.cap_qtiles,
            ws.data_ptr(), ws.numel(), _lib.stream_ptr(q.device)),
        "sdag_prefill_fwd",
    )
    return out


def _unit_inner(t: torch.Tensor) -> torch.Tensor:
    if t.stride(-1) != 1 or any(s % 8 for s in t.stride()[:-1]) or t.data_ptr() % 16:
        return t.contiguous()
    return t


# ------------------------------------------------------------------------------------------------------
# build_blocked_causal_mask_full — SDAG.py:68-127
# ------------------------------------------------------------------------------------------------------
def build_blocked_causal_mask_full(
    base_seq_len: int,
    max_new_tokens: int,
    sys_user_len: int,
    doc_token_spans: list,
    qa_start: int,
    device: str = "cuda",
    doc_neighbors: Optional[List[List[int]]] = None,
) -> torch.Tensor:
    """Dense bool ``[L, L]`` allow-mask, bit-identical to the reference builder (``max_new_tokens`` is unused
    there too, SDAG.py:70).  Kept for callers / tests that want the mask itself; the generation path below
    never materialises it."""
    lib = _lib.load()
    lay = SdagLayout(int(base_seq_len), int(sys_user_len), [tuple(map(int, s)) for s in doc_token_spans], int(qa_start),
                     doc_neighbors)
    if len(lay.spans) > _lib.MAX_DOCS:
        raise _lib.SdagB200Error("more than 512 documents")
    max_docs = max(1, len(lay.spans))
    rec = _to_device_i32(_pack_layouts([lay], max_docs, lay.nbr_k), device)
    mask = torch.empty((lay.L, lay.L), dtype=torch.uint8, device=device)
    _lib.check(lib.sdag_build_mask(rec.data_ptr(), max_docs, lay.nbr_k, mask.data_ptr(), lay.L, _lib.stream_ptr(mask.device)),
               "sdag_build_mask")
    return mask.view(torch.bool)
