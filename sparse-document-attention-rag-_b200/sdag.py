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
USE_FUSED_PREFILL = True  # False forces the Hugging Face forward (+ attention hook) for the prefill


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
        self.launch_qtiles = cap_qtiles  # grid extent in query tiles (>= the device-side count)
        self._ok: Optional[bool] = None if kind == "spans" else True  # span maps: status words not read back yet

    def is_ok(self) -> bool:
        """True when every prompt's map was built (header status 0).  The span builder rejects layouts outside its
        canonical form (``qa_start`` inside a document span, overlapping / unsorted spans, ``S`` past the first span) with a
        nonzero status and an empty tile list; the attention kernel would then leave the output unwritten.  Read back
        once per map (one small device-to-host copy), then cached."""
        if self._ok is None:
            stride = self.stride_bytes // 4
            status = self.buf.view(-1)[2 : 2 + stride * self.n_prompts : stride]
            self._ok = not bool((status != 0).any())
        return self._ok

    def ensure_ok(self) -> None:
        if not self.is_ok():
            raise _lib.SdagB200Error(
                "block map has a nonzero status: the layout is not in the span builder's canonical form "
                "(use build_blockmap_checked / build_blockmap_from_mask for such layouts)")

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


def _count_qtiles(lay: SdagLayout) -> int:
    """Query-row tiles the device builder will emit for a canonical layout (an upper bound otherwise): one list
    of <=128-row tiles per segment (system block, gaps, documents, question block)."""
    T = _lib.SDAG_TILE
    tiles = lambda n: (max(n, 0) + T - 1) // T  # noqa: E731
    qa = min(lay.qa_start, lay.L)
    n, pos = tiles(lay.S), lay.S
    for s, e in lay.spans:
        n += tiles(s - pos) + tiles(e - s)
        pos = max(pos, e)
    return n + tiles(qa - pos) + tiles(lay.L - qa) + 1


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
    _lib.count_launches(1)
    _lib.check(
        lib.sdag_build_blockmap(rec.data_ptr(), len(layouts), max_docs, nbr_k, max_L, buf.data_ptr(), stride_bytes,
                                _lib.stream_ptr(buf.device)),
        "sdag_build_blockmap",
    )
    cap_q = (max_L + _lib.SDAG_TILE - 1) // _lib.SDAG_TILE + 2 * max_docs + 3
    bm = BlockMap(buf, stride_bytes, cap_q, len(layouts), max_L, "spans")
    bm.launch_qtiles = min(cap_q, max(_count_qtiles(l) for l in layouts))
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
        bm.no_uniform_rows = True  # plain causal: every row sees at least itself
        if len(_causal_cache) > 64:
            _causal_cache.clear()
        _causal_cache[key] = bm
    return bm


# ------------------------------------------------------------------------------------------------------
# Attention calls
# ------------------------------------------------------------------------------------------------------
_workspaces = {}
_profile_events = None  # set to a list to collect (start, end) CUDA events around each prefill attention call


def _workspace(device, nbytes: int) -> torch.Tensor:
    key = str(device)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def sdag_prefill_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, blockmap: BlockMap,
                           scaling: Optional[float] = None, ragged: bool = False, kernel: Optional[str] = None,
                           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q ``[B,H,L,D]`` (any strides with unit innermost), k/v ``[B,Hkv,S,D]`` bf16 -> out ``[B,L,H,D]`` bf16.

    ``kernel``: None = the library's choice; ``"pair"`` / ``"single"`` force the CTA-pair kernel (head_dim 128 and an even
    number of query heads per kv head, else an error) or the single-CTA kernel (tests, A/B measurements).

    ``ragged``: the block maps were built for prompts shorter than L (right-padded batch).  Rows past a prompt's end
    are not covered by any tile; they are returned as zeros so that everything downstream of them — in particular the
    K / V rows the next layer derives from them, which sit inside the last kv tile of the prompt — stays finite.

    ``out``: caller-owned contiguous ``[B,L,H,D]`` bf16 result buffer (a layer loop re-uses one instead of allocating — and,
    when ``ragged``, zero-filling — per call; with ``ragged`` its rows past each prompt's end must already be zero: the
    kernel never writes them)."""
    lib = _lib.load()
    if q.dtype != torch.bfloat16 or k.dtype != torch.bfloat16 or v.dtype != torch.bfloat16:
        raise TypeError("sdag_b200 attention computes in bf16 (fp32 softmax / accumulation)")
    B, H, L, D = q.shape
    Hkv, S = k.shape[1], k.shape[2]
    if blockmap.n_prompts not in (1, B):
        raise ValueError("one block map per batch entry (or a single shared one) expected")
    if blockmap.max_L != L or S != L:
        raise ValueError(f"block map was built for L={blockmap.max_L}, got Lq={L}, Skv={S}")
    blockmap.ensure_ok()  # a rejected layout must not reach the kernel (its output rows would stay uninitialised)
    q, k, v = (_unit_inner(t) for t in (q, k, v))
    if out is None:
        out = (torch.zeros if ragged else torch.empty)((B, L, H, D), dtype=torch.bfloat16, device=q.device)
    elif out.shape != (B, L, H, D) or out.dtype != torch.bfloat16 or not out.is_contiguous() or out.device != q.device:
        raise ValueError("out must be a contiguous bf16 [B, L, H, D] tensor on q's device")
    ws_bytes = int(lib.sdag_prefill_workspace_bytes(B, Hkv, D))
    ws = _workspace(q.device, ws_bytes)
    scaling = float(D ** -0.5 if scaling is None else scaling)
    ev = None
    if _profile_events is not None:  # bench.py: CUDA-event timing of the attention launches, on their stream
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _lib.count_launches(1 if getattr(blockmap, "no_uniform_rows", False) else 3)  # (v_colsum + v_mean +) sdag_prefill
    _lib.check(
        lib.sdag_prefill_fwd(
            q.data_ptr(), q.stride(0), q.stride(1), q.stride(2),
            k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
            v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
            out.data_ptr(), out.stride(0), out.stride(1), out.stride(2),
            B, H, Hkv, L, S, D, scaling,
            blockmap.buf.data_ptr(), blockmap.stride_bytes if blockmap.n_prompts > 1 else 0, blockmap.launch_qtiles,
            ws.data_ptr(), ws.numel(),
            (1 if getattr(blockmap, "no_uniform_rows", False) else 0) | {None: 0, "single": 2, "pair": 4}[kernel],
            _lib.stream_ptr(q.device)),
        "sdag_prefill_fwd",
    )
    if ev is not None:
        ev[1].record()
        _profile_events.append(ev)
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


# ------------------------------------------------------------------------------------------------------
# compute_doc_knn_for_docs — SDAG.py:14-65
# ------------------------------------------------------------------------------------------------------
def doc_knn_device(emb: torch.Tensor, k_neighbors: int) -> torch.Tensor:
    """``emb [n, d]`` fp32 on the GPU -> int32 ``[n, k]`` neighbour rows, most similar first (-1 = none)."""
    lib = _lib.load()
    emb = emb.to(torch.float32).contiguous()
    n, d = emb.shape
    out = torch.empty((n, k_neighbors), dtype=torch.int32, device=emb.device)
    _lib.check(lib.doc_knn(emb.data_ptr(), n, d, k_neighbors, out.data_ptr(), _lib.stream_ptr(emb.device)), "doc_knn")
    return out


def compute_doc_knn_for_docs(ranker, docs: List[str], k_neighbors: int, device="cuda") -> List[List[int]]:
    """Directed kNN lists among the prompt documents (indices into ``docs``).  Same conventions as the
    reference: empty / whitespace docs take no part and get ``[]``; the ``"passage: "`` prefix is unconditional
    (SDAG.py:47); fewer than two usable docs or ``k_neighbors <= 0`` -> all lists empty."""
    n = len(docs)
    if k_neighbors <= 0 or n == 0:
        return [[] for _ in range(n)]
    usable = [i for i, d in enumerate(docs) if d and d.strip()]
    if len(usable) <= 1:
        return [[] for _ in range(n)]
    emb = ranker.encode(["passage: " + docs[i] for i in usable], convert_to_tensor=True, normalize_embeddings=True)
    emb = torch.as_tensor(emb).to(device=device, dtype=torch.float32)
    local = doc_knn_device(emb, int(k_neighbors)).cpu().tolist()
    out: List[List[int]] = [[] for _ in range(n)]
    for row, i_full in enumerate(usable):
        out[i_full] = [usable[j] for j in local[row] if j >= 0]
    return out


# ------------------------------------------------------------------------------------------------------
# Hugging Face attention hook
# ------------------------------------------------------------------------------------------------------
def decode_workspace_bytes(B: int, H: int, D: int, S: int) -> int:
    """Bytes of split-KV scratch one ``sdag_decode_attention`` call needs (``sdag_decode_workspace_bytes``)."""
    return int(_lib.load().sdag_decode_workspace_bytes(B, H, D, S))


def sdag_decode_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scaling: Optional[float] = None,
                          kv_len_dev: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                          workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One decode step, dense over the cache: q ``[B,H,1,D]``, k/v ``[B,Hkv,S,D]`` bf16 -> ``[B,1,H,D]``.
    ``kv_len_dev`` (int32 device scalar, or one value per prompt for ragged batches) makes the valid cache length a
    device-side value (``S`` is then the bound the grid is sized for) so the call can sit inside a CUDA graph that is
    replayed per token.  ``workspace`` (uint8, at least ``decode_workspace_bytes``) is the split-KV scratch; a caller
    that captures the call into a CUDA graph MUST pass its own, kept alive with the graph — the shared per-device
    scratch is replaced when a later call needs a larger one, and a captured pointer to it would dangle."""
    lib = _lib.load()
    B, H, _, D = q.shape
    Hkv, S = k.shape[1], k.shape[2]
    q2 = q.reshape(B, H, D)
    if q2.stride(-1) != 1:
        q2 = q2.contiguous()
    k, v = _unit_inner(k), _unit_inner(v)
    if out is None:
        out = torch.empty((B, 1, H, D), dtype=torch.bfloat16, device=q.device)
    ws_bytes = int(lib.sdag_decode_workspace_bytes(B, H, D, S))
    if workspace is not None:
        if workspace.numel() * workspace.element_size() < ws_bytes or workspace.device != q.device:
            raise _lib.SdagB200Error("decode workspace too small or on another device")
        ws = workspace
    else:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.SdagB200Error("sdag_decode_attention inside a CUDA graph needs a caller-owned workspace")
        ws = _workspace(q.device, ws_bytes)
    scaling = float(D ** -0.5 if scaling is None else scaling)
    _lib.count_launches(2)  # split + merge
    _lib.check(
        lib.sdag_decode_fwd(q2.data_ptr(), q2.stride(0), q2.stride(1),
                            k.data_ptr(), k.stride(0), k.stride(1), k.stride(2),
                            v.data_ptr(), v.stride(0), v.stride(1), v.stride(2),
                            out.data_ptr(), B, H, Hkv, S, D, scaling, ws.data_ptr(), ws.numel() * ws.element_size(),
                            _lib.ptr(kv_len_dev),
                            1 if (kv_len_dev is not None and kv_len_dev.numel() == B and B > 1) else 0,
                            _lib.stream_ptr(q.device)),
        "sdag_decode_fwd",
    )
    return out


def sdag_b200_attention(module, query, key, value, attention_mask=None, dropout: float = 0.0, scaling=None,
                        sdag_meta: Optional[BlockMap] = None, **kwargs):
    """``AttentionInterface`` function (same contract as HF ``sdpa_attention_forward``,
    integrations/sdpa_attention.py:40-104): returns ``(attn_output [B,Lq,H,D], None)``.

    * prefill (``Lq == S``): block-sparse attention under ``sdag_meta`` (a ``BlockMap``), or plain causal when
      no metadata is given (the CARG path);
    * decode (``Lq == 1``): dense over the cache, exactly what the reference does after the prefill
      (SDAG.py:193-197 passes no mask).
    """
    if dropout:
        raise NotImplementedError("sdag_b200 attention is inference-only (dropout must be 0)")
    B, H, Lq, D = query.shape
    S = key.shape[2]
    orig = query.dtype
    if orig != torch.bfloat16:
        query, key, value = query.to(torch.bfloat16), key.to(torch.bfloat16), value.to(torch.bfloat16)
    if Lq == 1:
        out = sdag_decode_attention(query, key, value, scaling)
    elif Lq == S:
        bm = sdag_meta if sdag_meta is not None else causal_blockmap(Lq, query.device)
        out = sdag_prefill_attention(query, key, value, bm, scaling)
    else:
        raise NotImplementedError("chunked prefill against an existing cache is not part of the reference path")
    return (out if orig == torch.bfloat16 else out.to(orig)), None


def register_attention() -> str:
    """Register the hook with transformers (idempotent) and return its name."""
    from transformers import AttentionInterface

    AttentionInterface.register(ATTN_NAME, sdag_b200_attention)
    return ATTN_NAME


def prepare_model(model):
    """Point a HF causal LM at the ``sdag_b200`` attention function."""
    register_attention()
    if getattr(model.config, "_attn_implementation", None) != ATTN_NAME:
        try:
            model.set_attn_implementation(ATTN_NAME)
        except (AttributeError, ValueError, TypeError) as exc:
            # older transformers / non-PreTrainedModel wrappers: no setter, the config field is what the layers read
            _lib.note_fallback("prepare_model", "set_attn_implementation unavailable, config field set directly", exc)
            model.config._attn_implementation = ATTN_NAME
    return model


# ------------------------------------------------------------------------------------------------------
# generate_with_custom_mask — SDAG.py:130-213
# ------------------------------------------------------------------------------------------------------
def _pick(logits: torch.Tensor) -> torch.Tensor:
    if TEMPERATURE > 0:  # SDAG.py:182-185 — global torch RNG, like the reference
        probs = torch.softmax(logits / TEMPERATURE, dim=-1)
        return torch.multinomial(probs, num_samples=1)
    return torch.argmax(logits, dim=-1, keepdim=True)


def prefill_forward(model, input_ids: torch.Tensor, blockmap: Optional[BlockMap]):
    """Last-position logits ``[B, vocab]`` and the KV cache of a prompt batch under a block map (None = causal).

    Llama-family bf16 models go through ``llama_runner.FusedPrefill`` (fused element-wise glue around the same
    attention kernel); anything else through the Hugging Face forward with the ``sdag_b200`` attention hook."""
    from .llama_runner import get_runner

    prepare_model(model)
    runner = get_runner(model) if USE_FUSED_PREFILL else None
    with torch.no_grad():
        if runner is not None:
            bm = blockmap if blockmap is not None else causal_blockmap(input_ids.shape[1], input_ids.device)
            return runner.prefill(input_ids, bm)
        out = model(input_ids=input_ids, use_cache=True, sdag_meta=blockmap, logits_to_keep=1)
        return out.logits[:, -1, :], out.past_key_values


def _generate(model, tokenizer, input_ids: torch.Tensor, blockmap: BlockMap, max_new_tokens: int) -> str:
    prepare_model(model)
    eos = tokenizer.eos_token_id
    greedy = not (TEMPERATURE > 0)
    from .llama_runner import get_runner

    runner = get_runner(model) if (USE_FUSED_PREFILL and greedy) else None
    if runner is not None:
        # Llama-family bf16 model, greedy: fused prefill into a preallocated cache + CUDA-graph decode loop.
        # (Sampling keeps the reference's per-token loop below so the global torch RNG is consumed identically.)
        ids = runner.generate(input_ids, blockmap, max_new_tokens, eos, temperature=0.0)[0].tolist()
        return tokenizer.decode(_trim_at_eos(ids, eos), skip_special_tokens=True).strip()
    with torch.no_grad():
        logits, past = prefill_forward(model, input_ids, blockmap)
        nxt = _pick(logits)
        new_tokens = [nxt]
        # One token comes from the prefill, up to max_new_tokens more from the loop; EOS is only tested inside
        # the loop (Appendix A7).  Greedy decoding keeps tokens on the device and polls for EOS every 16 steps;
        # the tail after the first in-loop EOS is cut below, so the result equals a per-token check.
        for step in range(max_new_tokens):
            out = model(input_ids=nxt, past_key_values=past, use_cache=True)
            past = out.past_key_values
            nxt = _pick(out.logits[:, -1, :])
            new_tokens.append(nxt)
            if not greedy:
                if nxt.item() == eos:
                    break
            elif eos is not None and (step % 16 == 15) and bool((torch.cat(new_tokens[1:], dim=1) == eos).any()):
                break
    ids = torch.cat(new_tokens, dim=1)[0].tolist()
    return tokenizer.decode(_trim_at_eos(ids, eos), skip_special_tokens=True).strip()


def _trim_at_eos(ids: List[int], eos) -> List[int]:
    """Cut after the first EOS produced inside the decode loop; the prefill token (index 0) never stops generation
    (SURVEY.md Appendix A7)."""
    if eos is not None:
        for pos in range(1, len(ids)):
            if ids[pos] == eos:
                return ids[: pos + 1]
    return ids


def generate_with_custom_mask(model, tokenizer, input_ids: torch.Tensor, prompt_mask: torch.Tensor,
                              max_new_tokens: int = 128) -> str:
    """Prefill under an arbitrary ``[L, L]`` allow-mask (bool, or the additive {0, finfo.min} form), then dense
    token-by-token decoding; returns the decoded new tokens (special tokens removed, stripped)."""
    device = input_ids.device
    if device.type != "cuda":
        raise _lib.SdagB200Error("sdag_b200 runs on CUDA devices only (no CPU fallback)")
    bm = build_blockmap_from_mask(prompt_mask.to(device))
    return _generate(model, tokenizer, input_ids, bm, max_new_tokens)


# ------------------------------------------------------------------------------------------------------
# build_rag_chat_and_spans — SDAG.py:216-304
# ------------------------------------------------------------------------------------------------------
def _token_len(tokenizer, text: str) -> int:
    return len(tokenizer(text, return_tensors="pt").input_ids[0])


def build_rag_chat_and_spans(tokenizer, system_prompt: str, user_template: str, query: str, malicious_docs: List[str],
                             retrieved_docs: list, add_attack_in_rank: int):
    """Prompt string plus the token offsets the mask needs: ``(chat_str, sys_user_len, doc_token_spans, qa_start,
    ranked_docs)``.  Offsets are counted by tokenising string prefixes, as the reference does (Appendix A10), so
    they agree with it for any tokenizer; all prefixes go through the tokenizer in one batched call."""
    ranked_docs = inject_malicious_docs_into_ranked_list(base_docs=retrieved_docs, malicious_docs=malicious_docs,
                                                         attack_pos=add_attack_in_rank)
    bullets = ["- " + d.strip() for d in ranked_docs if d and d.strip()]
    user_content = user_template.format(query=query, docs_text="\n\n".join(bullets))
    chat_str = tokenizer.apply_chat_template(
        [{"role": "system", "content": system_prompt}, {"role": "user", "content": user_content}],
        tokenize=False, add_generation_prompt=True)

    starts, cursor = [], 0
    for b in bullets:
        at = chat_str.find(b, cursor)
        if at < 0:
            at = cursor
        starts.append(at)
        cursor = at + len(b)
    q_at = chat_str.find(QUESTION_MARKER)
    if q_at < 0:
        q_at = len(chat_str)

    cuts = [starts[0] if starts else q_at]
    for b, at in zip(bullets, starts):
        cuts += [at, at + len(b)]
    cuts.append(q_at)
    try:
        lens = [len(ids) for ids in tokenizer([chat_str[:c] for c in cuts]).input_ids]
    except Exception:  # tokenizers without batched __call__
        lens = [_token_len(tokenizer, chat_str[:c]) for c in cuts]
    sys_user_len = lens[0]
    spans = [(lens[1 + 2 * i], lens[2 + 2 * i]) for i in range(len(bullets))]
    qa_start = lens[-1]
    return chat_str, sys_user_len, spans, qa_start, ranked_docs


# ------------------------------------------------------------------------------------------------------
# run_rag_with_doc_isolation — SDAG.py:307-375
# ------------------------------------------------------------------------------------------------------
def run_rag_with_doc_isolation(model, tokenizer, query: str, malicious_docs: List[str], retrieved_docs: list,
                               max_new_tokens: int, device: str, add_attack_in_rank: int, ranker,
                               doc_neighbors_k: int) -> str:
    """SDAG answer for one query.  Same steps as the reference, except that the ``[L, L]`` mask is never built:
    the token offsets go to the device and come back as a block map consumed by the attention kernel."""
    if torch.device(device).type != "cuda":
        raise _lib.SdagB200Error("sdag_b200 runs on CUDA devices only (no CPU fallback)")
    chat_str, sys_user_len, spans, qa_start, ranked_docs = build_rag_chat_and_spans(
        tokenizer=tokenizer, system_prompt=SYSTEM_PROMPT_RAG, user_template=USER_RAG_PROMPT, query=query,
        malicious_docs=malicious_docs, retrieved_docs=retrieved_docs, add_attack_in_rank=add_attack_in_rank)
    neighbors = None
    if doc_neighbors_k and doc_neighbors_k > 0:
        neighbors = compute_doc_knn_for_docs(ranker=ranker, docs=ranked_docs, k_neighbors=doc_neighbors_k, device=device)
    input_ids = tokenizer(chat_str, return_tensors="pt").to(device)["input_ids"]
    layout = SdagLayout(L=input_ids.size(1), S=sys_user_len, spans=spans, qa_start=qa_start, neighbors=neighbors)
    bm = build_blockmap_checked(layout, device)
    return _generate(model, tokenizer, input_ids, bm, max_new_tokens)


def run_rag_with_doc_isolation_batch(model, tokenizer, queries: Sequence[str], malicious_docs_batch: Sequence[List[str]],
                                     retrieved_docs_batch: Sequence[list], max_new_tokens: int, device: str,
                                     add_attack_in_rank: int, ranker, doc_neighbors_k: int) -> List[str]:
    """SDAG answers for a batch of queries in one ragged pass: the reference calls ``run_rag_with_doc_isolation`` once
    per query (main.py:469-496); here every query keeps its own prompt, token offsets, block map, RoPE positions and
    cache length, and the batch shares the prefill launches and one CUDA-graph decode step per token.  Per query the
    result is what ``run_rag_with_doc_isolation`` returns (greedy; with sampling the draws differ: the reference
    consumes the global torch RNG one query at a time).  Models the fused engine cannot run, and layouts the span
    path cannot express, fall back to the per-query call."""
    from .llama_runner import get_runner

    if torch.device(device).type != "cuda":
        raise _lib.SdagB200Error("sdag_b200 runs on CUDA devices only (no CPU fallback)")
    queries = list(queries)
    prepare_model(model)
    runner = get_runner(model) if USE_FUSED_PREFILL else None

    def one_by_one():
        return [run_rag_with_doc_isolation(model, tokenizer, q, list(m), list(r), max_new_tokens, device, add_attack_in_rank,
                                           ranker, doc_neighbors_k)
                for q, m, r in zip(queries, malicious_docs_batch, retrieved_docs_batch)]

    if runner is None or not queries:
        return one_by_one()
    rows, layouts = [], []
    for q, mal, ret in zip(queries, malicious_docs_batch, retrieved_docs_batch):
        chat_str, S, spans, qa_start, ranked = build_rag_chat_and_spans(
            tokenizer=tokenizer, system_prompt=SYSTEM_PROMPT_RAG, user_template=USER_RAG_PROMPT, query=q,
            malicious_docs=list(mal), retrieved_docs=list(ret), add_attack_in_rank=add_attack_in_rank)
        nbrs = None
        if doc_neighbors_k and doc_neighbors_k > 0:
            nbrs = compute_doc_knn_for_docs(ranker=ranker, docs=ranked, k_neighbors=doc_neighbors_k, device=device)
        ids = tokenizer(chat_str, return_tensors="pt")["input_ids"][0]
        rows.append(ids)
        layouts.append(SdagLayout(L=int(ids.numel()), S=S, spans=spans, qa_start=qa_start, neighbors=nbrs))
    if any(len(l.spans) > _lib.MAX_DOCS or l.nbr_k > _lib.MAX_NBR_K for l in layouts):
        return one_by_one()
    bm = build_blockmap(layouts, device)
    if not bm.is_ok():  # some layout is outside the span builder's canonical form: dense-mask path per query
        return one_by_one()
    eos = tokenizer.eos_token_id
    lengths = torch.tensor([l.L for l in layouts], dtype=torch.int64, device=device)
    ids = torch.full((len(rows), int(lengths.max())), eos if eos is not None else 0, dtype=torch.int64)
    for b, r in enumerate(rows):
        ids[b, : r.numel()] = r
    temperature = float(TEMPERATURE) if TEMPERATURE > 0 else 0.0
    gen = runner.generate(ids.to(device), bm, max_new_tokens, eos, temperature=temperature, lengths=lengths)
    return [tokenizer.decode(_trim_at_eos(row, eos), skip_special_tokens=True).strip() for row in gen.tolist()]


# ------------------------------------------------------------------------------------------------------
# Batched prefill (serving-side extension of the batch-of-1 reference loop, main.py:467-496)
# ------------------------------------------------------------------------------------------------------
def prefill_batch(model, input_ids_host: torch.Tensor, layouts: Sequence[SdagLayout], device="cuda"):
    """First generated token of B prompts of equal length in one pass: ``input_ids_host`` int64 ``[B, L]`` in
    (pinned) host memory, one ``SdagLayout`` per prompt.  Copies the ids to the device, builds the block maps
    from the token offsets there, runs the SDAG prefill (last-position logits only) and returns
    ``(first_tokens_host [B] int64, past_key_values)``.  Greedy or sampled exactly like ``_pick``."""
    ids = input_ids_host.to(device, non_blocking=True)
    layouts = list(layouts)
    bm = build_blockmap(layouts, device) if all(len(l.spans) <= _lib.MAX_DOCS and l.nbr_k <= _lib.MAX_NBR_K for l in layouts) else None
    if bm is not None and bm.is_ok():
        logits, past = prefill_forward(model, ids, bm)
        nxt = _pick(logits)
        return nxt[:, 0].cpu(), past
    # some layout is not canonical (the reference still builds a mask for it): one prompt at a time through the
    # dense-mask block map, which represents any mask
    toks, pasts = [], []
    for b, lay in enumerate(layouts):
        logits, past = prefill_forward(model, ids[b : b + 1], build_blockmap_checked(lay, device))
        toks.append(_pick(logits)[:, 0])
        pasts.append(past)
    return torch.cat(toks).cpu(), pasts
