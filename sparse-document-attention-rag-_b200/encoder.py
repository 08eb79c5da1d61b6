"""Sentence encoder of the retrieval path, B200-native ("encode corpus").

The reference embeds every text with a ``SentenceTransformer`` (e5-large-v2: BERT-large, masked mean pooling,
L2 normalisation) through ``ranker.encode(texts, convert_to_tensor=True, normalize_embeddings=True)`` —
``embed_queries`` (src/pipeline/retrieval/dense.py:44-65), ``encode_texts_with_ranker``
(src/pipeline/attack/malicious_selection.py:12-34), ``compute_doc_knn_for_docs``
(src/pipeline/sparse_attention_RAG/SDAG.py:46-48).  ``B200Ranker`` is an object with that interface (``encode``,
``get_sentence_embedding_dimension``) that every function of this package — and of the reference — accepts as its
``ranker``:

* texts are tokenised, sorted by length and packed into right-padded batches of bounded token count (the
  sentence-transformers batching, minus the padding waste: padded rows never enter the attention);
* ``BertEncoderRunner`` runs the HF ``BertModel`` weights: one cuBLAS GEMM for q/k/v, the **prefill attention kernel**
  under a bidirectional block map (``sdag_build_blockmap_full``), fused residual + LayerNorm, exact GELU;
* ``enc_mean_pool_normalize`` pools and normalises on the device and writes fp32 rows (and their bf16 shadow)
  straight into the destination — ``encode_corpus`` fills the rows of a ``FlatIPIndex`` shard in place.

The arithmetic is bf16 with fp32 accumulation / statistics (the reference's encoder runs fp32): embeddings agree with
the fp32 oracle to ~1e-2 absolute (cosine > 0.999, tests/test_encoder_gpu.py) — a stated tolerance, not bit parity.
There is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib


def full_blockmap(lengths_dev: torch.Tensor, max_L: int):
    """Bidirectional block maps for a right-padded batch (``lengths_dev`` int32 ``[B]`` on the device)."""
    from .sdag import BlockMap

    lib = _lib.load()
    B = int(lengths_dev.numel())
    stride = int(lib.sdag_full_blockmap_bytes(max_L))
    buf = torch.empty((B, stride // 4), dtype=torch.int32, device=lengths_dev.device)
    _lib.count_launches(1)
    _lib.check(lib.sdag_build_blockmap_full(lengths_dev.data_ptr(), B, max_L, buf.data_ptr(), stride,
                                            _lib.stream_ptr(buf.device)), "sdag_build_blockmap_full")
    T = (max_L + _lib.SDAG_TILE - 1) // _lib.SDAG_TILE
    bm = BlockMap(buf, stride, T, B, max_L, "full")
    bm._lengths = lengths_dev  # keep alive until the kernel has run
    bm.no_uniform_rows = True  # every token of a sequence sees the whole sequence
    return bm


def _linear(x, w, bias, residual=None, epilogue: int = 0):
    """bf16 ``F.linear`` on this library's CTA-pair GEMM: plain (+ bias), + residual (``llm_gemm_residual``: the
    ``dense(x) + input_tensor`` of BertSelfOutput / BertOutput, bf16 roundings kept), or bias + erf-GELU (``epilogue=2``,
    BertIntermediate)."""
    lib = _lib.load()
    N, K = w.shape
    ok = (x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.is_contiguous() and N % 8 == 0 and K % 8 == 0
          and x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0
          and (residual is None or (residual.stride(1) == 1 and residual.stride(0) % 8 == 0)))
    if not ok:
        _lib.note_fallback("encoder._linear", "operand layout not accepted by llm_gemm: cuBLAS")
        y = F.linear(x, w, bias)
        if epilogue == 2:
            _lib.check(lib.enc_gelu(y.data_ptr(), y.data_ptr(), y.numel(), _lib.stream_ptr(y.device)), "enc_gelu")
        return y if residual is None else y + residual
    out = torch.empty((x.shape[0], N), dtype=torch.bfloat16, device=x.device)
    _lib.count_launches(1)
    if residual is None:
        _lib.check(lib.llm_gemm(x.data_ptr(), x.stride(0), w.data_ptr(), _lib.ptr(bias), out.data_ptr(), N, x.shape[0], N, K,
                                epilogue, _lib.stream_ptr(x.device)), "llm_gemm")
    else:
        _lib.check(lib.llm_gemm_residual(x.data_ptr(), x.stride(0), w.data_ptr(), _lib.ptr(bias), residual.data_ptr(),
                                         residual.stride(0), out.data_ptr(), N, x.shape[0], N, K, _lib.stream_ptr(x.device)),
                   "llm_gemm_residual")
    return out


def _layernorm(x, residual, weight, bias, eps):
    lib = _lib.load()
    out = torch.empty_like(x)
    _lib.count_launches(1)
    _lib.check(lib.enc_layernorm(x.data_ptr(), _lib.ptr(residual), weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                 x.shape[0], x.shape[1], float(eps), _lib.stream_ptr(x.device)), "enc_layernorm")
    return out


class BertEncoderRunner:
    """Forward of a Hugging Face ``BertModel`` (bf16, CUDA, absolute position embeddings, GELU) on this library's
    kernels.  ``try_build`` returns None for anything else."""

    def __init__(self, bert):
        cfg = bert.config
        self.bert = bert
        self.H = cfg.num_attention_heads
        self.hidden = cfg.hidden_size
        self.D = cfg.hidden_size // cfg.num_attention_heads
        self.layers = []
        for layer in bert.encoder.layer:
            sa, so = layer.attention.self, layer.attention.output
            wqkv = torch.cat([sa.query.weight.data, sa.key.weight.data, sa.value.weight.data], dim=0).contiguous()
            bqkv = torch.cat([sa.query.bias.data, sa.key.bias.data, sa.value.bias.data]).contiguous()
            h = self.hidden
            sa.query.weight.data, sa.key.weight.data, sa.value.weight.data = wqkv[:h], wqkv[h : 2 * h], wqkv[2 * h :]
            self.layers.append(dict(
                wqkv=wqkv, bqkv=bqkv, wo=so.dense.weight, bo=so.dense.bias, ln1_w=so.LayerNorm.weight, ln1_b=so.LayerNorm.bias,
                eps1=so.LayerNorm.eps, wi=layer.intermediate.dense.weight, bi=layer.intermediate.dense.bias,
                wd=layer.output.dense.weight, bd=layer.output.dense.bias, ln2_w=layer.output.LayerNorm.weight,
                ln2_b=layer.output.LayerNorm.bias, eps2=layer.output.LayerNorm.eps))

    @staticmethod
    def try_build(bert) -> Optional["BertEncoderRunner"]:
        try:
            cfg = bert.config
            p0 = next(bert.parameters())
            D = cfg.hidden_size // cfg.num_attention_heads
            ok = (p0.dtype == torch.bfloat16 and p0.is_cuda and cfg.hidden_act == "gelu" and D in (64, 128)
                  and getattr(cfg, "position_embedding_type", "absolute") == "absolute" and cfg.hidden_size % 8 == 0
                  and cfg.intermediate_size % 8 == 0 and hasattr(bert, "embeddings") and hasattr(bert, "encoder")
                  and not getattr(cfg, "is_decoder", False))
            return BertEncoderRunner(bert) if ok else None
        except Exception:
            return None

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, lengths: torch.Tensor, token_type_ids: Optional[torch.Tensor] = None):
        """ids int64 ``[B, L]`` right-padded, ``lengths`` int32 ``[B]`` (device) -> last hidden state ``[B, L, hidden]``
        (rows past a sequence's end hold finite don't-care values)."""
        from . import sdag

        lib = _lib.load()
        B, L = input_ids.shape
        T, H, D, hid = B * L, self.H, self.D, self.hidden
        dev = input_ids.device
        emb = self.bert.embeddings
        pos = torch.arange(L, device=dev)
        tt = token_type_ids if token_type_ids is not None else torch.zeros_like(input_ids)
        x = emb.word_embeddings(input_ids) + emb.token_type_embeddings(tt)
        x = (x + emb.position_embeddings(pos)[None]).view(T, hid)
        x = _layernorm(x, None, emb.LayerNorm.weight, emb.LayerNorm.bias, emb.LayerNorm.eps)
        bm = full_blockmap(lengths, L)
        scaling = float(D) ** -0.5
        # one attention output buffer for all layers, zeroed once: rows past a sequence's end are never written
        ctx_buf = torch.zeros((B, L, H, D), dtype=torch.bfloat16, device=dev)
        for ly in self.layers:
            qkv = _linear(x, ly["wqkv"], ly["bqkv"]).view(B, L, 3, H, D)
            # strided views straight into the fused projection: q [B,H,L,D], k / v [B,H,L,D] (no copies, TMA takes strides)
            q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
            ctx = sdag.sdag_prefill_attention(q, k, v, bm, scaling, ragged=True, out=ctx_buf).view(T, hid)
            del qkv
            # dense + bias + residual in the GEMM epilogue; LayerNorm then reads one tensor
            x = _layernorm(_linear(ctx, ly["wo"], ly["bo"], residual=x), None, ly["ln1_w"], ly["ln1_b"], ly["eps1"])
            inter = _linear(x, ly["wi"], ly["bi"], epilogue=2)  # bias + GELU epilogue (same roundings as the two-step form)
            x = _layernorm(_linear(inter, ly["wd"], ly["bd"], residual=x), None, ly["ln2_w"], ly["ln2_b"], ly["eps2"])
        return x.view(B, L, hid)


class B200Ranker:
    """Drop-in for the ``SentenceTransformer`` the reference passes around as ``ranker``.

    ``model``: HF ``BertModel`` (e5-large-v2 geometry; converted to bf16 on the device), ``tokenizer``: its tokenizer.
    ``max_seq_length`` 512 and mean pooling are the e5-large-v2 module stack.  ``max_batch_tokens``
    bounds the padded size of one forward batch."""

    def __init__(self, model, tokenizer, device="cuda", max_seq_length: int = 512, max_batch_tokens: int = 32768):
        if torch.device(device).type != "cuda":
            raise _lib.SdagB200Error("sdag_b200 runs on CUDA devices only (no CPU fallback)")
        _lib.load()
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.model = model.to(self.device, dtype=torch.bfloat16).eval()
        self.tokenizer = tokenizer
        self.max_seq_length = int(max_seq_length)
        self.max_batch_tokens = int(max_batch_tokens)
        self.runner = BertEncoderRunner.try_build(self.model)
        if self.runner is None:
            raise _lib.SdagB200Error("B200Ranker needs a BERT-family encoder (GELU, absolute positions, head_dim 64 or 128)")

    def get_sentence_embedding_dimension(self) -> int:
        return int(self.model.config.hidden_size)

    # -- batching ---------------------------------------------------------------------------------------
    def _tokenize(self, texts: Sequence[str]) -> List[List[int]]:
        enc = self.tokenizer(list(texts), padding=False, truncation=True, max_length=self.max_seq_length)
        return [list(ids) for ids in enc["input_ids"]]

    def _batches(self, token_lists: List[List[int]]):
        """Indices sorted by descending length (sentence-transformers order), cut into batches whose padded size
        stays below ``max_batch_tokens``."""
        order = sorted(range(len(token_lists)), key=lambda i: -len(token_lists[i]))
        batch, width = [], 0
        for i in order:
            n = max(len(token_lists[i]), 1)
            w = max(width, n)
            if batch and w * (len(batch) + 1) > self.max_batch_tokens:
                yield batch, width
                batch, w = [], n
            batch.append(i)
            width = w
        if batch:
            yield batch, width

    @torch.no_grad()
    def encode_into(self, texts: Sequence[str], out_f32: torch.Tensor, out_bf16: Optional[torch.Tensor] = None, row0: int = 0,
                    normalize_embeddings: bool = True) -> None:
        """Embeddings of ``texts`` written to rows ``row0 + i`` of ``out_f32`` (``[>= row0 + n, dim]`` fp32 on the
        device) and, if given, of its bf16 shadow — in input order, whatever order the batches run in."""
        self.encode_tokens_into(self._tokenize(texts), out_f32, out_bf16, row0, normalize_embeddings)

    @torch.no_grad()
    def encode_tokens_into(self, toks: List[List[int]], out_f32: torch.Tensor, out_bf16: Optional[torch.Tensor] = None,
                           row0: int = 0, normalize_embeddings: bool = True) -> None:
        """Same, from token id lists (already truncated to ``max_seq_length``)."""
        lib = _lib.load()
        dim = self.get_sentence_embedding_dimension()
        if out_f32.dtype != torch.float32 or out_f32.shape[1] != dim or not out_f32.is_contiguous() or out_f32.device != self.device:
            raise ValueError("out_f32 must be a contiguous fp32 [rows, dim] tensor on the ranker's device")
        if out_f32.shape[0] < row0 + len(toks):
            raise ValueError("destination has too few rows")
        pad = self.tokenizer.pad_token_id if self.tokenizer.pad_token_id is not None else 0
        lens_all = np.fromiter((len(t) for t in toks), dtype=np.int64, count=len(toks))
        for idx, width in self._batches(toks):
            ids_np = np.full((len(idx), width), pad, dtype=np.int64)
            for r, i in enumerate(idx):
                ids_np[r, : lens_all[i]] = toks[i]
            idx_np = np.asarray(idx, dtype=np.int64)
            lengths = torch.from_numpy(lens_all[idx_np].astype(np.int32)).to(self.device, non_blocking=True)
            rows = torch.from_numpy(idx_np + row0).to(self.device, non_blocking=True)
            hidden = self.runner.forward(torch.from_numpy(ids_np).to(self.device, non_blocking=True), lengths)
            _lib.count_launches(1)
            _lib.check(lib.enc_mean_pool_normalize(hidden.data_ptr(), hidden.stride(0), hidden.stride(1), lengths.data_ptr(),
                                                   len(idx), width, dim, rows.data_ptr(), 0, out_f32.data_ptr(),
                                                   _lib.ptr(out_bf16), 1 if normalize_embeddings else 0,
                                                   _lib.stream_ptr(self.device)), "enc_mean_pool_normalize")

    # -- SentenceTransformer surface ------------------------------------------------------------------------
    def encode(self, sentences, batch_size: int = 32, convert_to_tensor: bool = False, normalize_embeddings: bool = False,
               convert_to_numpy: bool = True, show_progress_bar: bool = False, **_):
        """``SentenceTransformer.encode`` for the arguments the reference uses; ``batch_size`` is accepted and ignored
        (batches are cut by token count).  Returns a float32 tensor on the device (``convert_to_tensor=True``) or a
        numpy array."""
        single = isinstance(sentences, str)
        texts = [sentences] if single else list(sentences)
        dim = self.get_sentence_embedding_dimension()
        out = torch.empty((len(texts), dim), dtype=torch.float32, device=self.device)
        if texts:
            self.encode_into(texts, out, normalize_embeddings=normalize_embeddings)
        if single:
            out = out[0]
        return out if convert_to_tensor else out.cpu().numpy()


def encode_corpus(ranker: B200Ranker, texts: Sequence[str], prefix: str = "passage: ", chunk: int = 8192):
    """"Encode corpus" (north star): a ``FlatIPIndex`` over the unit-norm embeddings of ``texts`` (e5 passage prefix,
    malicious_selection.py:25-29), with the pooled rows written by the encoder's last kernel directly into the index's
    fp32 rows and bf16 shadow — no host round trip, no second cast pass.  The host tokenizer of chunk i+1 runs on a
    worker thread (fast tokenizers release the GIL) while the GPU encodes chunk i."""
    from concurrent.futures import ThreadPoolExecutor

    from .retrieval import FlatIPIndex

    n, dim = len(texts), ranker.get_sentence_embedding_dimension()
    f32 = torch.empty((n, dim), dtype=torch.float32, device=ranker.device)
    b16 = torch.empty((n, dim), dtype=torch.bfloat16, device=ranker.device)
    starts = list(range(0, n, chunk))

    def tokenize(r0):
        return ranker._tokenize([prefix + t for t in texts[r0 : r0 + chunk]])

    with ThreadPoolExecutor(max_workers=1) as pool:
        pending = pool.submit(tokenize, starts[0]) if starts else None
        for k, r0 in enumerate(starts):
            toks = pending.result()
            pending = pool.submit(tokenize, starts[k + 1]) if k + 1 < len(starts) else None
            ranker.encode_tokens_into(toks, f32, b16, row0=r0)
    return FlatIPIndex.from_encoded(f32, b16)
