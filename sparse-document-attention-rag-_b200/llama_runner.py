"""Fused prefill forward for Llama-family decoders (path A4: the forward that ``generate_with_custom_mask`` drives,
SDAG.py:175-176).

Hugging Face's decoder layer runs ~25 small torch kernels around each attention call (RMSNorm in fp32 with six
passes, RoPE via slice / neg / cat, SiLU then mul, residual adds) — about a third of a Llama-3-8B prefill step on a
B200.  ``FusedPrefill`` keeps the HF module's weights (the q/k/v and gate/up projections are re-pointed at views of
fused weight matrices, no copy is kept) and drives, per layer: ``llm_gemm_rope`` (q/k/v projection with RoPE and the
KV-cache write in its epilogue), the SDAG attention kernel, ``llm_gemm_residual`` (o_proj + residual add),
``llm_rmsnorm``, ``llm_gemm`` with the SwiGLU epilogue (gate + up), ``llm_gemm_residual`` (down_proj + residual add),
``llm_rmsnorm`` — no library GEMM in the prefill.  Only the last position of each prompt goes through the final norm
and the LM head (the reference computes all L rows and uses one, SDAG.py:181).  The returned KV cache is a HF ``DynamicCache``,
so decoding continues through the normal HF path (with the ``sdag_b200`` decode attention).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn.functional as F

from . import _lib


class _GemmTimer:
    """bench.py: CUDA-event timing of one prefill GEMM launch on its stream, collected into ``_gemm_events`` as
    ``(start, end, name, rows, N, K)`` — inactive (no events recorded) unless the list is set."""

    def __init__(self, name: str, rows: int, N: int, K: int):
        self.meta = (name, rows, N, K)
        self.ev = None

    def __enter__(self):
        if _gemm_events is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            _gemm_events.append(self.ev + self.meta)
        return False


def _gemm_ok(x, w) -> bool:
    """Shapes / strides llm_gemm takes: 16-byte rows everywhere (K, N and the row strides multiples of 8 elements)."""
    return (x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.is_contiguous() and w.shape[0] % 8 == 0
            and w.shape[1] % 8 == 0 and x.stride(1) == 1 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0)


def _on_cublas(name: str) -> bool:
    """A/B knob: SDAG_CUBLAS_PROJ=o,down,... sends the named prefill projections back to cuBLAS (+ stand-alone glue)."""
    return name.split("+")[0] in CUBLAS_PROJ


def _linear(x, w, bias, residual=None, prefill: bool = True, name: str = "linear"):
    """``F.linear(x, w, bias)`` (+ ``residual``, added with the bf16 roundings of ``residual + linear(x)``) for bf16
    ``x [rows, K]``.  The prefill runs every projection on this library's CTA-pair tcgen05 GEMM (``llm_gemm`` /
    ``llm_gemm_residual``); the few-row GEMMs of a decode step with more than two prompts stay on cuBLAS, where a
    256-row tile would be almost empty (``prefill=False``)."""
    if (prefill and USE_PAIR_GEMM and not _on_cublas(name) and _gemm_ok(x, w)
            and (residual is None or (residual.stride(1) == 1 and residual.stride(0) % 8 == 0))):
        lib = _lib.load()
        N, K = w.shape
        out = torch.empty((x.shape[0], N), dtype=torch.bfloat16, device=x.device)
        _lib.count_launches(1)
        with _GemmTimer(name, x.shape[0], N, K):
            if residual is None:
                _lib.check(lib.llm_gemm(x.data_ptr(), x.stride(0), w.data_ptr(), _lib.ptr(bias), out.data_ptr(), N, x.shape[0], N, K,
                                        0, _lib.stream_ptr(x.device)), "llm_gemm")
            else:
                _lib.check(lib.llm_gemm_residual(x.data_ptr(), x.stride(0), w.data_ptr(), _lib.ptr(bias), residual.data_ptr(),
                                                 residual.stride(0), out.data_ptr(), N, x.shape[0], N, K,
                                                 _lib.stream_ptr(x.device)), "llm_gemm_residual")
        return out
    if prefill and USE_PAIR_GEMM and not _on_cublas(name):
        _lib.note_fallback("llama_runner._linear", "operand layout not accepted by llm_gemm: cuBLAS")
    with _GemmTimer(name + " (cuBLAS)", x.shape[0], w.shape[0], w.shape[1]):
        y = F.linear(x, w, bias)
    return y if residual is None else y + residual


def _rmsnorm(x, residual, weight, eps):
    """(new residual stream | None, normalised rows) for bf16 [rows, dim]."""
    lib = _lib.load()
    rows, dim = x.shape
    out = torch.empty_like(x)
    sum_out = torch.empty_like(x) if residual is not None else None
    _lib.count_launches(1)
    _lib.check(lib.llm_rmsnorm(x.data_ptr(), _lib.ptr(residual), weight.data_ptr(), _lib.ptr(sum_out), out.data_ptr(), rows, dim,
                               float(eps), _lib.stream_ptr(x.device)), "llm_rmsnorm")
    return sum_out, out


class FusedPrefill:
    """Prefill engine for one HF causal LM; build with ``FusedPrefill.try_build(model)`` (None if unsupported)."""

    def __init__(self, model):
        cfg = model.config
        self.model = model
        self.H = cfg.num_attention_heads
        self.Hkv = cfg.num_key_value_heads
        self.D = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        self.hidden = cfg.hidden_size
        self.inter = cfg.intermediate_size
        self.layers = []
        for layer in model.model.layers:
            at, mlp = layer.self_attn, layer.mlp
            wqkv = torch.cat([at.q_proj.weight.data, at.k_proj.weight.data, at.v_proj.weight.data], dim=0).contiguous()
            nq, nk = at.q_proj.weight.shape[0], at.k_proj.weight.shape[0]
            at.q_proj.weight.data, at.k_proj.weight.data, at.v_proj.weight.data = wqkv[:nq], wqkv[nq : nq + nk], wqkv[nq + nk :]
            bqkv = None
            if at.q_proj.bias is not None:
                bqkv = torch.cat([at.q_proj.bias.data, at.k_proj.bias.data, at.v_proj.bias.data]).contiguous()
            wgu = torch.cat([mlp.gate_proj.weight.data, mlp.up_proj.weight.data], dim=0).contiguous()
            mlp.gate_proj.weight.data, mlp.up_proj.weight.data = wgu[: self.inter], wgu[self.inter :]
            self.layers.append(dict(
                wqkv=wqkv, bqkv=bqkv, wo=at.o_proj.weight, bo=at.o_proj.bias, wgu=wgu, wdown=mlp.down_proj.weight,
                w_in=layer.input_layernorm.weight, eps_in=layer.input_layernorm.variance_epsilon,
                w_post=layer.post_attention_layernorm.weight, eps_post=layer.post_attention_layernorm.variance_epsilon,
                scaling=at.scaling))
        torch.cuda.empty_cache()

    @staticmethod
    def unsupported_reason(model) -> Optional[str]:
        """None when the fused engine covers this model, else why not (architecture / dtype / device)."""
        cfg = model.config
        inner = getattr(model, "model", None)
        try:
            p0 = next(model.parameters())
        except StopIteration:
            return "model has no parameters"
        if p0.dtype != torch.bfloat16:
            return f"weights are {p0.dtype}, the engine computes in bf16"
        if not p0.is_cuda:
            return "weights are not on a CUDA device"
        if getattr(cfg, "hidden_act", "") != "silu":
            return f"hidden_act={getattr(cfg, 'hidden_act', None)!r} (SwiGLU MLP expected)"
        if inner is None or not (hasattr(inner, "rotary_emb") and hasattr(inner, "embed_tokens") and hasattr(model, "lm_head")):
            return "not a Llama-shaped decoder (rotary_emb / embed_tokens / lm_head)"
        if cfg.hidden_size % 8 or cfg.intermediate_size % 8:
            return "hidden / intermediate size not a multiple of 8"
        D = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
        if D not in (64, 128):
            return f"head_dim={D} (64 or 128 supported)"
        for layer in inner.layers:
            at, mlp = layer.self_attn, layer.mlp
            for name in ("q_proj", "k_proj", "v_proj", "o_proj"):
                if not isinstance(getattr(at, name, None), torch.nn.Linear):
                    return f"self_attn.{name} is not nn.Linear"
            for name in ("gate_proj", "up_proj", "down_proj"):
                lin = getattr(mlp, name, None)
                if not isinstance(lin, torch.nn.Linear) or lin.bias is not None:
                    return f"mlp.{name} is not a bias-free nn.Linear"
            if not (hasattr(layer.input_layernorm, "variance_epsilon") and hasattr(layer.post_attention_layernorm, "variance_epsilon")):
                return "layer norms are not RMSNorm"
            if hasattr(at, "q_norm"):
                return "per-head q/k norms (Qwen3, Gemma) are not handled"
        return None

    @staticmethod
    def try_build(model) -> Optional["FusedPrefill"]:
        """The engine for ``model``, or None (Hugging Face forward + attention hook) when the architecture is not covered;
        the reason is recorded through ``_lib.note_fallback`` (an error under SDAG_STRICT=1).  Errors while building the
        engine for a covered model are not swallowed."""
        reason = FusedPrefill.unsupported_reason(model)
        if reason is not None:
            _lib.note_fallback("FusedPrefill.try_build", reason)
            return None
        return FusedPrefill(model)

    @property
    def path(self) -> str:
        """Which kernels the prefill of this engine launches (bench.py prints it next to the numbers)."""
        if USE_PAIR_GEMM:
            qkv = "llm_gemm_rope (q/k/v + RoPE + KV write)" if FUSED_ROPE_GEMM else "llm_gemm + llm_rope_qkv"
            lin = f"{qkv}, llm_gemm_residual (o, down), llm_gemm SwiGLU (gate/up), llm_gemm (lm_head)"
        else:
            lin = "cuBLAS via torch (SDAG_PAIR_GEMM=off)"
        return f"FusedPrefill: {lin}; sdag_prefill_kernel; llm_rmsnorm"

    def rope_tables(self, n_pos: int, device):
        """bf16 cos / sin tables ``[n_pos, D]`` indexed by absolute position (HF rotary module, rounded like HF does)."""
        inner = self.model.model
        dummy = torch.empty((1, 1, self.hidden), dtype=torch.bfloat16, device=device)
        cos, sin = inner.rotary_emb(dummy, torch.arange(n_pos, device=device).unsqueeze(0))
        return cos[0].contiguous(), sin[0].contiguous()

    @torch.no_grad()
    def prefill(self, input_ids: torch.Tensor, blockmap, want_cache: bool = True, kv=None, lengths=None):
        """input_ids int64 [B, L] on the model's device -> (last-position logits [B, vocab], DynamicCache | None).

        ``lengths`` (int64 ``[B]`` on the device) marks a ragged batch: prompt b occupies columns ``[0, lengths[b])``
        (right-padded), its block map was built for that length, and its logits are taken at ``lengths[b] - 1``.
        Rows past a prompt's end are computed but never read (every op between attention calls is row-wise).

        ``kv = (k_cache, v_cache, cos, sin)`` with caches ``[n_layers, B, Hkv, S_cap, D]`` makes the prefill write K/V
        straight into a preallocated cache (the decode engine's) instead of building a ``DynamicCache``."""
        from . import sdag  # late import (sdag imports this module lazily too)

        lib = _lib.load()
        model, inner = self.model, self.model.model
        B, L = input_ids.shape
        T = B * L
        dev = input_ids.device
        H, Hkv, D = self.H, self.Hkv, self.D
        stream = _lib.stream_ptr(dev)
        h = inner.embed_tokens(input_ids)
        if kv is not None:
            kc_all, vc_all, cos, sin = kv
        else:
            cos, sin = self.rope_tables(L, dev)
        resid = h.view(T, self.hidden)
        _, normed = _rmsnorm(resid, None, self.layers[0]["w_in"], self.layers[0]["eps_in"])
        cache = None
        if want_cache and kv is None:
            from transformers import DynamicCache

            cache = DynamicCache(config=model.config) if _accepts_config(DynamicCache) else DynamicCache()
        last_rows = torch.arange(B, device=dev) * L + ((lengths - 1) if lengths is not None else (L - 1))
        n_layers = len(self.layers)
        per_tile = 256 // D
        fused_qkv = (FUSED_ROPE_GEMM and USE_PAIR_GEMM and not _on_cublas("qkv") and H % per_tile == 0 and Hkv % per_tile == 0 and self.hidden % 8 == 0)
        for i, ly in enumerate(self.layers):
            q = torch.empty((B, L, H, D), dtype=torch.bfloat16, device=dev)
            if kv is not None:
                kbuf, vbuf = kc_all[i], vc_all[i]
                k, v = kbuf[:, :, :L], vbuf[:, :, :L]
            else:
                kbuf = k = torch.empty((B, Hkv, L, D), dtype=torch.bfloat16, device=dev)
                vbuf = v = torch.empty((B, Hkv, L, D), dtype=torch.bfloat16, device=dev)
            if fused_qkv:  # q/k/v projection on the CTA-pair GEMM with RoPE + cache write in its epilogue
                _lib.count_launches(1)
                with _GemmTimer("qkv+rope", T, (H + 2 * Hkv) * D, self.hidden):
                    _lib.check(lib.llm_gemm_rope(normed.data_ptr(), normed.stride(0), ly["wqkv"].data_ptr(), _lib.ptr(ly["bqkv"]),
                                                 cos.data_ptr(), sin.data_ptr(), q.data_ptr(), kbuf.data_ptr(), vbuf.data_ptr(), B, L,
                                                 H, Hkv, D, self.hidden, kbuf.stride(0), kbuf.stride(1), 0, stream), "llm_gemm_rope")
            else:
                qkv = _linear(normed, ly["wqkv"], ly["bqkv"], name="qkv")
                _lib.count_launches(1)
                _lib.check(lib.llm_rope_qkv(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), q.data_ptr(), kbuf.data_ptr(),
                                            vbuf.data_ptr(), B, L, H, Hkv, D, kbuf.stride(0), kbuf.stride(1), 0, 0, 0, stream),
                           "llm_rope_qkv")
                del qkv
            if cache is not None:
                _cache_adopt(cache, k, v, i)
            attn = sdag.sdag_prefill_attention(q.transpose(1, 2), k, v, blockmap, ly["scaling"],
                                               ragged=lengths is not None)  # [B, L, H, D]
            attn = attn.view(T, H * D)
            last = i == n_layers - 1
            if last:  # only the last position of each prompt is read after the final layer
                attn, resid = attn[last_rows], resid[last_rows]
            normed, resid = self._mlp_block(attn, resid, i)
        logits = _linear(normed, model.lm_head.weight, model.lm_head.bias, name="lm_head")
        return logits, cache

    def _mlp_block(self, attn_rows, resid, i, prefill: bool = True):
        """o_proj + residual + post-norm + SwiGLU MLP + residual + the next layer's (or the final) norm.  In the prefill
        both residual adds happen in the epilogues of the o / down GEMMs (``llm_gemm_residual``), so each RMSNorm reads one
        tensor and writes one."""
        lib = _lib.load()
        ly = self.layers[i]
        inner = self.model.model
        if prefill and USE_PAIR_GEMM and not _on_cublas("o"):
            resid = _linear(attn_rows, ly["wo"], ly["bo"], residual=resid, name="o+residual")
            _, normed = _rmsnorm(resid, None, ly["w_post"], ly["eps_post"])
        else:
            o = _linear(attn_rows, ly["wo"], ly["bo"], prefill=prefill, name="o")
            resid, normed = _rmsnorm(o, resid, ly["w_post"], ly["eps_post"])
        rows = normed.shape[0]
        act = torch.empty((rows, self.inter), dtype=torch.bfloat16, device=normed.device)
        _lib.count_launches(1)
        if (prefill or rows >= PAIR_GEMM_MIN_ROWS) and USE_PAIR_GEMM and not _on_cublas("gate_up") and self.inter % 8 == 0:
            # gate/up GEMM on CTA pairs with the SwiGLU in its epilogue: the [rows, 2*inter] intermediate is never written
            with _GemmTimer("gate_up+swiglu", rows, 2 * self.inter, self.hidden):
                _lib.check(lib.llm_gemm(normed.data_ptr(), normed.stride(0), ly["wgu"].data_ptr(), 0, act.data_ptr(), self.inter,
                                        rows, 2 * self.inter, self.hidden, 1, _lib.stream_ptr(normed.device)), "llm_gemm")
        else:
            gu = F.linear(normed, ly["wgu"])
            _lib.check(lib.llm_swiglu(gu.data_ptr(), act.data_ptr(), rows, self.inter, _lib.stream_ptr(gu.device)), "llm_swiglu")
        nw, ne = ((self.layers[i + 1]["w_in"], self.layers[i + 1]["eps_in"]) if i + 1 < len(self.layers)
                  else (inner.norm.weight, inner.norm.variance_epsilon))
        if prefill and USE_PAIR_GEMM and not _on_cublas("down"):
            resid = _linear(act, ly["wdown"], None, residual=resid, name="down+residual")
            _, normed = _rmsnorm(resid, None, nw, ne)
        else:
            d = _linear(act, ly["wdown"], None, prefill=prefill, name="down")
            resid, normed = _rmsnorm(d, resid, nw, ne)
        return normed, resid

    # ------------------------------------------------------------------------------------------------
    # generation: prefill into a preallocated cache, then a CUDA-graph-captured decode step replayed per token
    # ------------------------------------------------------------------------------------------------
    def _decode_context(self, B: int, need_cap: int, temperature: float, dev, use_graph: bool, nucleus: bool = False):
        """Preallocated KV cache + device-side step state + captured decode graph, cached per (batch, capacity bucket,
        sampling mode).  Capturing costs a few hundred ms, so it is done once per bucket and replayed across calls."""
        from . import sdag

        lib = _lib.load()
        S_cap = max(2048, (need_cap + 2047) // 2048 * 2048)
        key = (B, S_cap, temperature > 0, nucleus, use_graph, USE_FUSED_GEMV, FUSED_GEMV_MAX_B, tuple(sorted(DECODE_GEMV.items())), str(dev))
        ctxs = self.__dict__.setdefault("_ctx", {})
        ctx = ctxs.get(key)
        if ctx is not None:
            return ctx
        # keep HBM bounded: drop the oldest contexts once their KV caches exceed the budget.  (A fixed limit of two made a loop
        # that alternates greedy SDAG and sampled CARG batches around a capacity-bucket boundary re-capture its graphs again
        # and again: seconds per hundred queries.)
        n_layers_ = len(self.layers)
        new_bytes = 2 * n_layers_ * B * self.Hkv * S_cap * self.D * 2
        while ctxs and sum(c_["kc"].numel() * 4 for c_ in ctxs.values()) + new_bytes > DECODE_CTX_BUDGET_BYTES:
            ctxs.pop(next(iter(ctxs)))
            torch.cuda.empty_cache()
        model, inner = self.model, self.model.model
        H, Hkv, D = self.H, self.Hkv, self.D
        n_layers = len(self.layers)
        c = dict(S_cap=S_cap)
        c["kc"] = torch.zeros((n_layers, B, Hkv, S_cap, D), dtype=torch.bfloat16, device=dev)
        c["vc"] = torch.zeros_like(c["kc"])
        c["cos"], c["sin"] = self.rope_tables(S_cap, dev)
        c["gen"] = torch.zeros((B, S_cap), dtype=torch.int64, device=dev)
        c["tok"] = torch.zeros((B, 1), dtype=torch.int64, device=dev)
        # per prompt (ragged batches): position of the token being fed / cache length once its K and V are written
        c["pos"] = torch.zeros((B,), dtype=torch.int32, device=dev)
        c["len"] = torch.ones((B,), dtype=torch.int32, device=dev)
        c["slot"] = torch.ones((B, 1), dtype=torch.int64, device=dev)  # output column
        c["temp"] = torch.ones((1,), dtype=torch.float32, device=dev)
        c["top_p"] = torch.ones((1,), dtype=torch.float32, device=dev)
        q = torch.empty((B, 1, H, D), dtype=torch.bfloat16, device=dev)
        attn_out = torch.empty((B, 1, H, D), dtype=torch.bfloat16, device=dev)
        # split-KV scratch owned by this context: the captured graph bakes its address in, so it must live exactly as
        # long as the graph (the shared per-device scratch of sdag.py is replaced when a larger one is needed)
        c["dec_ws"] = dec_ws = torch.empty(sdag.decode_workspace_bytes(B, H, D, S_cap), dtype=torch.uint8, device=dev)
        kc, vc, cos, sin = c["kc"], c["vc"], c["cos"], c["sin"]
        sample = temperature > 0

        def step():
            h = inner.embed_tokens(c["tok"]).view(B, self.hidden)
            resid = h
            _, normed = _rmsnorm(resid, None, self.layers[0]["w_in"], self.layers[0]["eps_in"])
            for i, ly in enumerate(self.layers):
                qkv = F.linear(normed, ly["wqkv"], ly["bqkv"])
                _lib.count_launches(1)
                _lib.check(lib.llm_rope_qkv(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), q.data_ptr(), kc[i].data_ptr(),
                                            vc[i].data_ptr(), B, 1, H, Hkv, D, kc[i].stride(0), kc[i].stride(1), 0,
                                            c["pos"].data_ptr(), 1, _lib.stream_ptr(dev)), "llm_rope_qkv")
                sdag.sdag_decode_attention(q.transpose(1, 2), kc[i], vc[i], ly["scaling"], kv_len_dev=c["len"], out=attn_out,
                                           workspace=dec_ws)
                normed, resid = self._mlp_block(attn_out.view(B, H * D), resid, i, prefill=False)
            lg = model.lm_head(normed)
            if sample:
                nxt = sample_tokens(lg, c["temp"], c["top_p"] if nucleus else None)
            else:
                nxt = torch.argmax(lg, dim=-1, keepdim=True)
            c["gen"].scatter_(1, c["slot"], nxt)
            c["tok"].copy_(nxt)
            c["pos"].add_(1)
            c["len"].add_(1)
            c["slot"].add_(1)

        # B <= 2: the linear layers of the step run on llm_gemv — weights streamed once, RMSNorm applied to the staged
        # input, SwiGLU and the residual add in the epilogues — 7 launches per layer instead of 10 (no norm / activation
        # kernels).  CUDA-core arithmetic: beyond two prompts its unpack + FMA work per weight byte catches up with the
        # HBM stream (tools/gemv_bench.py) and cuBLAS is the faster path again.
        inter, hid = self.inter, self.hidden
        head_w = model.lm_head.weight
        fused_ok = (USE_FUSED_GEMV and B <= FUSED_GEMV_MAX_B and hid % 8 == 0 and inter % 8 == 0 and (H * D) % 8 == 0
                    and all(n % 4 == 0 for n in ((H + 2 * Hkv) * D, hid, 2 * inter, head_w.shape[0]))
                    and model.lm_head.bias is None and head_w.dtype == torch.bfloat16 and B * inter * 2 <= 200 * 1024)
        if fused_ok:
            qkv_b = torch.empty((B, (H + 2 * Hkv) * D), dtype=torch.bfloat16, device=dev)
            gu_b = torch.empty((B, inter), dtype=torch.bfloat16, device=dev)  # SwiGLU activation (gate|up never materialised)
            res_a = torch.empty((B, hid), dtype=torch.bfloat16, device=dev)
            res_b = torch.empty((B, hid), dtype=torch.bfloat16, device=dev)
            logits_b = torch.empty((B, head_w.shape[0]), dtype=torch.bfloat16, device=dev)
            st = _lib.stream_ptr(dev)

            def gemv(x, w, bias, norm_w, resid, out, K, N, pro, eps=0.0, pair=0):
                _lib.count_launches(1)
                _lib.check(lib.llm_gemv(x.data_ptr(), w.data_ptr(), _lib.ptr(bias), _lib.ptr(norm_w), _lib.ptr(resid), out.data_ptr(),
                                        B, K, N, pro, float(eps), pair, _lib.stream_ptr(dev)), "llm_gemv")

            pf_stream = torch.cuda.Stream(device=dev) if PREFETCH_O_PROJ else None

            def step_fused():
                resid = inner.embed_tokens(c["tok"]).view(B, hid)
                main = torch.cuda.current_stream(dev)
                for i, ly in enumerate(self.layers):
                    if FUSED_ROPE_GEMV:  # q/k/v projection with RMSNorm prologue, RoPE + cache write epilogue: one launch
                        _lib.count_launches(1)
                        _lib.check(lib.llm_gemv_rope(resid.data_ptr(), ly["wqkv"].data_ptr(), _lib.ptr(ly["bqkv"]), ly["w_in"].data_ptr(),
                                                     float(ly["eps_in"]), cos.data_ptr(), sin.data_ptr(), q.data_ptr(), kc[i].data_ptr(),
                                                     vc[i].data_ptr(), B, hid, H, Hkv, D, kc[i].stride(0), kc[i].stride(1), 0,
                                                     c["pos"].data_ptr(), 1, _lib.stream_ptr(dev)), "llm_gemv_rope")
                    else:
                        gemv(resid, ly["wqkv"], ly["bqkv"], ly["w_in"], None, qkv_b, hid, qkv_b.shape[1], 1, ly["eps_in"])
                        _lib.count_launches(1)
                        _lib.check(lib.llm_rope_qkv(qkv_b.data_ptr(), cos.data_ptr(), sin.data_ptr(), q.data_ptr(), kc[i].data_ptr(),
                                                    vc[i].data_ptr(), B, 1, H, Hkv, D, kc[i].stride(0), kc[i].stride(1), 0,
                                                    c["pos"].data_ptr(), 1, _lib.stream_ptr(dev)), "llm_rope_qkv")
                    if pf_stream is not None:
                        # HBM idles while the attention kernels run: pull the o_proj weights into L2 on a side stream (forked
                        # after the q/k/v GEMV; joined at the end of the step so the main chain keeps its dependent launches)
                        pf_stream.wait_stream(main)
                        _lib.count_launches(1)
                        _lib.check(lib.llm_prefetch_l2(ly["wo"].data_ptr(), ly["wo"].numel() * 2, pf_stream.cuda_stream), "llm_prefetch_l2")
                    sdag.sdag_decode_attention(q.transpose(1, 2), kc[i], vc[i], ly["scaling"], kv_len_dev=c["len"], out=attn_out,
                                           workspace=dec_ws)
                    gemv(attn_out, ly["wo"], ly["bo"], None, resid, res_a, H * D, hid, 0)
                    gemv(res_a, ly["wgu"], None, ly["w_post"], None, gu_b, hid, 2 * inter, 1, ly["eps_post"], pair=1)  # -> act [B, inter]
                    gemv(gu_b, ly["wdown"], None, None, res_a, res_b, inter, hid, 0)
                    resid = res_b
                gemv(resid, head_w, None, inner.norm.weight, None, logits_b, hid, head_w.shape[0], 1, inner.norm.variance_epsilon)
                lg = logits_b
                if sample:
                    nxt = sample_tokens(lg, c["temp"], c["top_p"] if nucleus else None)
                else:
                    nxt = torch.argmax(lg, dim=-1, keepdim=True)
                c["gen"].scatter_(1, c["slot"], nxt)
                c["tok"].copy_(nxt)
                c["pos"].add_(1)
                c["len"].add_(1)
                c["slot"].add_(1)
                if pf_stream is not None:
                    main.wait_stream(pf_stream)

            step = step_fused
        # 3 <= B <= 8: every linear layer named in DECODE_GEMV runs on llm_gemv's warp-MMA form (weights streamed once,
        # RMSNorm prologue / SwiGLU / residual epilogues fused), the others on cuBLAS + the stand-alone glue kernels.
        # The residual stream is (x + pending): a cuBLAS o / down output stays pending until the next RMSNorm adds it.
        mixed = DECODE_GEMV if (USE_FUSED_GEMV and 3 <= B <= 8 and not fused_ok and hid % 32 == 0 and inter % 32 == 0
                                and (H * D) % 32 == 0 and ((H + 2 * Hkv) * D) % 16 == 0 and hid % 16 == 0
                                and head_w.dtype == torch.bfloat16) else {}
        if "head" in mixed and head_w.shape[0] % (8 if mixed["head"] == "skinny" else 16):
            mixed = {k_: v_ for k_, v_ in mixed.items() if k_ != "head"}  # vocabulary size not a multiple the kernels take: cuBLAS
        if mixed:
            n_qkv = (H + 2 * Hkv) * D
            qkv_b = torch.empty((B, n_qkv), dtype=torch.bfloat16, device=dev)
            gu_b = torch.empty((B, inter), dtype=torch.bfloat16, device=dev)
            res_a = torch.empty((B, hid), dtype=torch.bfloat16, device=dev)
            res_b = torch.empty((B, hid), dtype=torch.bfloat16, device=dev)
            logits_b = torch.empty((B, head_w.shape[0]), dtype=torch.bfloat16, device=dev)

            def gemv_for(name):
                """llm_gemv (warp-MMA form, weights through registers) or, for ``name=skinny``, llm_skinny_linear (TMA ring +
                tcgen05, K split over a cluster for the narrow layers)."""
                skinny = mixed.get(name) == "skinny"

                def run(x, w, bias, norm_w, resid, out, K, N, pro, eps=0.0, pair=0):
                    _lib.count_launches(1)
                    if skinny:
                        _lib.check(lib.llm_skinny_linear(x.data_ptr(), x.stride(0), w.data_ptr(), _lib.ptr(bias), _lib.ptr(norm_w), float(eps),
                                                         _lib.ptr(resid), resid.stride(0) if resid is not None else 0, out.data_ptr(),
                                                         out.stride(0), B, N, K, pair, _lib.stream_ptr(dev)), "llm_skinny_linear")
                    else:
                        _lib.check(lib.llm_gemv(x.data_ptr(), w.data_ptr(), _lib.ptr(bias), _lib.ptr(norm_w), _lib.ptr(resid), out.data_ptr(),
                                                B, K, N, pro, float(eps), pair, _lib.stream_ptr(dev)), "llm_gemv")
                return run

            g_qkv, g_o, g_gu, g_down, g_head = (gemv_for(n) for n in ("qkv", "o", "gate_up", "down", "head"))

            def settle(x, pending, w, eps):
                """(stream, normed): adds a pending cuBLAS output into the stream inside the RMSNorm launch."""
                if pending is None:
                    return x, _rmsnorm(x, None, w, eps)[1]
                return _rmsnorm(pending, x, w, eps)

            def step_mixed():
                x, pending = inner.embed_tokens(c["tok"]).view(B, hid), None
                for i, ly in enumerate(self.layers):
                    if "qkv" in mixed and pending is None:
                        g_qkv(x, ly["wqkv"], ly["bqkv"], ly["w_in"], None, qkv_b, hid, n_qkv, 1, ly["eps_in"])
                        qkv = qkv_b
                    else:
                        x, normed = settle(x, pending, ly["w_in"], ly["eps_in"])
                        pending = None
                        if "qkv" in mixed:
                            g_qkv(normed, ly["wqkv"], ly["bqkv"], None, None, qkv_b, hid, n_qkv, 0)
                            qkv = qkv_b
                        else:
                            qkv = F.linear(normed, ly["wqkv"], ly["bqkv"])
                    _lib.count_launches(1)
                    _lib.check(lib.llm_rope_qkv(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), q.data_ptr(), kc[i].data_ptr(),
                                                vc[i].data_ptr(), B, 1, H, Hkv, D, kc[i].stride(0), kc[i].stride(1), 0,
                                                c["pos"].data_ptr(), 1, _lib.stream_ptr(dev)), "llm_rope_qkv")
                    sdag.sdag_decode_attention(q.transpose(1, 2), kc[i], vc[i], ly["scaling"], kv_len_dev=c["len"], out=attn_out,
                                               workspace=dec_ws)
                    a = attn_out.view(B, H * D)
                    if "o" in mixed:
                        g_o(a, ly["wo"], ly["bo"], None, x, res_a, H * D, hid, 0)
                        x = res_a
                    else:
                        pending = F.linear(a, ly["wo"], ly["bo"])
                    if "gate_up" in mixed and pending is None:
                        g_gu(x, ly["wgu"], None, ly["w_post"], None, gu_b, hid, 2 * inter, 1, ly["eps_post"], pair=1)
                    else:
                        x, normed = settle(x, pending, ly["w_post"], ly["eps_post"])
                        pending = None
                        if "gate_up" in mixed:
                            g_gu(normed, ly["wgu"], None, None, None, gu_b, hid, 2 * inter, 0, pair=1)
                        else:
                            gu = F.linear(normed, ly["wgu"])
                            _lib.count_launches(1)
                            _lib.check(lib.llm_swiglu(gu.data_ptr(), gu_b.data_ptr(), B, inter, _lib.stream_ptr(dev)), "llm_swiglu")
                    if "down" in mixed:
                        g_down(gu_b, ly["wdown"], None, None, x, res_b, inter, hid, 0)
                        x = res_b
                    else:
                        pending = F.linear(gu_b, ly["wdown"])
                if "head" in mixed and pending is None:
                    g_head(x, head_w, model.lm_head.bias, inner.norm.weight, None, logits_b, hid, head_w.shape[0], 1,
                         inner.norm.variance_epsilon)
                    lg = logits_b
                else:
                    x, normed = settle(x, pending, inner.norm.weight, inner.norm.variance_epsilon)
                    lg = model.lm_head(normed)
                if sample:
                    nxt = sample_tokens(lg, c["temp"], c["top_p"] if nucleus else None)
                else:
                    nxt = torch.argmax(lg, dim=-1, keepdim=True)
                c["gen"].scatter_(1, c["slot"], nxt)
                c["tok"].copy_(nxt)
                c["pos"].add_(1)
                c["len"].add_(1)
                c["slot"].add_(1)

            step = step_mixed
        c["step"] = step
        c["graph"] = None
        if use_graph:
            try:
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    step()  # warm-up outside capture: workspaces, cuBLAS handles
                torch.cuda.current_stream(dev).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    step()
                c["graph"] = graph
            except Exception as exc:  # eager replay of the same step: correct, but every launch pays host latency
                c["graph"] = None
                torch.cuda.synchronize(dev)
                _lib.note_fallback("FusedPrefill._decode_context", "CUDA-graph capture of the decode step failed", exc)
        ctxs[key] = c
        return c

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, blockmap, max_new_tokens: int, eos_token_id, temperature: float = 0.0,
                 use_graph: bool = True, poll_every: int = 16, lengths=None, top_p=None,
                 eos_from_first: bool = False) -> torch.Tensor:
        """New token ids ``[B, <= max_new_tokens + 1]`` (one from the prefill, the rest from the decode loop —
        SDAG.py:181-211; trimming at EOS is the caller's job).  The decode step (fused glue, split-KV decode
        attention over the preallocated cache, cuBLAS GEMMs, token pick) is captured once into a CUDA graph; position,
        cache length and output slot live in device memory and advance inside the graph, so the host only replays.

        Ragged batches (``lengths`` int64 ``[B]``, prompts right-padded to the common L): every prompt continues at its own
        position with its own cache length.  ``top_p`` adds nucleus filtering to temperature sampling (the CARG path,
        attack/content_generation.py:76-83)."""
        B, L = input_ids.shape
        dev = input_ids.device
        nucleus = temperature > 0 and top_p is not None and top_p < 1.0
        c = self._decode_context(B, L + max_new_tokens + 1, temperature, dev, use_graph, nucleus)
        self._last_ctx = c  # tools: the context this call used
        logits, _ = self.prefill(input_ids, blockmap, want_cache=False, kv=(c["kc"], c["vc"], c["cos"], c["sin"]),
                                 lengths=lengths)
        if temperature > 0:
            c["temp"].fill_(temperature)
            if nucleus:
                c["top_p"].fill_(float(top_p))
            tok = sample_tokens(logits, c["temp"], c["top_p"] if nucleus else None)
        else:
            tok = torch.argmax(logits, dim=-1, keepdim=True)
        c["gen"][:, 0:1] = tok
        c["tok"].copy_(tok)
        if lengths is None:
            c["pos"].fill_(L)
            c["len"].fill_(L + 1)
        else:
            c["pos"].copy_(lengths)
            c["len"].copy_(lengths + 1)
        c["slot"].fill_(1)
        eos_t = None  # one id or several (generation_config.eos_token_id may be a list)
        if eos_token_id is not None and not (isinstance(eos_token_id, (list, tuple)) and len(eos_token_id) == 0):
            eos_t = torch.as_tensor(eos_token_id, dtype=torch.int64, device=dev).reshape(-1)
        done = 0
        while done < max_new_tokens:
            n = min(poll_every, max_new_tokens - done)
            for _ in range(n):
                if c["graph"] is not None:
                    c["graph"].replay()
                else:
                    c["step"]()
            done += n
            first = 0 if eos_from_first else 1  # the SDAG loop never stops on the prefill token
            if eos_t is not None and bool(torch.isin(c["gen"][:, first : done + 1], eos_t).any(dim=1).all()):
                break
        return c["gen"][:, : done + 1].clone()


_gemm_events = None  # set to a list to collect (start, end, name, rows, N, K) around every prefill GEMM launch (_GemmTimer)
CUBLAS_PROJ = {p for p in os.environ.get("SDAG_CUBLAS_PROJ", "").split(",") if p}  # A/B measurements only: e.g. "o,down"
USE_PAIR_GEMM = os.environ.get("SDAG_PAIR_GEMM", "") != "off"  # "off": cuBLAS for every projection (A/B measurements only)
PAIR_GEMM_MIN_ROWS = 512  # decode steps of >= this many prompts would use llm_gemm for gate/up (never reached in practice)
PREFETCH_O_PROJ = os.environ.get("SDAG_L2_PREFETCH", "") == "1"  # decode step: o_proj weights prefetched into L2 during attention (measured 3 % slower: off)
FUSED_ROPE_GEMM = os.environ.get("SDAG_NO_FUSED_ROPE_GEMM", "") != "1"  # prefill q/k/v GEMM with the RoPE + cache-write epilogue
FUSED_ROPE_GEMV = os.environ.get("SDAG_NO_FUSED_ROPE", "") != "1"  # decode q/k/v GEMV with the RoPE + cache-write epilogue
# decode steps of 3-8 prompts: the linear layers to run on llm_gemv's warp-MMA form instead of cuBLAS ("qkv,o,gate_up,down,head")
DECODE_GEMV = {p.split("=")[0]: (p.split("=") + ["mma"])[1] for p in os.environ.get("SDAG_DECODE_GEMV", "").split(",") if p}
DECODE_CTX_BUDGET_BYTES = int(float(os.environ.get("SDAG_DECODE_CTX_GB", "24")) * 2 ** 30)  # KV caches of the cached decode contexts
FUSED_GEMV_MAX_B = min(4, max(1, int(os.environ.get("SDAG_FUSED_GEMV_MAX_B", "2"))))  # largest batch on the fused llm_gemv step
USE_FUSED_GEMV = True  # decode-step linear layers on llm_gemv (B <= 4); False: cuBLAS + stand-alone glue kernels


def _cache_adopt(cache, k: torch.Tensor, v: torch.Tensor, layer_idx: int) -> None:
    """Hand layer ``layer_idx`` of a fresh ``DynamicCache`` the prefill's K / V tensors themselves.  ``cache.update`` would
    ``torch.cat`` them onto an empty tensor — a full copy of the cache per prefill (HF:cache_utils.py DynamicLayer.update);
    the tensors are ours and nothing else aliases them, so the layer can simply own them.  Falls back to ``update`` when
    the cache class does not look like transformers' lazily initialised layers."""
    try:
        replicate = getattr(cache, "layer_class_to_replicate", None)
        if replicate is not None:
            while len(cache.layers) <= layer_idx:
                cache.layers.append(replicate())
        layer = cache.layers[layer_idx]
        if (not getattr(layer, "is_initialized", True) and hasattr(layer, "lazy_initialization") and hasattr(layer, "keys")
                and not getattr(cache, "offloading", False)):
            layer.lazy_initialization(k, v)
            layer.keys, layer.values = k, v
            return
    except Exception:
        pass
    cache.update(k, v, layer_idx)


def sample_tokens(logits: torch.Tensor, temperature: torch.Tensor, top_p: Optional[torch.Tensor]) -> torch.Tensor:
    """Temperature, then nucleus filtering, then one multinomial draw per row — the order of HF ``generate``'s logits
    warpers (temperature -> top-p).  A token survives when the probability mass ranked strictly above it is below
    ``top_p`` (the most likely token always survives).  Device-side scalars, no host sync: usable inside a CUDA graph."""
    probs = torch.softmax(logits.float() / temperature, dim=-1)
    if top_p is None:
        return torch.multinomial(probs, num_samples=1)
    sp, si = torch.sort(probs, dim=-1, descending=True)
    above = sp.cumsum(dim=-1) - sp
    sp = sp.masked_fill(above >= top_p, 0.0)
    return si.gather(1, torch.multinomial(sp, num_samples=1))


def _accepts_config(cls) -> bool:
    import inspect

    try:
        return "config" in inspect.signature(cls.__init__).parameters
    except (TypeError, ValueError):
        return False


def get_runner(model) -> Optional[FusedPrefill]:
    """Cached ``FusedPrefill`` of a model (None when the architecture / dtype is not covered: HF path then)."""
    r = getattr(model, "_sdag_b200_runner", False)
    if r is False:
        r = FusedPrefill.try_build(model)
        try:
            model._sdag_b200_runner = r
        except Exception:
            pass
    return r
