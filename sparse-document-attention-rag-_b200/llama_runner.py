"""Fused prefill forward for Llama-family decoders (path A4: the forward that ``generate_with_custom_mask`` drives,
SDAG.py:175-176).

Hugging Face's decoder layer runs ~25 small torch kernels around each attention call (RMSNorm in fp32 with six
passes, RoPE via slice / neg / cat, SiLU then mul, residual adds) — about a third of a Llama-3-8B prefill step on a
B200.  ``FusedPrefill`` keeps the HF module's weights (the q/k/v and gate/up projections are re-pointed at views of
fused weight matrices, no copy is kept) and drives: one cuBLAS GEMM for qkv, ``llm_rope_qkv`` (RoPE + KV-cache
write), the SDAG attention kernel, o_proj, ``llm_rmsnorm`` (fused residual add), one GEMM for gate+up,
``llm_swiglu``, down_proj.  Only the last position of each prompt goes through the final norm and the LM head
(the reference computes all L rows and uses one, SDAG.py:181).  The returned KV cache is a HF ``DynamicCache``,
so decoding continues through the normal HF path (with the ``sdag_b200`` decode attention).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn.functional as F

from . import _lib


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
    def try_build(model) -> Optional["FusedPrefill"]:
        try:
            cfg = model.config
            inner = model.model
            p0 = next(model.parameters())
            ok = (p0.dtype == torch.bfloat16 and p0.is_cuda and getattr(cfg, "hidden_act", "") == "silu"
                  and hasattr(inner, "rotary_emb") and hasattr(inner, "embed_tokens") and hasattr(model, "lm_head")
                  and cfg.hidden_size % 8 == 0 and cfg.intermediate_size % 8 == 0)
            D = getattr(cfg, "head_dim", None) or cfg.hidden_size // cfg.num_attention_heads
            ok = ok and D in (64, 128)
            for layer in inner.layers:
                at, mlp = layer.self_attn, layer.mlp
                for name in ("q_proj", "k_proj", "v_proj", "o_proj"):
                    ok = ok and isinstance(getattr(at, name, None), torch.nn.Linear)
                for name in ("gate_proj", "up_proj", "down_proj"):
                    ok = ok and isinstance(getattr(mlp, name, None), torch.nn.Linear) and getattr(mlp, name).bias is None
                ok = ok and hasattr(layer.input_layernorm, "variance_epsilon") and hasattr(layer.post_attention_layernorm, "variance_epsilon")
                ok = ok and not hasattr(at, "q_norm")  # per-head q/k norms (Qwen3, Gemma) are not handled
            return FusedPrefill(model) if ok else None
        except Exception:
            return None

    @torch.no_grad()
    def prefill(self, input_ids: torch.Tensor, blockmap, want_cache: bool = True):
        """input_ids int64 [B, L] on the model's device -> (last-position logits [B, vocab], DynamicCache | None)."""
        from . import sdag  # late import (sdag imports this module lazily too)

        lib = _lib.load()
        model, inner = self.model, self.model.model
        B, L = input_ids.shape
        T = B * L
        dev = input_ids.device
        H, Hkv, D = self.H, self.Hkv, self.D
        stream = _lib.stream_ptr(dev)
        h = inner.embed_tokens(input_ids)
        pos = torch.arange(L, device=dev).unsqueeze(0)
        cos, sin = inner.rotary_emb(h, pos)  # [1, L, D] in the model dtype
        cos, sin = cos[0].contiguous(), sin[0].contiguous()
        resid = h.view(T, self.hidden)
        _, normed = _rmsnorm(resid, None, self.layers[0]["w_in"], self.layers[0]["eps_in"])
        cache = None
        if want_cache:
            from transformers import DynamicCache

            cache = DynamicCache(config=model.config) if _accepts_config(DynamicCache) else DynamicCache()
        last_rows = torch.arange(B, device=dev) * L + (L - 1)
        n_layers = len(self.layers)
        for i, ly in enumerate(self.layers):
            qkv = F.linear(normed, ly["wqkv"], ly["bqkv"])
            q = torch.empty((B, L, H, D), dtype=torch.bfloat16, device=dev)
            k = torch.empty((B, Hkv, L, D), dtype=torch.bfloat16, device=dev)
            v = torch.empty((B, Hkv, L, D), dtype=torch.bfloat16, device=dev)
            _lib.count_launches(1)
            _lib.check(lib.llm_rope_qkv(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(),
                                        B, L, H, Hkv, D, k.stride(0), k.stride(1), 0, stream), "llm_rope_qkv")
            del qkv
            if cache is not None:
                cache.update(k, v, i)
            attn = sdag.sdag_prefill_attention(q.transpose(1, 2), k, v, blockmap, ly["scaling"])  # [B, L, H, D]
            attn = attn.view(T, H * D)
            last = i == n_layers - 1
            if last:  # only the last position of each prompt is read after the final layer
                attn, resid = attn[last_rows], resid[last_rows]
            o = F.linear(attn, ly["wo"], ly["bo"])
            resid, normed = _rmsnorm(o, resid, ly["w_post"], ly["eps_post"])
            gu = F.linear(normed, ly["wgu"])
            act = torch.empty((gu.shape[0], self.inter), dtype=torch.bfloat16, device=dev)
            _lib.count_launches(1)
            _lib.check(lib.llm_swiglu(gu.data_ptr(), act.data_ptr(), gu.shape[0], self.inter, stream), "llm_swiglu")
            del gu
            d = F.linear(act, ly["wdown"])
            if not last:
                nxt = self.layers[i + 1]
                resid, normed = _rmsnorm(d, resid, nxt["w_in"], nxt["eps_in"])
            else:
                _, normed = _rmsnorm(d, resid, inner.norm.weight, inner.norm.variance_epsilon)
        logits = model.lm_head(normed)
        return logits, cache


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
