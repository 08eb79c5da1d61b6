"""Oracle: masked attention exactly as the reference reaches it (TEST INFRASTRUCTURE).

The reference never computes attention itself: ``generate_with_custom_mask`` (SDAG.py:161-176) turns the
bool allow-mask into an additive ``[1,1,L,L]`` bias filled with ``torch.finfo(model_dtype).min`` and hands
it to the Hugging Face model, whose attention is
``softmax(Q K^T * D**-0.5 + bias, dtype=float32) @ V`` (HF ``models/llama/modeling_llama.py:199-221``,
"eager"; ``integrations/sdpa_attention.py:92-101`` computes the same thing).  Consequences restated here:

* a disallowed key gets weight exactly 0 whenever the row has at least one allowed key;
* a row with NO allowed key (gap rows) has every logit equal to ``finfo.min``, so
  its softmax is uniform over ALL L keys, future ones included: the output row is ``mean_j V[j]``;
* decode steps (SDAG.py:193-197) pass no mask: the new token attends to every cached position.

All arithmetic is float64 on the (already rounded) inputs, so the oracle is the exact value the bf16 CUDA
kernel approximates; the tolerance lives in the tests.
"""
from __future__ import annotations

import numpy as np


def masked_attention_oracle(q: np.ndarray, k: np.ndarray, v: np.ndarray, mask: np.ndarray, scaling: float) -> np.ndarray:
    """q [H, Lq, D], k/v [Hkv, S, D], mask bool [Lq, S] -> out [Lq, H, D] (float64).

    GQA: query head h reads kv head ``h // (H // Hkv)`` (HF ``repeat_kv``, modeling_llama.py:187-196).
    """
    q = np.asarray(q, dtype=np.float64)
    k = np.asarray(k, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    H, Lq, D = q.shape
    Hkv, S, _ = k.shape
    g = H // Hkv
    out = np.empty((Lq, H, D), dtype=np.float64)
    any_allowed = mask.any(axis=1)
    for h in range(H):
        kk = k[h // g]
        vv = v[h // g]
        s = (q[h] @ kk.T) * scaling
        s = np.where(mask, s, -np.inf)
        # rows with nothing allowed: all logits equal -> uniform over every key
        s[~any_allowed] = 0.0
        s -= s.max(axis=1, keepdims=True)
        p = np.exp(s)
        p /= p.sum(axis=1, keepdims=True)
        out[:, h, :] = p @ vv
    return out


def decode_attention_oracle(q: np.ndarray, k: np.ndarray, v: np.ndarray, scaling: float) -> np.ndarray:
    """One decode step, dense over the whole cache (SDAG.py:193-197): q [H, D], k/v [Hkv, S, D] -> [H, D]."""
    q = np.asarray(q, dtype=np.float64)
    H, D = q.shape
    S = k.shape[1]
    full = np.ones((1, S), dtype=bool)
    return masked_attention_oracle(q[:, None, :], k, v, full, scaling)[0]
