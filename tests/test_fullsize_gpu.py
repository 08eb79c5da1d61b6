"""GPU parity at BASELINE.json's full sizes, through size-independent checks the oracle can afford:
config 5 (L = 32 122, K = 2) attention on sampled rows, and the config-2 geometry against the same-GPU Hugging
Face path fed the reference's dense additive mask (what SDAG.py:161-176 does)."""
import numpy as np
import pytest
import torch

from oracle.sdag_mask import build_mask_oracle, synthetic_layout

pytestmark = pytest.mark.gpu


def _row_allowed(i, L, S, spans, qa, nbrs):
    """Allowed key indices of row i (same predicate as oracle.sdag_mask, one row at a time)."""
    cols = np.zeros(L, dtype=bool)
    if i < S or i >= qa:
        cols[: i + 1] = True
        return cols
    for d, (a, b) in enumerate(spans):
        if a <= i < b:
            cols[:S] = True
            cols[a : i + 1] = True
            if nbrs is not None:
                for n in nbrs[d]:
                    cols[spans[n][0] : spans[n][1]] = True
    return cols


def test_config5_long_context_attention_sampled_rows():
    import sdag_b200.sdag as sdag

    L, S, spans, qa = synthetic_layout(48, 50, 640, 1, 24)
    assert L == 32122
    rng = np.random.default_rng(42)
    emb = rng.standard_normal((50, 64)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    sims = emb @ emb.T
    nbrs = [[int(j) for j in np.argsort(-sims[r]) if j != r][:2] for r in range(50)]
    H, Hkv, D = 8, 2, 128
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(1, L, H, D, generator=g, device="cuda").to(torch.bfloat16).transpose(1, 2)
    k = torch.randn(1, Hkv, L, D, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(1, Hkv, L, D, generator=g, device="cuda").to(torch.bfloat16)
    bm = sdag.build_blockmap([sdag.SdagLayout(L, S, spans, qa, nbrs)])
    assert bm.header(0)["pairs"] == 53523828
    out = sdag.sdag_prefill_attention(q, k, v, bm, D ** -0.5)[0].float()  # [L, H, D]
    assert torch.isfinite(out).all()
    rows = [0, 47, 48, 300, 687, 688, 689, 5000, 20000, 32097, 32098, 32121] + list(rng.integers(0, L, 52))
    kf, vf = k[0].float(), v[0].float()
    vmean = vf.mean(dim=1)  # [Hkv, D]
    for i in rows:
        cols = torch.from_numpy(_row_allowed(int(i), L, S, spans, qa, nbrs)).cuda()
        for h in range(H):
            if not cols.any():
                ref = vmean[h // (H // Hkv)]
            else:
                s = (kf[h // (H // Hkv)][cols] @ q[0, h, i].float()) * D ** -0.5
                ref = torch.softmax(s, dim=0) @ vf[h // (H // Hkv)][cols]
            torch.testing.assert_close(out[i, h], ref, atol=2e-2, rtol=2e-2)


def test_config2_geometry_hook_matches_hf_with_reference_mask():
    """2 decoder layers of the Llama-3-8B geometry, bf16, L = 2642: logits of the hook path vs HF eager and HF sdpa
    given the dense additive finfo.min mask.  Also records what torch SDPA does with fully masked (gap) rows."""
    from transformers import LlamaConfig, LlamaForCausalLM

    import sdag_b200.sdag as sdag

    L, S, spans, qa = synthetic_layout(48, 10, 256, 1, 24)
    cfg = LlamaConfig(vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_hidden_layers=2,
                      num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192, rope_theta=500000.0)
    torch.manual_seed(42)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device("cuda"):
            model = LlamaForCausalLM(cfg).eval()
    finally:
        torch.set_default_dtype(prev)
    ids = torch.randint(0, 32000, (1, L), generator=torch.Generator().manual_seed(42)).cuda()
    mask = sdag.build_blocked_causal_mask_full(L, 500, S, spans, qa, "cuda", None)
    assert np.array_equal(mask.cpu().numpy(), build_mask_oracle(L, S, spans, qa))
    bias = torch.zeros(L, L, dtype=torch.bfloat16, device="cuda").masked_fill(~mask, torch.finfo(torch.bfloat16).min)[None, None]
    outs = {}
    with torch.no_grad():
        for impl in ("eager", "sdpa"):
            model.set_attn_implementation(impl)
            o = model(input_ids=ids, attention_mask=bias, use_cache=False, output_hidden_states=True)
            outs[impl] = (o.logits[0].float(), o.hidden_states[-1][0].float())
        sdag.prepare_model(model)
        bm = sdag.build_blockmap([sdag.SdagLayout(L, S, spans, qa)])
        o = model(input_ids=ids, use_cache=False, sdag_meta=bm, output_hidden_states=True)
        mine = (o.logits[0].float(), o.hidden_states[-1][0].float())
    print("sdpa finite on gap rows:", bool(torch.isfinite(outs["sdpa"][1]).all()))
    ref_logits, ref_hidden = outs["eager"]
    assert torch.isfinite(mine[1]).all()
    # bf16 end-to-end: compare relative to the spread of the reference's own two implementations
    spread = (outs["sdpa"][1] - ref_hidden).abs().max().item() if torch.isfinite(outs["sdpa"][1]).all() else 0.0
    err = (mine[1] - ref_hidden).abs().max().item()
    scale = ref_hidden.abs().max().item()
    print(f"hidden max-abs err vs eager {err:.4f} (eager-vs-sdpa spread {spread:.4f}, |hidden| max {scale:.2f})")
    # (on B200 torch's SDPA backend does NOT reproduce eager on the fully masked gap rows — spread ~6 — so eager,
    # the canonical semantics of SURVEY.md §8c, is the only yardstick for the hidden states)
    assert err <= 0.05 * scale
    # random-init logits are nearly flat, so arg-max flips under any bf16 perturbation: use the agreement between
    # the reference's own two implementations (eager vs sdpa) as the yardstick
    agree = (mine[0].argmax(-1) == ref_logits.argmax(-1)).float().mean().item()
    yard = (outs["sdpa"][0].argmax(-1) == ref_logits.argmax(-1)).float().mean().item()
    print(f"arg-max agreement over all {L} positions: mine-vs-eager {agree:.4f}, sdpa-vs-eager {yard:.4f}")
    assert agree >= yard - 0.03
    lerr = (mine[0] - ref_logits).abs().max().item()
    lspread = (outs["sdpa"][0] - ref_logits).abs().max().item()
    print(f"logits max-abs err vs eager {lerr:.4f} (eager-vs-sdpa {lspread:.4f})")
    assert lerr <= 0.3
