"""GPU parity of the CTA-pair prefill attention kernel (sdag_prefill_pair_kernel: tcgen05 cta_group::2, two query heads
of a GQA group per CTA pair), forced with kernel="pair", against the float64 oracle / torch and against the single-CTA
kernel.  Same tolerance as the single-CTA kernel: atol = rtol = 2e-2 on N(0,1) bf16 inputs (P and the output are bf16)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.attention import masked_attention_oracle
from oracle.sdag_mask import build_mask_oracle, synthetic_layout

pytestmark = pytest.mark.gpu
TOL = dict(atol=2e-2, rtol=2e-2)


def _sdag():
    import sdag_b200.sdag as sdag

    return sdag


def _qkv(B, H, Hkv, L, D, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(B, H, L, D, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Hkv, L, D, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Hkv, L, D, generator=g, device="cuda").to(torch.bfloat16)
    return q, k, v


def _check(out, q, k, v, mask):
    D = q.shape[-1]
    for b in range(q.shape[0]):
        ref = masked_attention_oracle(q[b].float().cpu().numpy(), k[b].float().cpu().numpy(), v[b].float().cpu().numpy(), mask, D ** -0.5)
        got = out[b].float().cpu().numpy()
        assert np.isfinite(got).all()
        np.testing.assert_allclose(got, ref, **TOL)


@pytest.mark.parametrize("H,Hkv", [(4, 2), (8, 2), (2, 1)])  # G = 2 (one head per CTA), G = 4 (short tiles pack 2 heads per CTA), G = 2
def test_pair_kernel_matches_oracle_on_golden_layouts(golden_dir, H, Hkv):
    sdag = _sdag()
    with open(os.path.join(golden_dir, "masks.json")) as f:
        cases = json.load(f)
    for n, e in enumerate(cases):
        if e["L"] > 700:
            continue
        L = e["L"]
        q, k, v = _qkv(1, H, Hkv, L, 128, 42 + n)
        lay = sdag.SdagLayout(L, e["S"], [tuple(s) for s in e["spans"]], e["qa_start"], e["neighbors"])
        bm = sdag.build_blockmap([lay])
        if bm.header(0)["status"] != 0:
            continue
        mask = build_mask_oracle(L, e["S"], e["spans"], e["qa_start"], e["neighbors"])
        out = sdag.sdag_prefill_attention(q, k, v, bm, 128 ** -0.5, kernel="pair")
        torch.cuda.synchronize()
        _check(out, q, k, v, mask)


def test_pair_kernel_config2_batch_and_agreement_with_single_cta_kernel():
    """Config-2 layout (L = 2642), 8 q / 2 kv heads, two prompts with different layouts; then the same call through the
    single-CTA kernel: both are within tolerance of the oracle and of each other."""
    sdag = _sdag()
    L, S, spans, qa = synthetic_layout(48, 10, 256, 1, 24)
    L2, S2, spans2, qa2 = synthetic_layout(64, 10, 250, 2, L - (64 + 10 * 252))
    assert L2 == L
    q, k, v = _qkv(2, 8, 2, L, 128, 7)
    bm = sdag.build_blockmap([sdag.SdagLayout(L, S, spans, qa), sdag.SdagLayout(L, S2, spans2, qa2)])
    out = sdag.sdag_prefill_attention(q, k, v, bm, 128 ** -0.5, kernel="pair")
    single = sdag.sdag_prefill_attention(q, k, v, bm, 128 ** -0.5, kernel="single")
    for b, (s_, sp_, qa_) in enumerate(((S, spans, qa), (S2, spans2, qa2))):
        mask = build_mask_oracle(L, s_, sp_, qa_)
        _check(out[b : b + 1], q[b : b + 1], k[b : b + 1], v[b : b + 1], mask)
    torch.testing.assert_close(out.float(), single.float(), **TOL)


def test_pair_kernel_dense_mask_path_and_strided_q():
    sdag = _sdag()
    rng = np.random.default_rng(11)
    L, H, Hkv, D = 333, 4, 2, 128
    m = rng.random((L, L)) < 0.2
    m[5] = False
    m[200:210] = False
    g = torch.Generator(device="cuda").manual_seed(1)
    q = torch.randn(1, L, H, D, generator=g, device="cuda").to(torch.bfloat16).transpose(1, 2)
    k = torch.randn(1, Hkv, L, D, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(1, Hkv, L, D, generator=g, device="cuda").to(torch.bfloat16)
    bm = sdag.build_blockmap_from_mask(torch.from_numpy(m).cuda())
    out = sdag.sdag_prefill_attention(q, k, v, bm, D ** -0.5, kernel="pair")
    _check(out, q, k, v, m)


@pytest.mark.parametrize("L,B,H,Hkv", [(1000, 1, 8, 2), (4096, 2, 32, 8), (130, 3, 4, 2)])
def test_pair_kernel_causal_matches_torch_sdpa(L, B, H, Hkv):
    sdag = _sdag()
    q, k, v = _qkv(B, H, Hkv, L, 128, L)
    out = sdag.sdag_prefill_attention(q, k, v, sdag.causal_blockmap(L, q.device), 128 ** -0.5, kernel="pair")
    ref = torch.nn.functional.scaled_dot_product_attention(
        q.float(), k.float().repeat_interleave(H // Hkv, 1), v.float().repeat_interleave(H // Hkv, 1), is_causal=True)
    torch.testing.assert_close(out.float(), ref.transpose(1, 2), **TOL)


def test_pair_kernel_refuses_shapes_it_cannot_take():
    from sdag_b200._lib import SdagB200Error

    sdag = _sdag()
    for H, Hkv, D in ((8, 8, 128), (6, 2, 128), (4, 2, 64)):  # one head per kv head, odd group, head_dim 64
        q, k, v = _qkv(1, H, Hkv, 256, D, 3)
        with pytest.raises(SdagB200Error):
            sdag.sdag_prefill_attention(q, k, v, sdag.causal_blockmap(256, q.device), kernel="pair")
