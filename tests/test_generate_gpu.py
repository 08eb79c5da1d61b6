"""GPU parity: decode attention, the Hugging Face hook, and the end-to-end SDAG generation functions vs the
reference outputs stored in tests/golden/generation.json (made by running the reference's SDAG.py on CPU)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle.attention import decode_attention_oracle
from oracle.sdag_mask import build_mask_oracle

import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,H,Hkv,D,B", [(1, 8, 2, 128, 1), (63, 8, 2, 128, 2), (64, 14, 2, 64, 1), (65, 4, 4, 128, 1),
                                         (1000, 32, 8, 128, 1), (5000, 8, 2, 128, 3), (33000, 32, 8, 128, 1)])
def test_decode_attention_matches_oracle(S, H, Hkv, D, B):
    import sdag_b200.sdag as sdag

    g = torch.Generator(device="cuda").manual_seed(S)
    q = torch.randn(B, H, 1, D, generator=g, device="cuda").to(torch.bfloat16)
    cap = S + 17  # cache allocated larger than the valid length: garbage beyond S must not be read
    kc = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    vc = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    kc[:, :, S:] = float("nan")
    vc[:, :, S:] = float("nan")
    out = sdag.sdag_decode_attention(q, kc[:, :, :S], vc[:, :, :S], D ** -0.5)
    assert out.shape == (B, 1, H, D)
    for b in range(B if S < 10000 else 1):
        ref = decode_attention_oracle(q[b, :, 0].float().cpu().numpy(), kc[b, :, :S].float().cpu().numpy(),
                                      vc[b, :, :S].float().cpu().numpy(), D ** -0.5)
        np.testing.assert_allclose(out[b, 0].float().cpu().numpy(), ref, atol=1e-2, rtol=1e-2)


@pytest.mark.parametrize("cap,lens,H,Hkv,D,B", [(4096, [1, 127, 128, 129, 2643, 4096], 32, 8, 128, 2),
                                                (2048, [5, 1000, 2047], 14, 2, 64, 1)])
def test_decode_attention_device_side_length(cap, lens, H, Hkv, D, B):
    """CUDA-graph form of the call: the grid is sized for the capacity, the valid length is read from device memory.
    Rows at or beyond the length hold NaN and must not leak into the result; every length must agree with the oracle."""
    import sdag_b200.sdag as sdag

    g = torch.Generator(device="cuda").manual_seed(cap)
    q = torch.randn(B, H, 1, D, generator=g, device="cuda").to(torch.bfloat16)
    k0 = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    v0 = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    n_dev = torch.zeros(1, dtype=torch.int32, device="cuda")
    for n in lens:
        kc, vc = k0.clone(), v0.clone()
        kc[:, :, n:] = float("nan")
        vc[:, :, n:] = float("nan")
        n_dev.fill_(n)
        out = sdag.sdag_decode_attention(q, kc, vc, D ** -0.5, kv_len_dev=n_dev)
        ref = decode_attention_oracle(q[0, :, 0].float().cpu().numpy(), kc[0, :, :n].float().cpu().numpy(),
                                      vc[0, :, :n].float().cpu().numpy(), D ** -0.5)
        np.testing.assert_allclose(out[0, 0].float().cpu().numpy(), ref, atol=1e-2, rtol=1e-2, err_msg=f"len {n}")


def test_decode_attention_per_prompt_lengths():
    """Ragged batch: every prompt has its own cache length (kv_len_dev_stride = 1), NaN beyond it."""
    import sdag_b200.sdag as sdag

    B, H, Hkv, D, cap = 4, 32, 8, 128, 3000
    lens = [1, 200, 1537, 3000]
    g = torch.Generator(device="cuda").manual_seed(9)
    q = torch.randn(B, H, 1, D, generator=g, device="cuda").to(torch.bfloat16)
    kc = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    vc = torch.randn(B, Hkv, cap, D, generator=g, device="cuda").to(torch.bfloat16)
    for b, n in enumerate(lens):
        kc[b, :, n:] = float("nan")
        vc[b, :, n:] = float("nan")
    out = sdag.sdag_decode_attention(q, kc, vc, D ** -0.5, kv_len_dev=torch.tensor(lens, dtype=torch.int32, device="cuda"))
    for b, n in enumerate(lens):
        ref = decode_attention_oracle(q[b, :, 0].float().cpu().numpy(), kc[b, :, :n].float().cpu().numpy(),
                                      vc[b, :, :n].float().cpu().numpy(), D ** -0.5)
        np.testing.assert_allclose(out[b, 0].float().cpu().numpy(), ref, atol=1e-2, rtol=1e-2, err_msg=f"prompt {b}")


# Greedy parity rule: every prompt's tokens equal the reference's; a deviation is accepted only at a position where the
# reference's OWN fp32 decision was a near-tie — the gap between its two best logits (stored with the goldens as
# ``top2_gaps``) is at most GAP_EXCUSE.  bf16 attention moves these logits by up to ~0.1 (the prefill-logit test below
# bounds the top-8 values by atol 0.15), so a gap beyond GAP_EXCUSE cannot be explained by rounding.  Tokens after an
# excused flip follow a different continuation and are not compared.
GAP_EXCUSE = 0.1


def tokens_match_or_excused(mine, ref, gaps, what):
    """True if equal; False if the first deviation is an excused near-tie; fails the test otherwise."""
    for t, (a, b) in enumerate(zip(mine, ref)):
        if a != b:
            assert t < len(gaps) and gaps[t] <= GAP_EXCUSE, \
                f"{what}: token {t} is {a}, reference {b}, and the reference's fp32 top-2 logit gap there is {gaps[t] if t < len(gaps) else None}"
            print(f"[excused flip] {what}: token {t} {a} vs {b}, reference top-2 gap {gaps[t]:.4f}")
            return False
    assert len(mine) == len(ref), (what, mine, ref)
    return True


def _records(golden_dir):
    with open(os.path.join(golden_dir, "generation.json")) as f:
        return json.load(f)


def test_prompt_builder_matches_reference_golden(golden_dir):
    import sdag_b200.sdag as sdag
    from sdag_b200.prompts import SYSTEM_PROMPT_RAG, USER_RAG_PROMPT

    tok = synth.make_tokenizer()
    for r in _records(golden_dir):
        chat, S, spans, qa, ranked = sdag.build_rag_chat_and_spans(tok, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT, synth.SAMPLE_QUERY,
                                                                   r["malicious"], r["retrieved"], r["rank"])
        assert hashlib.sha256(chat.encode()).hexdigest() == r["chat_sha256"]
        assert (S, [list(s) for s in spans], qa, ranked) == (r["S"], r["spans"], r["qa_start"], r["ranked_docs"])


def test_hook_prefill_logits_match_reference_golden(golden_dir):
    """Same random-init model as the golden run; fp32 weights, bf16 attention kernel.  The reference's prefill
    logits (fp32 CPU) must be reproduced: same arg-max, top-8 values within bf16-attention tolerance."""
    import sdag_b200.sdag as sdag

    tok = synth.make_tokenizer()
    model = sdag.prepare_model(synth.tiny_llama(len(tok)).cuda())
    for r in _records(golden_dir):
        from sdag_b200.prompts import SYSTEM_PROMPT_RAG, USER_RAG_PROMPT
        chat, S, spans, qa, ranked = sdag.build_rag_chat_and_spans(tok, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT, synth.SAMPLE_QUERY,
                                                                   r["malicious"], r["retrieved"], r["rank"])
        ids = tok(chat, return_tensors="pt").input_ids.cuda()
        bm = sdag.build_blockmap([sdag.SdagLayout(ids.shape[1], S, spans, qa, r["neighbors"])])
        with torch.no_grad():
            logits = model(input_ids=ids, use_cache=True, sdag_meta=bm, logits_to_keep=1).logits[0, -1].float().cpu()
        top_ids = r["prefill_top_ids"]
        assert int(logits.argmax()) == top_ids[0], r["case"]
        np.testing.assert_allclose(logits[top_ids].numpy(), np.array(r["prefill_top_logits"]), atol=0.15, rtol=0.05)


def test_run_rag_with_doc_isolation_greedy_matches_reference_golden(golden_dir):
    import sdag_b200.sdag as sdag

    sdag.TEMPERATURE = 0
    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok)).cuda()
    ranker = synth.HashRanker()
    agree = 0
    recs = _records(golden_dir)
    for r in recs:
        seen = {}
        orig = tok.decode

        def spy(ids, **kw):
            seen["ids"] = [int(x) for x in ids]
            return orig(ids, **kw)

        tok.decode = spy
        ans = sdag.run_rag_with_doc_isolation(model=model, tokenizer=tok, query=synth.SAMPLE_QUERY, malicious_docs=r["malicious"],
                                              retrieved_docs=r["retrieved"], max_new_tokens=6, device="cuda",
                                              add_attack_in_rank=r["rank"], ranker=ranker, doc_neighbors_k=r["doc_neighbors_k"])
        tok.decode = orig
        assert len(seen["ids"]) == 7  # one token from the prefill + max_new_tokens (Appendix A7)
        if tokens_match_or_excused(seen["ids"], r["new_token_ids"], r["top2_gaps"], r["case"]):
            assert ans == r["answer"]
            agree += 1
    assert agree >= 2, agree  # the goldens hold two prompts whose every decision has a gap > GAP_EXCUSE


def test_generate_with_custom_mask_dense_path_equals_span_path(golden_dir):
    import sdag_b200.sdag as sdag

    sdag.TEMPERATURE = 0
    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok)).cuda()
    r = _records(golden_dir)[1]
    from sdag_b200.prompts import SYSTEM_PROMPT_RAG, USER_RAG_PROMPT
    chat, S, spans, qa, _ = sdag.build_rag_chat_and_spans(tok, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT, synth.SAMPLE_QUERY,
                                                          r["malicious"], r["retrieved"], r["rank"])
    ids = tok(chat, return_tensors="pt").input_ids.cuda()
    L = ids.shape[1]
    mask = sdag.build_blocked_causal_mask_full(L, 500, S, spans, qa, "cuda", r["neighbors"])
    assert np.array_equal(mask.cpu().numpy(), build_mask_oracle(L, S, spans, qa, r["neighbors"]))
    a_bool = sdag.generate_with_custom_mask(model, tok, ids, mask, max_new_tokens=6)
    additive = torch.zeros(L, L, device="cuda").masked_fill(~mask, torch.finfo(torch.float32).min)
    a_float = sdag.generate_with_custom_mask(model, tok, ids, additive, max_new_tokens=6)
    a_span = sdag.run_rag_with_doc_isolation(model=model, tokenizer=tok, query=synth.SAMPLE_QUERY, malicious_docs=r["malicious"],
                                             retrieved_docs=r["retrieved"], max_new_tokens=6, device="cuda",
                                             add_attack_in_rank=r["rank"], ranker=synth.HashRanker(),
                                             doc_neighbors_k=r["doc_neighbors_k"])
    assert a_bool == a_float == a_span
    with pytest.raises(NotImplementedError):
        sdag.generate_with_custom_mask(model, tok, ids, torch.full((L, L), -1.0, device="cuda"), max_new_tokens=1)


def test_eos_stops_generation_inside_loop_only():
    """EOS as the first (prefill) token does not stop; an EOS produced in the loop does, and is stripped."""
    import sdag_b200.sdag as sdag

    sdag.TEMPERATURE = 0
    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok)).cuda()
    ids = tok("hello world, this is a prompt", return_tensors="pt").input_ids.cuda()
    L = ids.shape[1]
    mask = torch.ones(L, L, dtype=torch.bool, device="cuda").tril()
    base = sdag.generate_with_custom_mask(model, tok, ids, mask, max_new_tokens=40)
    seen = {}
    orig = tok.decode
    tok.decode = lambda i, **kw: (seen.__setitem__("ids", [int(x) for x in i]), orig(i, **kw))[1]
    sdag.generate_with_custom_mask(model, tok, ids, mask, max_new_tokens=40)
    full = list(seen["ids"])
    # pretend the 20th generated token is EOS
    tok2 = synth.make_tokenizer()
    tok2.decode = lambda i, **kw: (seen.__setitem__("ids2", [int(x) for x in i]), orig(i, **kw))[1]
    target = full[19]
    first = full.index(target)
    tok2.eos_token_id  # property exists
    class Tok:  # thin proxy overriding eos_token_id
        eos_token_id = target
        def decode(self, i, **kw):
            seen["ids2"] = [int(x) for x in i]
            return orig(i, **kw)
    sdag.generate_with_custom_mask(model, Tok(), ids, mask, max_new_tokens=40)
    cut = first if first >= 1 else full.index(target, 1)
    assert seen["ids2"] == full[: cut + 1]
    tok.decode = orig


def test_glue_kernels_match_torch_module_math():
    """llm_rmsnorm / llm_rope_qkv / llm_swiglu against the HF module formulas evaluated with torch in bf16."""
    import sdag_b200._lib as L
    from sdag_b200.llama_runner import _rmsnorm

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(0)
    rows, dim = 37, 4096
    x = torch.randn(rows, dim, generator=g, device="cuda").to(torch.bfloat16)
    res = torch.randn(rows, dim, generator=g, device="cuda").to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(dim, generator=g, device="cuda")).to(torch.bfloat16)

    def hf_norm(h, eps=1e-5):
        f = h.float()
        f = f * torch.rsqrt(f.pow(2).mean(-1, keepdim=True) + eps)
        return w * f.to(torch.bfloat16)

    _, y = _rmsnorm(x, None, w, 1e-5)
    torch.testing.assert_close(y.float(), hf_norm(x).float(), atol=2e-2, rtol=2e-2)
    s, y2 = _rmsnorm(x, res, w, 1e-5)
    assert torch.equal(s, x + res)
    torch.testing.assert_close(y2.float(), hf_norm(x + res).float(), atol=2e-2, rtol=2e-2)
    assert (y2 == hf_norm(x + res)).float().mean() > 0.99  # same rounding order: almost every element bit-equal

    inter = 1024
    gu = torch.randn(rows, 2 * inter, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.empty(rows, inter, dtype=torch.bfloat16, device="cuda")
    L.check(lib.llm_swiglu(gu.data_ptr(), out.data_ptr(), rows, inter, L.stream_ptr(out.device)), "swiglu")
    ref = torch.nn.functional.silu(gu[:, :inter]) * gu[:, inter:]
    torch.testing.assert_close(out.float(), ref.float(), atol=2e-2, rtol=2e-2)
    assert (out == ref).float().mean() > 0.99

    B, Lq, H, Hkv, D = 2, 19, 4, 2, 128
    qkv = torch.randn(B * Lq, (H + 2 * Hkv) * D, generator=g, device="cuda").to(torch.bfloat16)
    ang = torch.rand(Lq + 3, D // 2, generator=g, device="cuda") * 6.0  # tables are indexed by absolute position
    emb = torch.cat([ang, ang], dim=-1)
    cos_t, sin_t = emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)
    cos, sin = cos_t[3:], sin_t[3:]
    q = torch.empty(B, Lq, H, D, dtype=torch.bfloat16, device="cuda")
    k = torch.zeros(B, Hkv, Lq + 5, D, dtype=torch.bfloat16, device="cuda")
    v = torch.zeros_like(k)
    L.check(lib.llm_rope_qkv(qkv.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(), B, Lq, H,
                             Hkv, D, k.stride(0), k.stride(1), 3, 0, 0, L.stream_ptr(q.device)), "rope")

    def rot(t):
        return torch.cat((-t[..., D // 2:], t[..., : D // 2]), dim=-1)

    x3 = qkv.view(B, Lq, H + 2 * Hkv, D)
    qr, kr, vr = x3[:, :, :H], x3[:, :, H : H + Hkv], x3[:, :, H + Hkv :]
    c, s_ = cos[None, :, None, :], sin[None, :, None, :]
    assert torch.equal(q, qr * c + rot(qr) * s_)
    assert torch.equal(k[:, :, 3 : 3 + Lq], (kr * c + rot(kr) * s_).transpose(1, 2))
    assert torch.equal(v[:, :, 3 : 3 + Lq], vr.transpose(1, 2))
    assert (k[:, :, :3] == 0).all() and (k[:, :, 3 + Lq :] == 0).all()
    # ragged decode step: one token per prompt, each at its own device-side position (pos0_dev_stride = 1)
    pos = torch.tensor([2, 7], dtype=torch.int32, device="cuda")
    q1 = torch.empty(B, 1, H, D, dtype=torch.bfloat16, device="cuda")
    k1 = torch.zeros(B, Hkv, Lq + 5, D, dtype=torch.bfloat16, device="cuda")
    v1 = torch.zeros_like(k1)
    one = qkv.view(B, Lq, -1)[:, 0].contiguous()
    L.check(lib.llm_rope_qkv(one.data_ptr(), cos_t.data_ptr(), sin_t.data_ptr(), q1.data_ptr(), k1.data_ptr(), v1.data_ptr(), B, 1, H,
                             Hkv, D, k1.stride(0), k1.stride(1), 0, pos.data_ptr(), 1, L.stream_ptr(q.device)), "rope ragged")
    for b, pb in enumerate(pos.tolist()):
        cb, sb = cos_t[pb][None, :], sin_t[pb][None, :]
        assert torch.equal(q1[b, 0], qr[b, 0] * cb + rot(qr[b, 0]) * sb)
        assert torch.equal(k1[b, :, pb], kr[b, 0] * cb + rot(kr[b, 0]) * sb)
        assert torch.equal(v1[b, :, pb], vr[b, 0])
        assert int((k1[b] != 0).any(dim=-1).sum()) == Hkv  # exactly one row per kv head written


def test_fused_prefill_runner_matches_hf_hook_path(golden_dir):
    """bf16 model: the fused runner (one GEMM for qkv / gate+up, fused glue) vs the HF forward with the same
    attention kernel hooked in: same arg-max, logits within bf16 GEMM-order noise; the returned cache decodes."""
    import sdag_b200.sdag as sdag
    from sdag_b200.llama_runner import get_runner
    from sdag_b200.prompts import SYSTEM_PROMPT_RAG, USER_RAG_PROMPT

    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok), dtype=torch.bfloat16, hidden=256, heads=4, kv_heads=2, head_dim=64, inter=512,
                             init_scale=3.0).cuda()
    sdag.prepare_model(model)
    assert get_runner(model) is not None
    r = _records(golden_dir)[1]
    chat, S, spans, qa, _ = sdag.build_rag_chat_and_spans(tok, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT, synth.SAMPLE_QUERY,
                                                          r["malicious"], r["retrieved"], r["rank"])
    ids = tok(chat, return_tensors="pt").input_ids.cuda()
    ids = torch.cat([ids, ids.flip(1)], dim=0)  # batch of 2 different prompts, same layout
    bm = sdag.build_blockmap([sdag.SdagLayout(ids.shape[1], S, spans, qa, r["neighbors"])])
    sdag.USE_FUSED_PREFILL = False
    try:
        ref_logits, ref_cache = sdag.prefill_forward(model, ids, bm)
    finally:
        sdag.USE_FUSED_PREFILL = True
    logits, cache = sdag.prefill_forward(model, ids, bm)
    assert torch.equal(logits.argmax(-1), ref_logits.argmax(-1))
    torch.testing.assert_close(logits.float(), ref_logits.float(), atol=0.1, rtol=0.05)
    # one decode step on each cache gives the same next token
    nxt = logits.argmax(-1, keepdim=True)
    with torch.no_grad():
        a = model(input_ids=nxt, past_key_values=cache, use_cache=True).logits[:, -1]
        b = model(input_ids=nxt, past_key_values=ref_cache, use_cache=True).logits[:, -1]
    assert torch.equal(a.argmax(-1), b.argmax(-1))
    torch.testing.assert_close(a.float(), b.float(), atol=0.1, rtol=0.05)


def test_fused_generate_graph_decode_matches_hf_loop(golden_dir):
    """bf16 Llama-family model: prefill into the preallocated cache + CUDA-graph decode loop must give the same
    greedy tokens as the per-token Hugging Face loop with the same attention kernels, incl. EOS trimming."""
    import sdag_b200.sdag as sdag
    from sdag_b200.llama_runner import get_runner

    sdag.TEMPERATURE = 0
    tok = synth.make_tokenizer()
    model = synth.tiny_llama(len(tok), dtype=torch.bfloat16, hidden=256, heads=4, kv_heads=2, head_dim=64, inter=512,
                             init_scale=3.0).cuda()
    r = _records(golden_dir)[0]
    kwargs = dict(model=model, tokenizer=tok, query=synth.SAMPLE_QUERY, malicious_docs=r["malicious"], retrieved_docs=r["retrieved"],
                  max_new_tokens=40, device="cuda", add_attack_in_rank=r["rank"], ranker=synth.HashRanker(), doc_neighbors_k=0)
    seen = {}
    orig = tok.decode
    tok.decode = lambda i, **kw: (seen.__setitem__("ids", [int(x) for x in i]), orig(i, **kw))[1]
    try:
        sdag.USE_FUSED_PREFILL = False
        a = sdag.run_rag_with_doc_isolation(**kwargs)
        ids_hf = list(seen["ids"])
        sdag.USE_FUSED_PREFILL = True
        assert get_runner(model) is not None
        b = sdag.run_rag_with_doc_isolation(**kwargs)
        ids_fused = list(seen["ids"])
    finally:
        sdag.USE_FUSED_PREFILL = True
        tok.decode = orig
    assert len(ids_hf) == 41
    # bf16 GEMM-order noise may flip a late near-tie on a random model; the bulk of the sequence must be identical
    same = sum(1 for x, y in zip(ids_hf, ids_fused) if x == y)
    first_diff = next((i for i, (x, y) in enumerate(zip(ids_hf, ids_fused)) if x != y), len(ids_hf))
    assert first_diff >= 8, (first_diff, ids_hf, ids_fused)
    if first_diff == len(ids_hf):
        assert a == b
    # graph replay == eager stepping of the same engine (bit-identical kernels and inputs)
    runner = get_runner(model)
    from sdag_b200.prompts import SYSTEM_PROMPT_RAG, USER_RAG_PROMPT
    chat, S, spans, qa, _ = sdag.build_rag_chat_and_spans(tok, SYSTEM_PROMPT_RAG, USER_RAG_PROMPT, synth.SAMPLE_QUERY,
                                                          r["malicious"], r["retrieved"], r["rank"])
    ids = tok(chat, return_tensors="pt").input_ids.cuda()
    bm = sdag.build_blockmap([sdag.SdagLayout(ids.shape[1], S, spans, qa)])
    g1 = runner.generate(ids, bm, 24, None, use_graph=True)
    g2 = runner.generate(ids, bm, 24, None, use_graph=False)
    assert torch.equal(g1, g2)
    assert g1.shape == (1, 25)


# ------------------------------------------------------------------------------------------------------
# CARG: generate_batch_seq2seq / generate_noiso_batch (attack/content_generation.py:44-94, main.py:308-378)
# ------------------------------------------------------------------------------------------------------
def _carg_gold(golden_dir):
    with open(os.path.join(golden_dir, "carg.json")) as f:
        return json.load(f)


def _spy_decode(tok, sink):
    orig = tok.decode
    tok.decode = lambda i, **kw: (sink.append([int(x) for x in i]), orig(i, **kw))[1]
    return orig


def test_carg_batch_generation_matches_reference_golden(golden_dir):
    """Same random-init fp32 model as the golden run (reference code: HF ``generate`` on a left-padded batch, CPU).
    Here every prompt runs unpadded through the attention hook; greedy tokens must be the reference's (bf16 attention
    on a random-init model may flip a near-tie: at most one prompt per call may deviate)."""
    import sdag_b200.config as cfg
    from sdag_b200.generation import generate_batch_seq2seq, generate_noiso_batch

    gold = _carg_gold(golden_dir)
    tok = synth.make_tokenizer()
    tok.padding_side = "left"
    model = synth.tiny_llama(len(tok)).cuda()
    seen = []
    orig = _spy_decode(tok, seen)
    try:
        out = generate_batch_seq2seq(tok, model, gold["raw_prompts"], max_tokens=gold["raw_max_tokens"], temperature=0, top_p=1.0)
        raw_ids, seen[:] = list(seen), []
        saved = (cfg.ORACLE, cfg.LLM_BATCH_SIZE, cfg.MAX_GEN_TOKENS_RAG, cfg.TEMPERATURE, cfg.TOP_P)
        cfg.ORACLE, cfg.LLM_BATCH_SIZE, cfg.MAX_GEN_TOKENS_RAG = True, gold["llm_batch_size"], gold["max_gen_tokens"]
        cfg.TEMPERATURE, cfg.TOP_P = 0, 1.0
        try:
            answers = generate_noiso_batch(tok, model, gold["queries"], gold["defended"], gold["malicious"],
                                           attacker_pos=gold["attacker_pos"], order_mode=gold["order_mode"])
        finally:
            cfg.ORACLE, cfg.LLM_BATCH_SIZE, cfg.MAX_GEN_TOKENS_RAG, cfg.TEMPERATURE, cfg.TOP_P = saved
        noiso_ids = list(seen)
    finally:
        tok.decode = orig
    assert len(raw_ids) == len(gold["raw_token_ids"]) and len(noiso_ids) == len(gold["token_ids"])
    same_raw = [tokens_match_or_excused(a, b, g, f"raw prompt {n}")
                for n, (a, b, g) in enumerate(zip(raw_ids, gold["raw_token_ids"], gold["raw_top2_gaps"]))]
    for ok, mine, ref in zip(same_raw, out, gold["raw_answers"]):
        if ok:
            assert mine == ref
    same_noiso = [tokens_match_or_excused(a, b, g, f"noiso prompt {n}")
                  for n, (a, b, g) in enumerate(zip(noiso_ids, gold["token_ids"], gold["top2_gaps"]))]
    for ok, mine, ref in zip(same_noiso, answers, gold["answers"]):
        if ok:
            assert mine == ref
    assert sum(same_raw) >= 2 and sum(same_noiso) >= 1


def _bf16_model(tok):
    return synth.tiny_llama(len(tok), dtype=torch.bfloat16, hidden=256, heads=4, kv_heads=2, head_dim=64, inter=512,
                            init_scale=3.0).cuda()


def test_carg_ragged_fused_batch_equals_hf_generate_and_single_runs():
    """bf16 Llama-family model on the fused engine: a ragged batch (per-prompt block maps, RoPE positions and cache
    lengths, one CUDA-graph decode step for all) against (a) HF ``model.generate`` on the left-padded batch with the
    stock sdpa attention — what the reference's generate_batch_seq2seq executes — and (b) each prompt generated alone."""
    import sdag_b200.sdag as sdag
    from sdag_b200.generation import generate_batch_seq2seq
    from sdag_b200.llama_runner import get_runner

    tok = synth.make_tokenizer()
    tok.padding_side = "left"
    model = _bf16_model(tok)
    prompts = ["Paris is the capital of", "Honey", "The highest mountain above sea level is called Mount Everest and it",
               "Octopuses have three hearts", "a"]
    n_new = 24
    # (a) the reference's call, stock attention
    enc = tok(prompts, return_tensors="pt", padding=True, truncation=True).to("cuda")
    model.set_attn_implementation("sdpa")
    with torch.no_grad():
        ref = model.generate(input_ids=enc["input_ids"], attention_mask=enc["attention_mask"], max_new_tokens=n_new,
                             do_sample=False, pad_token_id=tok.eos_token_id)
    ref_new = [row[enc["input_ids"].shape[1]:].tolist() for row in ref]
    seen = []
    orig = _spy_decode(tok, seen)
    try:
        assert get_runner(sdag.prepare_model(model)) is not None
        generate_batch_seq2seq(tok, model, prompts, max_tokens=n_new, temperature=0, top_p=1.0)
        batch_ids, seen[:] = list(seen), []
        for p in prompts:
            generate_batch_seq2seq(tok, model, [p], max_tokens=n_new, temperature=0, top_p=1.0)
        single_ids = list(seen)
    finally:
        tok.decode = orig
    from sdag_b200.generation import _eos_ids

    stops = set(_eos_ids(tok, model))  # what HF generate stops on: generation_config.eos_token_id, not the tokenizer's id

    def cut(ids):
        return next((ids[: i + 1] for i, t in enumerate(ids) if t in stops), ids)

    def first_diff(a, b):
        return next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))

    for b, (mine, alone, hf) in enumerate(zip(batch_ids, single_ids, ref_new)):
        hf = cut(hf)
        assert len(mine) == len(hf) or first_diff(mine, hf) < min(len(mine), len(hf)), (b, mine, hf)
        # bf16 GEMM-order noise (different batch shapes) may flip a late near-tie on a random model
        assert first_diff(mine, alone) >= min(8, len(alone)), (b, mine, alone)
        assert first_diff(mine, hf) >= min(6, len(hf)), (b, mine, hf)
    agree = sum(1 for mine, hf in zip(batch_ids, ref_new) if mine == cut(hf))
    assert agree >= len(prompts) - 2, (batch_ids, ref_new)


def test_carg_sampling_top_p_limits_and_runs_in_graph():
    """Nucleus sampling inside the captured decode step: a vanishing top_p keeps only the most likely token (= greedy);
    ordinary settings produce tokens and respect max_tokens."""
    from sdag_b200.generation import generate_batch_seq2seq

    tok = synth.make_tokenizer()
    tok.padding_side = "left"
    model = _bf16_model(tok)
    prompts = ["Paris is the capital of", "Honey never"]
    seen = []
    orig = _spy_decode(tok, seen)
    try:
        generate_batch_seq2seq(tok, model, prompts, max_tokens=12, temperature=0, top_p=1.0)
        greedy, seen[:] = list(seen), []
        generate_batch_seq2seq(tok, model, prompts, max_tokens=12, temperature=1.0, top_p=1e-6)
        tiny_p, seen[:] = list(seen), []
        torch.manual_seed(0)
        generate_batch_seq2seq(tok, model, prompts, max_tokens=12, temperature=0.7, top_p=0.9)
        sampled = list(seen)
    finally:
        tok.decode = orig
    assert tiny_p == greedy
    assert all(1 <= len(s) <= 12 for s in sampled)


def test_sdag_ragged_batch_equals_per_query_calls(golden_dir):
    """run_rag_with_doc_isolation_batch (prompts of different lengths, gap rows, neighbour lists) against the per-query
    function on the fused engine: same greedy tokens up to late bf16 near-ties."""
    import sdag_b200.sdag as sdag

    sdag.TEMPERATURE = 0
    tok = synth.make_tokenizer()
    model = _bf16_model(tok)
    recs = _records(golden_dir)
    queries = [synth.SAMPLE_QUERY, "Which ocean is the largest?", "How many hearts?", synth.SAMPLE_QUERY + " Please answer."]
    mals = [recs[0]["malicious"], [], recs[1]["malicious"], recs[0]["malicious"]]
    rets = [recs[0]["retrieved"], recs[3]["retrieved"], recs[1]["retrieved"], recs[0]["retrieved"][:2]]
    for nbr_k in (0, 2):
        seen = []
        orig = _spy_decode(tok, seen)
        try:
            batch = sdag.run_rag_with_doc_isolation_batch(model, tok, queries, mals, rets, max_new_tokens=20, device="cuda",
                                                          add_attack_in_rank=1, ranker=synth.HashRanker(), doc_neighbors_k=nbr_k)
            batch_ids, seen[:] = list(seen), []
            single = [sdag.run_rag_with_doc_isolation(model, tok, q, m, r, max_new_tokens=20, device="cuda", add_attack_in_rank=1,
                                                      ranker=synth.HashRanker(), doc_neighbors_k=nbr_k)
                      for q, m, r in zip(queries, mals, rets)]
            single_ids = list(seen)
        finally:
            tok.decode = orig
        assert len(batch) == len(queries)
        exact = 0
        for a, b_, ta, tb in zip(batch_ids, single_ids, batch, single):
            fd = next((i for i, (x, y) in enumerate(zip(a, b_)) if x != y), min(len(a), len(b_)))
            assert fd >= min(8, len(b_)), (nbr_k, a, b_)
            if a == b_:
                assert ta == tb
                exact += 1
        assert exact >= len(queries) - 1, (nbr_k, batch_ids, single_ids)


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 8])
def test_decode_gemv_matches_module_math(B):
    """llm_gemv: plain, RMSNorm-prologue, SwiGLU-prologue, bias and residual epilogues vs the torch bf16 module order."""
    import sdag_b200._lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(B)
    K, N, I = 1024, 528, 2048  # N % 16 == 0: the tensor-core form (B > 2) works on 16-row blocks
    x = torch.randn(B, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16)
    nw = (1 + 0.1 * torch.randn(K, generator=g, device="cuda")).to(torch.bfloat16)
    res = torch.randn(B, N, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.empty(B, N, dtype=torch.bfloat16, device="cuda")
    st = L.stream_ptr(x.device)

    def close(a, ref):
        torch.testing.assert_close(a.float(), ref.float(), atol=3e-2, rtol=3e-2)
        assert (a == ref).float().mean() > 0.9  # same roundings; fp32 summation order differs from cuBLAS

    L.check(lib.llm_gemv(x.data_ptr(), W.data_ptr(), bias.data_ptr(), 0, 0, out.data_ptr(), B, K, N, 0, 0.0, 0, st), "plain")
    close(out, torch.nn.functional.linear(x, W, bias))
    L.check(lib.llm_gemv(x.data_ptr(), W.data_ptr(), 0, 0, res.data_ptr(), out.data_ptr(), B, K, N, 0, 0.0, 0, st), "residual")
    close(out, res + torch.nn.functional.linear(x, W))
    xf = x.float()
    normed = nw * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16)
    L.check(lib.llm_gemv(x.data_ptr(), W.data_ptr(), 0, nw.data_ptr(), 0, out.data_ptr(), B, K, N, 1, 1e-5, 0, st), "rmsnorm")
    close(out, torch.nn.functional.linear(normed, W))
    gu = torch.randn(B, 2 * I, generator=g, device="cuda").to(torch.bfloat16)
    Wd = (torch.randn(N, I, generator=g, device="cuda") * I ** -0.5).to(torch.bfloat16)
    act = torch.nn.functional.silu(gu[:, :I]) * gu[:, I:]
    L.check(lib.llm_gemv(gu.data_ptr(), Wd.data_ptr(), 0, 0, res.data_ptr(), out.data_ptr(), B, I, N, 2, 0.0, 0, st), "swiglu")
    close(out, res + torch.nn.functional.linear(act, Wd))
    # fused gate|up layer with the SwiGLU in the epilogue: weight [2*I2, K], out [B, I2]
    I2 = 264
    Wgu = (torch.randn(2 * I2, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    out2 = torch.empty(B, I2, dtype=torch.bfloat16, device="cuda")
    L.check(lib.llm_gemv(x.data_ptr(), Wgu.data_ptr(), 0, nw.data_ptr(), 0, out2.data_ptr(), B, K, 2 * I2, 1, 1e-5, 1, st), "pair")
    gu2 = torch.nn.functional.linear(normed, Wgu)
    close(out2, torch.nn.functional.silu(gu2[:, :I2]) * gu2[:, I2:])
    assert lib.llm_gemv(x.data_ptr(), W.data_ptr(), 0, 0, 0, out.data_ptr(), 9, K, N, 0, 0.0, 0, st) == 1002  # B > 8: unsupported


def test_fused_gemv_decode_step_equals_cublas_step():
    """The generation engine with the decode step on llm_gemv vs the same engine on cuBLAS + stand-alone glue kernels."""
    import sdag_b200.llama_runner as lr
    import sdag_b200.sdag as sdag

    tok = synth.make_tokenizer()
    model = _bf16_model(tok)
    runner = lr.get_runner(sdag.prepare_model(model))
    ids = tok(["The Eiffel Tower was completed in 1889 and stands in Paris.", "Honey never spoils because of its low moisture."],
              return_tensors="pt", padding=True)["input_ids"][:, :40].cuda()
    bm = sdag.causal_blockmap(ids.shape[1], ids.device)
    try:
        lr.USE_FUSED_GEMV = True
        a = runner.generate(ids, bm, 32, None)
        a_eager = runner.generate(ids, bm, 32, None, use_graph=False)
        lr.USE_FUSED_GEMV = False
        b = runner.generate(ids, bm, 32, None)
    finally:
        lr.USE_FUSED_GEMV = True
    assert torch.equal(a, a_eager)
    for row_a, row_b in zip(a.tolist(), b.tolist()):
        fd = next((i for i, (x, y) in enumerate(zip(row_a, row_b)) if x != y), len(row_a))
        assert fd >= 8, (row_a, row_b)
    assert sum(ra == rb for ra, rb in zip(a.tolist(), b.tolist())) >= 1


@pytest.mark.parametrize("M,N,K,ldx_pad", [(300, 264, 72, 0), (1000, 512, 256, 8), (129, 8, 64, 0), (2048, 768, 1024, 0), (1, 256, 4096, 0)])
def test_prefill_gemm_matches_linear(M, N, K, ldx_pad):
    """llm_gemm (CTA-pair tcgen05 GEMM) vs torch's linear on the same bf16 operands, ragged M / N / K edges, strided rows,
    bias; rows and columns outside [M, N] are never written."""
    import sdag_b200._lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    xs = torch.randn(M, K + ldx_pad, generator=g, device="cuda").to(torch.bfloat16)
    x = xs[:, :K]
    W = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16)
    st = L.stream_ptr(x.device)
    for b in (None, bias):
        buf = torch.full((M + 3, N + 8), 7.0, dtype=torch.bfloat16, device="cuda")
        L.check(lib.llm_gemm(x.data_ptr(), xs.stride(0), W.data_ptr(), L.ptr(b), buf.data_ptr(), buf.stride(0), M, N, K, 0, st), "llm_gemm")
        ref = torch.nn.functional.linear(x.float(), W.float(), None if b is None else b.float()).to(torch.bfloat16)
        torch.testing.assert_close(buf[:M, :N].float(), ref.float(), atol=2e-2, rtol=2e-2)
        assert (buf[:M, :N] == ref).float().mean() > 0.97  # one rounding of an fp32 sum; only the summation order differs
        assert (buf[M:] == 7.0).all() and (buf[:, N:] == 7.0).all()
    assert lib.llm_gemm(x.data_ptr(), xs.stride(0), W.data_ptr(), 0, buf.data_ptr(), buf.stride(0), M, N + 1, K, 0, st) == 1001
    assert lib.llm_gemm(0, K, W.data_ptr(), 0, buf.data_ptr(), buf.stride(0), M, N, K, 0, st) == 1001


@pytest.mark.parametrize("M,I,K", [(700, 200, 136), (512, 1024, 512), (260, 64, 64), (3000, 14336 // 8, 4096 // 4)])
def test_prefill_gemm_swiglu_epilogue_matches_mlp_roundings(M, I, K):
    """llm_gemm with the SwiGLU epilogue vs linear -> llm_swiglu (the unfused path, LlamaMLP's bf16 rounding order)."""
    import sdag_b200._lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M + I)
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    Wgu = (torch.randn(2 * I, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    st = L.stream_ptr(x.device)
    out = torch.full((M + 1, I + 8), 7.0, dtype=torch.bfloat16, device="cuda")
    L.check(lib.llm_gemm(x.data_ptr(), K, Wgu.data_ptr(), 0, out.data_ptr(), out.stride(0), M, 2 * I, K, 1, st), "llm_gemm swiglu")
    gu = torch.nn.functional.linear(x.float(), Wgu.float()).to(torch.bfloat16)
    ref = torch.nn.functional.silu(gu[:, :I]) * gu[:, I:]
    torch.testing.assert_close(out[:M, :I].float(), ref.float(), atol=3e-2, rtol=3e-2)
    assert (out[:M, :I] == ref).float().mean() > 0.9
    assert (out[M:] == 7.0).all() and (out[:, I:] == 7.0).all()
    unfused = torch.empty(M, I, dtype=torch.bfloat16, device="cuda")
    L.check(lib.llm_swiglu(gu.data_ptr(), unfused.data_ptr(), M, I, st), "llm_swiglu")
    torch.testing.assert_close(out[:M, :I].float(), unfused.float(), atol=3e-2, rtol=3e-2)
    assert lib.llm_gemm(x.data_ptr(), K, Wgu.data_ptr(), x.data_ptr(), out.data_ptr(), out.stride(0), M, 2 * I, K, 1, st) == 1001  # no bias here


def L_count():
    import sdag_b200._lib as L

    return L.launch_count


def test_fused_prefill_with_pair_gemm_matches_cublas_path(monkeypatch):
    """The prefill engine with every projection on llm_gemm (RoPE / residual / SwiGLU epilogues) vs the same engine on
    cuBLAS + the stand-alone glue kernels."""
    import sdag_b200.llama_runner as lr
    import sdag_b200.sdag as sdag

    tok = synth.make_tokenizer()
    model = _bf16_model(tok)
    runner = lr.get_runner(sdag.prepare_model(model))
    assert runner is not None
    ids = torch.randint(5, len(tok), (2, 384), device="cuda")
    bm = sdag.causal_blockmap(384, ids.device)
    monkeypatch.setattr(lr, "USE_PAIR_GEMM", False)
    ref, _ = runner.prefill(ids, bm, want_cache=False)
    monkeypatch.setattr(lr, "USE_PAIR_GEMM", True)
    launches0 = L_count()
    got, _ = runner.prefill(ids, bm, want_cache=False)
    assert "llm_gemm_residual" in runner.path and L_count() > launches0
    torch.testing.assert_close(got.float(), ref.float(), atol=5e-2, rtol=5e-2)
    assert (got.argmax(-1) == ref.argmax(-1)).all()


@pytest.mark.parametrize("B,bias", [(1, False), (2, True), (4, True)])
def test_decode_qkv_gemv_with_rope_epilogue_equals_two_launch_form(B, bias):
    """llm_gemv_rope (q/k/v projection + RoPE + cache write in one launch) is bit-identical to llm_gemv -> llm_rope_qkv."""
    import sdag_b200._lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(B)
    H, Hkv, D, K, S = 4, 2, 64, 512, 96
    N = (H + 2 * Hkv) * D
    x = torch.randn(B, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bq = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16) if bias else None
    nw = (1 + 0.1 * torch.randn(K, generator=g, device="cuda")).to(torch.bfloat16)
    ang = torch.rand(S, D, generator=g, device="cuda") * 6.28
    cos, sin = ang.cos().to(torch.bfloat16), ang.sin().to(torch.bfloat16)
    pos = torch.tensor([5, 17, 0, 95][:B], dtype=torch.int32, device="cuda")
    st = L.stream_ptr(x.device)

    def buffers():
        return (torch.zeros(B, 1, H, D, dtype=torch.bfloat16, device="cuda"),
                torch.zeros(B, Hkv, S, D, dtype=torch.bfloat16, device="cuda"), torch.zeros(B, Hkv, S, D, dtype=torch.bfloat16, device="cuda"))

    q1, k1, v1 = buffers()
    qkv = torch.empty(B, N, dtype=torch.bfloat16, device="cuda")
    L.check(lib.llm_gemv(x.data_ptr(), W.data_ptr(), L.ptr(bq), nw.data_ptr(), 0, qkv.data_ptr(), B, K, N, 1, 1e-5, 0, st), "gemv")
    L.check(lib.llm_rope_qkv(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), q1.data_ptr(), k1.data_ptr(), v1.data_ptr(), B, 1, H, Hkv, D,
                             k1.stride(0), k1.stride(1), 0, pos.data_ptr(), 1, st), "rope")
    q2, k2, v2 = buffers()
    L.check(lib.llm_gemv_rope(x.data_ptr(), W.data_ptr(), L.ptr(bq), nw.data_ptr(), 1e-5, cos.data_ptr(), sin.data_ptr(), q2.data_ptr(),
                              k2.data_ptr(), v2.data_ptr(), B, K, H, Hkv, D, k2.stride(0), k2.stride(1), 0, pos.data_ptr(), 1, st), "gemv_rope")
    assert q1.abs().sum() > 0 and k1.abs().sum() > 0 and v1.abs().sum() > 0
    assert torch.equal(q1, q2) and torch.equal(k1, k2) and torch.equal(v1, v2)
    assert lib.llm_gemv_rope(x.data_ptr(), W.data_ptr(), 0, nw.data_ptr(), 1e-5, cos.data_ptr(), sin.data_ptr(), q2.data_ptr(),
                             k2.data_ptr(), v2.data_ptr(), 5, K, H, Hkv, D, k2.stride(0), k2.stride(1), 0, pos.data_ptr(), 1, st) == 1002


@pytest.mark.parametrize("B,L,H,Hkv,D,K,bias,pos0", [(2, 300, 4, 2, 128, 256, False, 0), (3, 130, 8, 4, 64, 192, True, 5), (1, 700, 2, 2, 128, 64, True, 0)])
def test_prefill_qkv_gemm_with_rope_epilogue_equals_two_launch_form(B, L, H, Hkv, D, K, bias, pos0):
    """llm_gemm_rope (q/k/v projection + RoPE + cache write in the GEMM epilogue) vs llm_gemm -> llm_rope_qkv: same
    accumulators, same roundings, so the three outputs are bit-identical; cache rows outside [pos0, pos0 + L) stay untouched."""
    import sdag_b200._lib as L_

    lib = L_.load()
    g = torch.Generator(device="cuda").manual_seed(B * L)
    N, M, S = (H + 2 * Hkv) * D, B * L, pos0 + L + 7
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bq = torch.randn(N, generator=g, device="cuda").to(torch.bfloat16) if bias else None
    ang = torch.rand(S, D, generator=g, device="cuda") * 6.28
    cos, sin = ang.cos().to(torch.bfloat16), ang.sin().to(torch.bfloat16)
    st = L_.stream_ptr(x.device)

    def buffers():
        return (torch.full((B, L, H, D), 3.0, dtype=torch.bfloat16, device="cuda"),
                torch.full((B, Hkv, S, D), 3.0, dtype=torch.bfloat16, device="cuda"), torch.full((B, Hkv, S, D), 3.0, dtype=torch.bfloat16, device="cuda"))

    q1, k1, v1 = buffers()
    qkv = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    L_.check(lib.llm_gemm(x.data_ptr(), K, W.data_ptr(), L_.ptr(bq), qkv.data_ptr(), N, M, N, K, 0, st), "gemm")
    L_.check(lib.llm_rope_qkv(qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), q1.data_ptr(), k1.data_ptr(), v1.data_ptr(), B, L, H, Hkv, D,
                              k1.stride(0), k1.stride(1), pos0, 0, 0, st), "rope")
    q2, k2, v2 = buffers()
    L_.check(lib.llm_gemm_rope(x.data_ptr(), K, W.data_ptr(), L_.ptr(bq), cos.data_ptr(), sin.data_ptr(), q2.data_ptr(), k2.data_ptr(),
                               v2.data_ptr(), B, L, H, Hkv, D, K, k2.stride(0), k2.stride(1), pos0, st), "gemm_rope")
    assert torch.equal(q1, q2) and torch.equal(k1, k2) and torch.equal(v1, v2)
    assert (k2[:, :, :pos0] == 3.0).all() and (k2[:, :, pos0 + L:] == 3.0).all() and (v2[:, :, pos0 + L:] == 3.0).all()
    assert not (q2 == 3.0).all()
    # geometry the tiles cannot hold: 3 heads of 128 (odd), head_dim 96
    assert lib.llm_gemm_rope(x.data_ptr(), K, W.data_ptr(), 0, cos.data_ptr(), sin.data_ptr(), q2.data_ptr(), k2.data_ptr(), v2.data_ptr(),
                             B, L, 3, 1, 128, K, k2.stride(0), k2.stride(1), 0, st) == 1002
    assert lib.llm_gemm_rope(x.data_ptr(), K, W.data_ptr(), 0, cos.data_ptr(), sin.data_ptr(), q2.data_ptr(), k2.data_ptr(), v2.data_ptr(),
                             B, L, H, Hkv, 96, K, k2.stride(0), k2.stride(1), 0, st) == 1002


def test_prefill_gemm_wide_tile_variant_in_subprocess():
    """The 256 x 512 pair-tile variant of llm_gemm (SDAG_GEMM_WIDE=1, read once per process) against torch on a plain and a
    SwiGLU shape with ragged edges."""
    import subprocess
    import sys

    code = r'''
import torch, sdag_b200._lib as L
lib = L.load(); st = L.stream_ptr(torch.device("cuda:0"))
g = torch.Generator(device="cuda").manual_seed(3)
for M, N, K, epi in ((700, 1288, 200, 0), (1000, 1040, 136, 1), (300, 512, 64, 0)):
    x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
    n_out = N // 2 if epi else N
    out = torch.full((M + 2, n_out + 8), 7.0, dtype=torch.bfloat16, device="cuda")
    L.check(lib.llm_gemm(x.data_ptr(), K, W.data_ptr(), 0, out.data_ptr(), out.stride(0), M, N, K, epi, st), "llm_gemm wide")
    y = torch.nn.functional.linear(x.float(), W.float()).to(torch.bfloat16)
    ref = torch.nn.functional.silu(y[:, :n_out]) * y[:, n_out:] if epi else y
    torch.testing.assert_close(out[:M, :n_out].float(), ref.float(), atol=3e-2, rtol=3e-2)
    assert (out[:M, :n_out] == ref).float().mean() > 0.9
    assert (out[M:] == 7.0).all() and (out[:, n_out:] == 7.0).all()
# the residual epilogue on wide tiles (only reachable with the knob: the default keeps 256 x 256 tiles for it)
M, N, K = 2304, 520, 8192
x = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
W = (torch.randn(N, K, generator=g, device="cuda") * K ** -0.5).to(torch.bfloat16)
res = torch.randn(M, N, generator=g, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
L.check(lib.llm_gemm_residual(x.data_ptr(), K, W.data_ptr(), 0, res.data_ptr(), N, out.data_ptr(), N, M, N, K, st), "llm_gemm_residual wide")
ref = (torch.nn.functional.linear(x.float(), W.float()).to(torch.bfloat16).float() + res.float()).to(torch.bfloat16)
torch.testing.assert_close(out.float(), ref.float(), atol=3e-2, rtol=3e-2)
assert (out == ref).float().mean() > 0.9
print("wide ok")
'''
    env = dict(os.environ, SDAG_GEMM_WIDE="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "wide ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("choice", [{"o": "mma", "down": "mma"},
                                    {"qkv": "skinny", "o": "skinny", "gate_up": "skinny", "down": "skinny", "head": "skinny"}])
def test_decode_step_kernel_choices_agree_with_the_default_step(choice, monkeypatch):
    """``SDAG_DECODE_GEMV`` (opt-in): the decode step of 3-8 prompts with its linear layers on ``llm_gemv``'s warp-MMA
    form / on ``llm_skinny_linear`` (RMSNorm prologue, SwiGLU and residual epilogues fused, K split over a cluster)
    against the default step (cuBLAS + stand-alone glue) on the same ragged batch: the first decode tokens equal, later
    ones up to a bf16 near-tie flip (different GEMM orders), as between any two batch shapes of this engine."""
    import sdag_b200.llama_runner as lr
    import sdag_b200.sdag as sdag
    from sdag_b200.generation import generate_batch_seq2seq

    tok = synth.make_tokenizer()
    tok.padding_side = "left"
    model = _bf16_model(tok)
    assert lr.get_runner(sdag.prepare_model(model)) is not None
    prompts = ["Paris is the capital of", "Honey", "The highest mountain above sea level is called Mount Everest and it",
               "Octopuses have three hearts"]
    n_new = 16

    def run():
        seen = []
        orig = _spy_decode(tok, seen)
        try:
            generate_batch_seq2seq(tok, model, prompts, max_tokens=n_new, temperature=0, top_p=1.0)
        finally:
            tok.decode = orig
        return list(seen)

    base = run()
    launches0 = dict(lr._lib.launch_counts) if hasattr(lr._lib, "launch_counts") else None
    monkeypatch.setattr(lr, "DECODE_GEMV", dict(choice))  # part of the decode-context key: a new graph is captured
    mixed = run()
    ctx = lr.get_runner(model)._last_ctx
    assert ctx["graph"] is not None and ctx["step"].__name__ == "step_mixed"
    del launches0
    for b, (x, y) in enumerate(zip(base, mixed)):
        d = next((i for i, (p, q) in enumerate(zip(x, y)) if p != q), min(len(x), len(y)))
        assert d >= min(4, len(x)), (b, x, y)
    assert sum(x == y for x, y in zip(base, mixed)) >= len(prompts) - 1, (base, mixed)
