"""GPU parity of the prefill linear layers (llm_gemm / llm_gemm_residual / llm_gemm_rope) against fp32 torch: small
ragged shapes exhaustively, and the production shapes of BASELINE config 2 (21 136 token rows, Llama-3-8B geometry) on
sampled rows x columns — the shapes bench.py times.  Tolerance: the output is bf16, so |err| <= 1 bf16 ulp of the value
(2^-8 relative) + the fp32 accumulation-order noise; stated as atol = rtol = 2e-2 with > 90 % of entries bit-equal to the
bf16 rounding of the fp32 reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = dict(atol=2e-2, rtol=2e-2)


def _lib():
    import sdag_b200._lib as L

    return L, L.load()


def _rand(shape, g, scale=1.0):
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K,bias", [(1, 8, 8, False), (300, 520, 72, True), (700, 1288, 200, False), (257, 256, 4096, True),
                                        (2304, 520, 8192, False)])  # long K (the wide tiles' range; a residual keeps 256 x 256 tiles)
def test_llm_gemm_residual_matches_torch_and_may_alias_out(M, N, K, bias):
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x, W = _rand((M, K), g), _rand((N, K), g, K ** -0.5)
    b = _rand((N,), g) if bias else None
    res = _rand((M + 1, N + 8), g)  # padded rows: ldr != N
    out = torch.full((M + 2, N + 16), 7.0, dtype=torch.bfloat16, device="cuda")
    st = L.stream_ptr(x.device)
    L.check(lib.llm_gemm_residual(x.data_ptr(), K, W.data_ptr(), L.ptr(b), res.data_ptr(), res.stride(0), out.data_ptr(),
                                  out.stride(0), M, N, K, st), "llm_gemm_residual")
    lin = torch.nn.functional.linear(x.float(), W.float(), None if b is None else b.float()).to(torch.bfloat16)
    ref = (lin.float() + res[:M, :N].float()).to(torch.bfloat16)
    torch.testing.assert_close(out[:M, :N].float(), ref.float(), **TOL)
    assert (out[:M, :N] == ref).float().mean() > 0.9
    assert (out[M:] == 7.0).all() and (out[:, N:] == 7.0).all()  # nothing written outside [M, N]
    # in place: the residual stream is updated where it lives
    inplace = res.clone()
    L.check(lib.llm_gemm_residual(x.data_ptr(), K, W.data_ptr(), L.ptr(b), inplace.data_ptr(), inplace.stride(0), inplace.data_ptr(),
                                  inplace.stride(0), M, N, K, st), "llm_gemm_residual in place")
    assert torch.equal(inplace[:M, :N], out[:M, :N])
    assert torch.equal(inplace[M:], res[M:]) and torch.equal(inplace[:, N:], res[:, N:])
    # argument checks: misaligned residual stride, residual with a non-plain epilogue is not expressible
    assert lib.llm_gemm_residual(x.data_ptr(), K, W.data_ptr(), 0, res.data_ptr(), N + 4, out.data_ptr(), out.stride(0), M, N, K, st) == 1001
    assert lib.llm_gemm_residual(x.data_ptr(), K, W.data_ptr(), 0, 0, N, out.data_ptr(), out.stride(0), M, N, K, st) == 1001


def _sample(n, count, g):
    idx = torch.randperm(n, generator=g, device="cuda")[:count].sort().values
    return idx


# BASELINE config 2, 8 prompts x 2642 tokens = 21 136 rows: the shapes of one decoder layer + the LM head on 8 rows
PROD_M = 21136


@pytest.mark.parametrize("name,M,N,K,kind", [
    ("o_proj + residual", PROD_M, 4096, 4096, "residual"),
    ("down_proj + residual", PROD_M, 4096, 14336, "residual"),
    ("down_proj, plain epilogue (wide tiles)", PROD_M, 4096, 14336, "plain"),
    ("gate/up + SwiGLU", PROD_M, 28672, 4096, "swiglu"),
    ("lm_head on the last rows", 8, 128256, 4096, "plain"),
])
def test_llm_gemm_production_shapes_on_sampled_tiles(name, M, N, K, kind):
    """The launches bench.py times, at their full size; verified on 96 sampled rows x 384 sampled output columns (plus
    the first / last rows and columns) against an fp32 product of the same bf16 operands."""
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(N + K)
    x, W = _rand((M, K), g), _rand((N, K), g, K ** -0.5)
    st = L.stream_ptr(x.device)
    n_out = N // 2 if kind == "swiglu" else N
    out = torch.empty((M, n_out), dtype=torch.bfloat16, device="cuda")
    res = _rand((M, n_out), g) if kind == "residual" else None
    if kind == "residual":
        L.check(lib.llm_gemm_residual(x.data_ptr(), K, W.data_ptr(), 0, res.data_ptr(), n_out, out.data_ptr(), n_out, M, N, K, st), name)
    else:
        L.check(lib.llm_gemm(x.data_ptr(), K, W.data_ptr(), 0, out.data_ptr(), n_out, M, N, K, 1 if kind == "swiglu" else 0, st), name)
    rows = torch.unique(torch.cat([_sample(M, min(M, 96), g), torch.tensor([0, M - 1], device="cuda")]))
    cols = torch.unique(torch.cat([_sample(n_out, 384, g), torch.tensor([0, n_out - 1], device="cuda")]))
    xr = x[rows].float()
    if kind == "swiglu":
        gate = (xr @ W[cols].float().T).to(torch.bfloat16)
        up = (xr @ W[cols + n_out].float().T).to(torch.bfloat16)
        ref = torch.nn.functional.silu(gate) * up
    else:
        ref = (xr @ W[cols].float().T).to(torch.bfloat16)
        if kind == "residual":
            ref = (ref.float() + res[rows][:, cols].float()).to(torch.bfloat16)
    got = out[rows][:, cols]
    torch.testing.assert_close(got.float(), ref.float(), **TOL)
    assert (got == ref).float().mean() > 0.9, name


def test_llm_gemm_rope_production_shape_on_sampled_rows():
    """q/k/v projection + RoPE + KV-cache write at 8 x 2642 tokens, Llama-3-8B geometry (32 q / 8 kv heads, d = 128, K = 4096):
    sampled token rows against fp32 linear -> bf16 -> HF apply_rotary_pos_emb in bf16."""
    L, lib = _lib()
    B, Lp, H, Hkv, D, K = 8, 2642, 32, 8, 128, 4096
    N, M = (H + 2 * Hkv) * D, B * Lp
    g = torch.Generator(device="cuda").manual_seed(11)
    x, W = _rand((M, K), g), _rand((N, K), g, K ** -0.5)
    inv = 1.0 / (500000.0 ** (torch.arange(0, D, 2, device="cuda").float() / D))
    ang = torch.arange(Lp, device="cuda").float()[:, None] * inv[None]
    emb = torch.cat([ang, ang], dim=-1)
    cos, sin = emb.cos().to(torch.bfloat16), emb.sin().to(torch.bfloat16)
    q = torch.empty((B, Lp, H, D), dtype=torch.bfloat16, device="cuda")
    k = torch.empty((B, Hkv, Lp, D), dtype=torch.bfloat16, device="cuda")
    v = torch.empty_like(k)
    st = L.stream_ptr(x.device)
    L.check(lib.llm_gemm_rope(x.data_ptr(), K, W.data_ptr(), 0, cos.data_ptr(), sin.data_ptr(), q.data_ptr(), k.data_ptr(), v.data_ptr(),
                              B, Lp, H, Hkv, D, K, k.stride(0), k.stride(1), 0, st), "llm_gemm_rope")
    rows = torch.unique(torch.cat([_sample(M, 64, g), torch.tensor([0, Lp - 1, Lp, M - 1], device="cuda")]))
    qkv = (x[rows].float() @ W.float().T).to(torch.bfloat16).view(-1, H + 2 * Hkv, D)
    pos = rows % Lp
    c, s = cos[pos][:, None], sin[pos][:, None]

    def rot(t):  # HF apply_rotary_pos_emb on bf16 tensors
        half = torch.cat([-t[..., D // 2:], t[..., : D // 2]], dim=-1)
        return t * c + half * s

    q_ref, k_ref, v_ref = rot(qkv[:, :H]), rot(qkv[:, H:H + Hkv]), qkv[:, H + Hkv:]
    b_idx = rows // Lp
    torch.testing.assert_close(q.view(M, H, D)[rows].float(), q_ref.float(), **TOL)
    torch.testing.assert_close(k[b_idx, :, pos].float(), k_ref.float(), **TOL)
    torch.testing.assert_close(v[b_idx, :, pos].float(), v_ref.float(), **TOL)
    assert (v[b_idx, :, pos] == v_ref).float().mean() > 0.9


@pytest.mark.parametrize("B,N,K,mode", [
    (1, 4096, 4096, "residual"), (3, 4096, 14336, "residual"), (8, 6144, 4096, "norm+bias"), (4, 28672, 4096, "norm+swiglu"),
    (2, 128256, 4096, "norm"), (8, 896 + 256, 896, "norm+bias"), (5, 9728, 896, "norm+swiglu"), (7, 896, 4864, "residual"),
    (17, 1000, 192, "plain"), (32, 4096, 4096, "residual"), (8, 4096, 7168, "norm"),
])
def test_llm_skinny_linear_matches_torch(B, N, K, mode):
    """The decode step's few-row linear layers as a TMA weight stream on tcgen05 (csrc/llm_skinny.cu): Llama-3-8B and
    Qwen2.5-0.5B layer shapes, 1..32 prompts, every prologue / epilogue form, padded row strides; against an fp32 product of
    the same bf16 operands with the Hugging Face roundings (LlamaRMSNorm, nn.Linear -> bf16, residual add, LlamaMLP)."""
    L, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(B * 7 + N)
    x = _rand((B, K + 8), g)[:, :K]  # ldx != K
    W = _rand((N, K), g, K ** -0.5)
    norm = "norm" in mode
    pair = "swiglu" in mode
    nw = (1.0 + 0.1 * torch.randn(K, generator=g, device="cuda")).to(torch.bfloat16) if norm else None
    bias = _rand((N,), g) if "bias" in mode else None
    res = _rand((B, N + 16), g)[:, :N] if mode == "residual" else None
    n_out = N // 2 if pair else N
    out = torch.full((B + 1, n_out + 8), 7.0, dtype=torch.bfloat16, device="cuda")
    st = L.stream_ptr(x.device)
    rc = lib.llm_skinny_linear(x.data_ptr(), x.stride(0), W.data_ptr(), L.ptr(bias), L.ptr(nw), 1e-5, L.ptr(res),
                               res.stride(0) if res is not None else 0, out.data_ptr(), out.stride(0), B, N, K, int(pair), st)
    L.check(rc, "llm_skinny_linear")
    xf = x.float()
    if norm:
        xf = (nw.float() * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-5)).to(torch.bfloat16).float()).to(torch.bfloat16).float()
    lin = xf @ W.float().T
    if bias is not None:
        lin = lin + bias.float()
    lin = lin.to(torch.bfloat16)
    if pair:
        ref = (torch.nn.functional.silu(lin[:, :n_out].float()).to(torch.bfloat16).float() * lin[:, n_out:].float()).to(torch.bfloat16)
    elif res is not None:
        ref = (lin.float() + res.float()).to(torch.bfloat16)
    else:
        ref = lin
    got = out[:B, :n_out]
    torch.testing.assert_close(got.float(), ref.float(), **TOL)
    assert (got == ref).float().mean() > 0.9
    assert (out[B:] == 7.0).all() and (out[:, n_out:] == 7.0).all()  # nothing written outside [B, n_out]
    if res is not None:  # in place: out aliases the residual
        inplace = res.clone()
        L.check(lib.llm_skinny_linear(x.data_ptr(), x.stride(0), W.data_ptr(), 0, 0, 0.0, inplace.data_ptr(), inplace.stride(0),
                                      inplace.data_ptr(), inplace.stride(0), B, N, K, 0, st), "in place")
        assert torch.equal(inplace, got)
