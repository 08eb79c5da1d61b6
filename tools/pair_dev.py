#!/usr/bin/env python
"""Bring-up driver of the CTA-pair prefill kernel: a ladder of small cases (each against torch SDPA / the single-CTA kernel),
printing max-abs errors, then timings.  Run each step in its own process under `timeout` on the GPU box."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdag_b200.sdag as sdag

def case(L, B, H, Hkv, layout=None, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(B, H, L, 128, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Hkv, L, 128, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Hkv, L, 128, generator=g, device="cuda").to(torch.bfloat16)
    bm = sdag.causal_blockmap(L, q.device) if layout is None else sdag.build_blockmap([sdag.SdagLayout(L, *layout)] * B)
    ref = sdag.sdag_prefill_attention(q, k, v, bm, kernel="single")
    torch.cuda.synchronize()
    out = sdag.sdag_prefill_attention(q, k, v, bm, kernel="pair")
    torch.cuda.synchronize()
    d = (out.float() - ref.float()).abs()
    print(f"L={L} B={B} H={H} Hkv={Hkv} layout={'causal' if layout is None else 'spans'}: max|pair - single| = {d.max().item():.4f} "
          f"mean {d.mean().item():.5f} finite={bool(torch.isfinite(out.float()).all())}", flush=True)
    if d.max().item() > 0.05:
        bad = (d > 0.05).nonzero()
        print("  first bad (b, l, h, d):", bad[:5].tolist(), " rows with errors:", sorted(set(bad[:, 1].tolist()))[:20], flush=True)
    return q, k, v, bm

if __name__ == "__main__":
    step = sys.argv[1] if len(sys.argv) > 1 else "all"
    if step in ("1", "all"):
        case(128, 1, 2, 1)          # one item, one tile
    if step in ("2", "all"):
        case(256, 1, 2, 1)          # two tiles, diagonal + full
        case(1000, 1, 4, 2)
    if step in ("3", "all"):
        case(300, 2, 8, 2)          # G = 4: short tiles pack two heads per CTA
        from bench import synthetic_layout
        L, S, spans, qa = synthetic_layout(48, 10, 256, 1, 24)
        case(L, 2, 8, 2, (S, spans, qa))
    if step in ("4", "all"):
        for L, B in ((8192, 4), (2642, 8)):
            from bench import synthetic_layout
            lay = None
            if L == 2642:
                _, S, spans, qa = synthetic_layout(48, 10, 256, 1, 24); lay = (S, spans, qa)
            q, k, v, bm = case(L, B, 32, 8, lay)
            for kern in ("single", "pair"):
                for _ in range(3): sdag.sdag_prefill_attention(q, k, v, bm, kernel=kern)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(10): sdag.sdag_prefill_attention(q, k, v, bm, kernel=kern)
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 10
                fl = 4.0 * 128 * bm.header(0)["pairs"] * 32 * B
                print(f"  {kern}: {ms:.4f} ms  {fl / ms / 1e9:.0f} TFLOP/s (algorithmic)", flush=True)
