"""llm_gemm (CTA-pair tcgen05 GEMM) vs cuBLAS on the Llama-3-8B prefill shapes.

    python tools/gemm_bench.py [--tokens 21248]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdag_b200  # noqa: E402
from sdag_b200 import _lib  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def sustained(fn, flops, seconds=1.5):
    """TFLOP/s over the second half of a `seconds`-long back-to-back run (power-capped clocks, not burst clocks)."""
    t1 = timed(fn, 3)
    n = max(4, int(seconds * 1e3 / t1 / 2))
    for _ in range(n):
        fn()
    t2 = timed(fn, n)
    return flops / t2 / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, default=21248)
    ap.add_argument("--only", default="", help="run one shape by name (qkv, o, gate_up, gate_up+swiglu, down)")
    ap.add_argument("--sustain", action="store_true", help="also report TFLOP/s of 1.5 s back-to-back runs")
    ap.add_argument("--encoder", action="store_true", help="the e5-large-v2 encoder's projections (hidden 1024, FFN 4096) instead")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    M = args.tokens
    st = _lib.stream_ptr(dev)
    torch.manual_seed(0)
    llama = (("qkv", 6144, 4096, 0), ("o", 4096, 4096, 0), ("gate_up", 28672, 4096, 0), ("gate_up+swiglu", 28672, 4096, 1),
             ("down", 4096, 14336, 0), ("o+residual", 4096, 4096, 2), ("down+residual", 4096, 14336, 2))
    # epilogue 3: bias + erf-GELU (llm_gemm epilogue code 2), 4: bias + residual
    e5 = (("enc qkv+bias", 3072, 1024, 5), ("enc o+bias+res", 1024, 1024, 4), ("enc ffn1+gelu", 4096, 1024, 3), ("enc ffn2+bias+res", 1024, 4096, 4))
    for name, N, K, epi in (e5 if args.encoder else llama):
        if args.only and name != args.only:
            continue
        x = torch.randn(M, K, device=dev, dtype=torch.bfloat16) * 0.5
        w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02
        n_out = N // 2 if epi == 1 else N
        out = torch.empty(M, n_out, device=dev, dtype=torch.bfloat16)
        res = torch.randn(M, n_out, device=dev, dtype=torch.bfloat16) if epi in (2, 4) else None
        bias = torch.randn(N, device=dev, dtype=torch.bfloat16) if epi >= 3 else None

        def ours():
            if epi in (2, 4):  # residual add in the epilogue (what the prefill runs for o / down)
                _lib.check(lib.llm_gemm_residual(x.data_ptr(), K, w.data_ptr(), _lib.ptr(bias), res.data_ptr(), n_out, out.data_ptr(), n_out,
                                                 M, N, K, st), "llm_gemm_residual")
            elif epi in (3, 5):
                _lib.check(lib.llm_gemm(x.data_ptr(), K, w.data_ptr(), bias.data_ptr(), out.data_ptr(), n_out, M, N, K, 2 if epi == 3 else 0, st),
                           "llm_gemm")
            else:
                _lib.check(lib.llm_gemm(x.data_ptr(), K, w.data_ptr(), 0, out.data_ptr(), n_out, M, N, K, epi, st), "llm_gemm")

        act = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)

        def cublas():
            y = torch.nn.functional.linear(x, w, bias)
            if epi == 3:
                y = torch.nn.functional.gelu(y)
            if epi == 1:
                _lib.check(lib.llm_swiglu(y.data_ptr(), act.data_ptr(), M, N // 2, st), "swiglu")
            elif epi in (2, 4):
                y = y + res
            return y

        t_o, t_c = timed(ours), timed(cublas)
        ref = cublas()
        ref = act if epi == 1 else ref
        ours()
        err = (out.float() - ref.float()).abs().max().item()
        fl = 2.0 * M * N * K
        print(f"{name:16s} M={M} N={N} K={K}: llm_gemm {t_o:.3f} ms ({fl / t_o / 1e9:.0f} TFLOP/s)  cuBLAS{'+swiglu' if epi == 1 else ('+add' if epi == 2 else '')} "
              f"{t_c:.3f} ms ({fl / t_c / 1e9:.0f} TFLOP/s)  max|diff| {err:.4f}", flush=True)
        if args.sustain:
            print(f"{'':16s} sustained: llm_gemm {sustained(ours, fl):.0f}  cuBLAS {sustained(cublas, fl):.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
