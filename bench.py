#!/usr/bin/env python
"""Benchmark of the two hot paths (BASELINE.json: "SDAG prefill tok/s (Llama-3-8B, k=10); dense top-k QPS,
10M x 1024 corpus").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (port, bounded sample)

One JSON line on stdout (rank 0).  Primary metric = SDAG prefill tokens/s on BASELINE config 2 (Llama-3-8B
geometry, bf16, random init, TOP_K=10 x 256-token documents, DOC_NEIGHBORS_K=0, L = 2642); the dense top-k QPS
over the 10M x 1024 corpus rides along under "secondary".  A "step" is one SDAG prefill of a batch of prompts
(+ one 4096-query search batch for the secondary metric, timed separately).

Multi-GPU: SDAG prefill is partitioned by independent query batches — no collective, weak scaling; the
retrieval corpus is sharded by rows over the ranks with one all-gather + merge (fixed 10M rows: strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SDAG prefill tok/s (Llama-3-8B, k=10)"
SECONDARY_METRIC = "dense top-k QPS, 10M x 1024 corpus"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="prompts per SDAG prefill step (per GPU)")
    ap.add_argument("--layers", type=int, default=32, help="decoder layers (32 = Llama-3-8B; fewer only for debugging)")
    ap.add_argument("--corpus-rows", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--carg", action="store_true", help="also time plain causal (CARG) prefill through the same kernels")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        p["_source"] = "measured (MEASURED_PEAKS.json)"
        return p
    p = dict(FALLBACK_PEAKS)
    p["_source"] = "fallback (B200_PROFILING.md)"
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (the profiling recipe's clocks line)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()  # exact PID we started
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def synthetic_layout(sys_user_len, n_docs, doc_len, gap, qa_len):
    """Fixed-shape prompt layout of SURVEY.md §8d: (L, S, doc spans, qa_start)."""
    spans, pos = [], sys_user_len
    for _ in range(n_docs):
        spans.append((pos, pos + doc_len))
        pos += doc_len + gap
    return pos + qa_len, sys_user_len, spans, pos


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local, world


# ------------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (port), bounded sample per step
# ------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle.cpu_reference import SdagPrefillCpu, flat_ip_cpu

    L, S, spans, qa = synthetic_layout(48, 10, 256, 1, 24)
    job = SdagPrefillCpu(L, S, spans, qa, n_layers=1, vocab_div=8)
    for _ in range(args.warmup):
        job.step()
    per_prompt, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = job.step()
        per_prompt.append(last["seconds_per_prompt"])
    wall = time.perf_counter() - t0
    mean_pp = sum(per_prompt) / len(per_prompt)
    value = L / mean_pp
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * mean_pp, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[1]: SDAG prefill, Llama-3-8B geometry random-init, TOP_K=10 x 256-token docs, "
                               "DOC_NEIGHBORS_K=0, L=2642 (S=48, 1 gap token per doc, 24 QA tokens), batch of 1 (reference loop)",
                   "device": "host CPU", "sampled_step_seconds": wall / max(args.steps, 1)},
        "cpu_baseline": {"value": value, "unit": "tok/s", "cores": job.threads, "kind": "port", "sample": job.describe(last)},
        "e2e": {"value": value, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_secondary:
        r = flat_ip_cpu(100_000, 1024, args.queries, args.topk, args.corpus_rows)
        line["secondary"] = {"metric": SECONDARY_METRIC, "value": r["qps"], "unit": "queries/s",
                             "cpu_baseline": {"value": r["qps"], "unit": "queries/s", "cores": r["threads"], "kind": "port",
                                              "sample": r["sample"]}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def build_llama(layers: int, device):
    import torch
    from transformers import LlamaConfig, LlamaForCausalLM

    cfg = LlamaConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=layers,
                      num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192, rope_theta=500000.0,
                      tie_word_embeddings=False)
    torch.manual_seed(42)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(prev)
    return model.eval()


def run_b200(args):
    import torch
    import torch.distributed as dist

    import sdag_b200._lib as L_
    import sdag_b200.llama_runner as llama_runner
    import sdag_b200.sdag as sdag
    from sdag_b200.retrieval import FlatIPIndex, ShardedFlatIPIndex

    rank, local, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    L_.load()
    pk = peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- SDAG prefill, config 2 ----------------
    Lp, S, spans, qa = synthetic_layout(48, 10, 256, 1, 24)
    B = args.batch
    model = sdag.prepare_model(build_llama(args.layers, device))
    g = torch.Generator().manual_seed(42 + rank)
    ids_host = torch.randint(0, 128000, (B, Lp), generator=g, dtype=torch.int64).pin_memory()
    layouts = [sdag.SdagLayout(Lp, S, spans, qa) for _ in range(B)]
    ids_dev = ids_host.to(device)
    bm = sdag.build_blockmap(layouts, device)
    hdr = bm.header(0)
    assert hdr["status"] == 0 and hdr["pairs"] == 516148, hdr  # P of SURVEY.md §8d for this layout
    H, D = 32, 128
    attn_flops_per_launch = 4.0 * D * hdr["pairs"] * H * B  # algorithmic: 4*D*P*H per prompt per layer

    def step_resident(meta):
        logits, _ = sdag.prefill_forward(model, ids_dev, meta)
        return logits.argmax(dim=-1)

    def step_e2e():
        toks, _ = sdag.prefill_batch(model, ids_host, layouts, device)
        return toks

    def timed(fn, steps, warmup, collect_attn=False):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        if collect_attn:
            sdag._profile_events = []
            llama_runner._gemm_events = []
        launches0 = L_.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        attn_ms = None
        if collect_attn:
            evs, sdag._profile_events = sdag._profile_events, None
            attn_ms = [a.elapsed_time(b) for a, b in evs]
            gevs, llama_runner._gemm_events = llama_runner._gemm_events, None
            gemm_samples[:] = [(a.elapsed_time(b), rows) for a, b, rows in gevs]
        return max_over_ranks(ms), clocks, L_.launch_count - launches0, attn_ms

    gemm_samples = []  # (ms, token rows) of the gate/up llm_gemm launches of the last timed run

    def bad_clocks(c):
        return any(r in c.get("reasons", []) for r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"))

    ms, clocks, launches, attn_ms = timed(lambda: step_resident(bm), args.steps, args.warmup, collect_attn=True)
    remeasured = False
    if bad_clocks(clocks):
        ms, clocks, launches, attn_ms = timed(lambda: step_resident(bm), args.steps, max(args.warmup, 3), collect_attn=True)
        remeasured = True
    tokens_per_step = B * Lp * world
    value = tokens_per_step * args.steps / (ms / 1000.0)
    attn_avg_ms = sum(attn_ms) / len(attn_ms)
    achieved = attn_flops_per_launch / (attn_avg_ms / 1000.0) / 1e12
    peak = pk.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"])

    ms_e2e, _, _, _ = timed(step_e2e, args.steps, args.warmup)
    e2e_value = tokens_per_step * args.steps / (ms_e2e / 1000.0)

    # the same kernel on unmasked (dense causal) tiles: north_star's ">= 60 % of dense bf16 tensor peak on the unmasked
    # tiles" — 4 prompts x 8192 tokens, 32 heads, inputs (403 MB of q/k/v/out) larger than L2, CUDA events on the stream
    def dense_tiles_rate():
        Ld, Bd = 8192, 4
        gq = torch.Generator(device=device).manual_seed(1)
        qd = torch.randn(Bd, Ld, H, D, generator=gq, device=device).to(torch.bfloat16).transpose(1, 2)
        kd = torch.randn(Bd, 8, Ld, D, generator=gq, device=device).to(torch.bfloat16)
        vd = torch.randn(Bd, 8, Ld, D, generator=gq, device=device).to(torch.bfloat16)
        bmd = sdag.causal_blockmap(Ld, device)
        pairs = bmd.header(0)["pairs"]
        for _ in range(3):
            sdag.sdag_prefill_attention(qd, kd, vd, bmd)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            sdag.sdag_prefill_attention(qd, kd, vd, bmd)
        b.record()
        torch.cuda.synchronize()
        ms_d = a.elapsed_time(b) / 10
        fl = 4.0 * D * pairs * H * Bd
        return {"workload": "dense causal, 4 x 8192 tokens, 32 q / 8 kv heads, d=128 (no masked tiles to skip)",
                "launch_ms": ms_d, "flops_per_launch": fl, "achieved": fl / (ms_d / 1000.0) / 1e12, "peak": peak,
                "unit": "TFLOP/s", "frac": fl / (ms_d / 1000.0) / 1e12 / peak}

    dense_tiles = dense_tiles_rate()

    # the launch that takes the largest share of the step (ncu launch list: ~50 %): the gate/up projection with the
    # SwiGLU epilogue on this repo's CTA-pair GEMM, timed live with CUDA events inside the timed region above
    step_dominant = None
    if gemm_samples:
        inter, hid = model.config.intermediate_size, model.config.hidden_size
        g_ms = sum(m for m, _ in gemm_samples) / len(gemm_samples)
        g_fl = sum(2.0 * r * 2 * inter * hid for _, r in gemm_samples) / len(gemm_samples)
        step_dominant = {"kernel": "llm_gemm_kernel<SWIGLU> (gate/up projection + SwiGLU epilogue, tcgen05 cta_group::2)",
                         "bound": "tensor", "launch_ms": g_ms, "flops_per_launch": g_fl,
                         "achieved": g_fl / (g_ms / 1000.0) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": g_fl / (g_ms / 1000.0) / 1e12 / peak,
                         "share_of_step": g_ms * len(gemm_samples) / args.steps / (ms / args.steps)}

    carg = None
    if args.carg:
        ms_c, _, _, _ = timed(lambda: step_resident(None), args.steps, args.warmup)
        carg = tokens_per_step * args.steps / (ms_c / 1000.0)

    line = {
        "metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {
            "workload": "configs[1]: SDAG prefill, Llama-3-8B geometry (32 layers, 32 q / 8 kv heads, d=128) bf16 random-init, "
                        "TOP_K=10 x 256-token docs, DOC_NEIGHBORS_K=0, L=2642 (S=48, 1 gap token per doc, 24 QA tokens)",
            "prompts_per_step_per_gpu": B, "tokens_per_step": tokens_per_step, "layers": args.layers,
            "partition": "independent query batches per GPU, weights replicated, no collective",
            "cache": "inputs larger than L2 (16 GB of weights stream through every step)",
            "linear_layers": "gate/up projection + SwiGLU and q/k/v projection + RoPE + KV write on this repo's CTA-pair tcgen05 GEMM "
                             "(llm_gemm, llm_gemm_rope); o and down projections on cuBLAS via torch (plain library GEMMs); "
                             "attention, block map, RMSNorm: this repo's kernels (llama_runner.FusedPrefill)",
            "remeasured_for_clocks": remeasured,
        },
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at 8 prompts
                     # (profiles/r01_ncu_sdag_prefill_cfg2.md); algorithmic q+k+v+out bytes are 433 MB
                     "traffic": 508.5e6 * B / 8 if B == 8 else None,
                     "kernel": "sdag_prefill_kernel<128> (+ v_colsum, v_mean) per layer launch",
                     "launch_ms": attn_avg_ms, "flops_per_launch": attn_flops_per_launch,
                     "peak_source": pk["_source"] + ", bf16 sustained", "unmasked_tiles": dense_tiles,
                     "step_dominant_kernel": step_dominant},
        "e2e": {"value": e2e_value, "unit": "tok/s", "h2d_bytes_per_step": int(ids_host.numel() * 8 + B * 28 * 4),
                "d2h_bytes_per_step": B * 8, "api": "sdag_b200.sdag.prefill_batch (pinned host ids -> first tokens on host)"},
        "gpu_launches": launches, "clocks": clocks,
    }
    if carg is not None:
        line["carg_tok_s"] = carg

    # free the model before the retrieval corpus is generated
    del model
    torch.cuda.empty_cache()

    # ---------------- dense top-k, config 3 ----------------
    if not args.no_secondary:
        N, d, Q, k = args.corpus_rows, 1024, args.queries, args.topk
        rows = N // world
        lo = rank * rows
        n_local = rows if rank < world - 1 else N - lo
        gen = torch.Generator(device=device).manual_seed(42 + rank)
        x = torch.empty((n_local, d), dtype=torch.float32, device=device)
        for r0 in range(0, n_local, 1 << 20):  # generate + normalise in 1M-row pieces
            r1 = min(n_local, r0 + (1 << 20))
            x[r0:r1].normal_(generator=gen)
            x[r0:r1] /= x[r0:r1].norm(dim=1, keepdim=True)
        index = ShardedFlatIPIndex(FlatIPIndex.from_device(x, id_offset=lo))
        qg = torch.Generator().manual_seed(7)
        xq_host = torch.randn(Q, d, generator=qg)
        xq_host = (xq_host / xq_host.norm(dim=1, keepdim=True)).pin_memory()
        xq_dev = xq_host.to(device)
        # enough steps for the clock sampler (100 ms period) to see the timed region: >= 0.6 s of searches
        search_index_probe = index.search_device(xq_dev, k)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        index.search_device(xq_dev, k)
        p1.record()
        torch.cuda.synchronize()
        del search_index_probe
        est_ms = max_over_ranks(max(p0.elapsed_time(p1), 0.05))
        s_steps = int(max(3, min(args.steps, 10), min(200, -(-600.0 // est_ms))))

        def search_resident():
            return index.search_device(xq_dev, k)

        def search_e2e():
            s, i = index.search_device(xq_host.to(device, non_blocking=True), k)
            return s.cpu(), i.cpu()

        ms_s, clocks_s, launches_s, _ = timed(search_resident, s_steps, args.warmup)
        ms_se, _, _, _ = timed(search_e2e, s_steps, args.warmup)
        qps = Q * s_steps / (ms_s / 1000.0)
        flops = 2.0 * Q * n_local * d  # per GPU per batch (algorithmic; re-scoring and merge excluded)
        ach = flops / (ms_s / s_steps / 1000.0) / 1e12
        stats = index.local.last_stats.cpu().tolist()

        # full-size parity: 8 sampled queries brute-forced in fp32 over this rank's whole shard (plain torch, chunked),
        # against the kernel path's local top-k — ids must match (exact ties aside), scores to 1e-5 relative
        def full_size_check(n_check=8):
            sel = torch.linspace(0, Q - 1, n_check, device=device).long()
            s_k, i_k = index.local.search_device(xq_dev[sel], k)
            best_s = torch.full((n_check, k), -float("inf"), device=device)
            best_i = torch.full((n_check, k), -1, dtype=torch.int64, device=device)
            for r0 in range(0, n_local, 1 << 20):
                r1 = min(n_local, r0 + (1 << 20))
                sc = xq_dev[sel] @ x[r0:r1].T
                cs, ci = torch.cat([best_s, sc], dim=1).topk(k, dim=1)
                ids = torch.cat([best_i, torch.arange(r0, r1, device=device).expand(n_check, -1) + lo], dim=1)
                best_s, best_i = cs, ids.gather(1, ci)
            same = (best_i == i_k)
            rel = ((best_s - s_k).abs() / best_s.abs().clamp(min=1e-12)).max()
            # a mismatching id is acceptable only where the two scores tie to fp32 summation-order noise
            tie_ok = ((best_s - s_k).abs() <= 2e-6 * best_s.abs()) | same
            return {"queries": n_check, "rows": n_local, "ids_equal_frac": float(same.float().mean()),
                    "all_ids_equal_or_tied": bool(tie_ok.all()), "max_rel_score_err": float(rel)}

        parity = full_size_check()
        ok_all = torch.tensor([1 if parity["all_ids_equal_or_tied"] and parity["max_rel_score_err"] < 1e-5 else 0], device=device)
        if world > 1:
            dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
        parity["all_ranks_ok"] = bool(int(ok_all))
        line["secondary"] = {
            "metric": SECONDARY_METRIC, "value": qps, "unit": "queries/s", "ms_per_step": ms_s / s_steps, "steps": s_steps,
            "scaling": "strong", "dtype": "bf16 tensor-core pass + fp32 exact re-score",
            "config": {"workload": f"configs[2]: {N} x {d} fp32 synthetic corpus (unit Gaussian rows) + bf16 shadow, "
                                   f"{Q}-query batch, k={k}, rows sharded over {world} GPU(s)",
                       "candidates_per_query": stats[1] / Q, "overflow_queries": stats[0]},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                         "kernel": "ip_gemm_filter_kernel_2cta (whole ip_search timed)",
                         "note": "peak is the cuBLAS bf16 figure measured on this pool; the CTA-pair kernel can exceed it (nominal dense bf16 2250)", "peak_source": pk["_source"] + ", bf16 sustained"},
            "e2e": {"value": Q * s_steps / (ms_se / 1000.0), "unit": "queries/s", "h2d_bytes_per_step": Q * d * 4,
                    "d2h_bytes_per_step": Q * k * 12, "api": "FlatIPIndex.search_device + host copies (faiss-shaped search)"},
            "gpu_launches": launches_s, "clocks": clocks_s, "parity_check": parity,
        }
        del index, x
        torch.cuda.empty_cache()

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_reference import flat_ip_cpu, sdag_prefill_cpu

        r = sdag_prefill_cpu(Lp, S, spans, qa, n_layers=1, vocab_div=8)
        line["cpu_baseline"] = {"value": r["tok_per_s"], "unit": "tok/s", "cores": r["threads"], "kind": "port", "sample": r["sample"]}
        if "secondary" in line:
            r2 = flat_ip_cpu(100_000, 1024, args.queries, args.topk, args.corpus_rows)
            line["secondary"]["cpu_baseline"] = {"value": r2["qps"], "unit": "queries/s", "cores": r2["threads"], "kind": "port",
                                                 "sample": r2["sample"]}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
