"""The BASELINE.json configurations beside the headline one, as functions bench.py calls (and that can be run alone):

* ``same_gpu_hf``  — the fair "before" for SDAG prefill tok/s: Hugging Face Llama on the same GPU, fed the reference's
  dense additive ``finfo.min`` mask exactly as ``generate_with_custom_mask`` does (SDAG.py:161-176), ``sdpa`` and ``eager``.
* ``config5``      — configs[4]: long-context SDAG, TOP_K = 50 x 640-token documents, DOC_NEIGHBORS_K = 2, L = 32 122,
  prefill + 128 decode steps on one GPU (SDAG.py:161-211).
* ``config4``      — configs[3]: retrieval + near / far centroid selection + SDAG generation (+ the CARG answers the
  reference's main loop also produces), TOP_K = 10, queries of the NQ attack set split round-robin over the GPUs
  (main.py:764-796, attack/malicious_selection.py:140-180) through ``sdag_b200.harness.main(query_slice=...)``.
* ``config1``      — configs[0]: 50 sampled NQ queries, 100 k-document corpus encoded on the GPU, TOP_K = 5, SDAG strict,
  Qwen2.5-0.5B geometry.

Everything is synthetic where the reference downloads (weights random-init, corpus = template documents, byte-level
tokenizer); the queries and the adversarial candidates are the reference's own NQ set (tests/golden/nq_attack_1000.json.gz).
"""
from __future__ import annotations

import contextlib
import gzip
import json
import os
import random
import sys
import tempfile
import time
from typing import Dict, List, Optional

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

NQ_FIXTURE = os.path.join(ROOT, "tests", "golden", "nq_attack_1000.json.gz")


def synthetic_layout(sys_user_len, n_docs, doc_len, gap, qa_len):
    """Fixed-shape prompt layout of SURVEY.md §8d: (L, S, doc spans, qa_start)."""
    spans, pos = [], sys_user_len
    for _ in range(n_docs):
        spans.append((pos, pos + doc_len))
        pos += doc_len + gap
    return pos + qa_len, sys_user_len, spans, pos


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _timed_ms(fn, reps: int, warmup: int = 1) -> List[float]:
    for _ in range(warmup):
        fn()
    out = []
    for _ in range(reps):
        a, b = _events()
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return out


# ------------------------------------------------------------------------------------------------------
# same-GPU Hugging Face baseline
# ------------------------------------------------------------------------------------------------------
def same_gpu_hf(model, ids_dev: torch.Tensor, layout, prompts: int = 4) -> Dict:
    """HF forward of one prompt at a time (the reference is batch-of-1) with the dense additive mask: tok/s for
    ``attn_implementation`` sdpa and eager.  The bool mask comes from this library's ``build_blocked_causal_mask_full``
    (bit-identical to the reference builder — tests/golden/masks.json) and is converted like SDAG.py:164-173; it is built
    once, outside the timed region (the reference rebuilds it per query with one slice write per row, which this number
    does not charge to it).  Logits are computed for all L rows, as the reference does (SDAG.py:176-181)."""
    import sdag_b200.sdag as sdag

    Lp, S, spans, qa = layout
    dev = ids_dev.device
    dtype = next(model.parameters()).dtype
    mask = sdag.build_blocked_causal_mask_full(Lp, 0, S, list(spans), qa, str(dev))
    add = torch.zeros((Lp, Lp), dtype=dtype, device=dev).masked_fill(~mask, torch.finfo(dtype).min)[None, None]
    n = max(1, min(prompts, ids_dev.shape[0]))
    res = {"layers": len(model.model.layers), "prompts_timed": n, "tokens_per_prompt": Lp, "batch": 1,
           "mask": "dense additive finfo.min [1,1,L,L] (SDAG.py:164-173), prebuilt", "logits_rows": "all L (SDAG.py:176)"}
    prev = getattr(model.config, "_attn_implementation", None)
    try:
        for impl in ("sdpa", "eager"):
            model.set_attn_implementation(impl)

            def fwd(b):
                with torch.no_grad():
                    out = model(input_ids=ids_dev[b : b + 1], attention_mask=add, use_cache=True)
                    return out.logits[:, -1, :].argmax(-1)

            fwd(0)
            fwd(0)
            ms = []
            for b in range(n):
                ms += _timed_ms(lambda: fwd(b), 1, warmup=0)
            res[f"{impl}_ms_per_prompt"] = sum(ms) / len(ms)
            res[f"{impl}_tok_s"] = Lp / (sum(ms) / len(ms) / 1000.0)
    finally:
        sdag.prepare_model(model)
        if prev is not None and getattr(model.config, "_attn_implementation", None) != prev:
            model.set_attn_implementation(prev)
        torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------------
# config 5: long context, neighbours, prefill + decode
# ------------------------------------------------------------------------------------------------------
def config5(model, device, new_tokens: int = 128, peak_tflops: float = 1360.8, peak_gbs: float = 6550.0) -> Dict:
    import sdag_b200.sdag as sdag
    from sdag_b200.llama_runner import get_runner

    Lp, S, spans, qa = synthetic_layout(48, 50, 640, 1, 24)
    g = torch.Generator().manual_seed(42)
    emb = torch.randn(len(spans), 64, generator=g)
    emb = (emb / emb.norm(dim=1, keepdim=True)).to(device)
    nbrs = sdag.doc_knn_device(emb, 2).cpu().tolist()  # directed kNN lists, forward (non-causal) neighbours included
    lay = sdag.SdagLayout(Lp, S, spans, qa, nbrs)
    bm = sdag.build_blockmap([lay], device)
    hdr = bm.header(0)
    assert hdr["status"] == 0 and hdr["pairs"] == 53523828, hdr  # P of SURVEY.md §8d (K = 2, all docs 640 tokens)
    ids = torch.randint(0, 128000, (1, Lp), generator=torch.Generator().manual_seed(5)).to(device)
    runner = get_runner(model)
    assert runner is not None
    cfg = model.config
    H, Hkv = cfg.num_attention_heads, cfg.num_key_value_heads
    D = getattr(cfg, "head_dim", None) or cfg.hidden_size // H
    n_layers = len(model.model.layers)

    runner.generate(ids, bm, new_tokens, None)  # builds the decode context of this capacity bucket (cache, captured graph)
    prefill_ms = min(_timed_ms(lambda: runner.generate(ids, bm, 0, None), 3, warmup=1))
    sdag._profile_events = []
    runner.generate(ids, bm, 0, None)
    torch.cuda.synchronize()
    evs, sdag._profile_events = sdag._profile_events, None
    attn_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    gen_ms = min(_timed_ms(lambda: runner.generate(ids, bm, new_tokens, None), 3, warmup=0))
    decode_ms = (gen_ms - prefill_ms) / new_tokens
    attn_flops = 4.0 * D * hdr["pairs"] * H
    ctx = runner._last_ctx  # the context generate() just used
    kv_bytes = 2.0 * (Lp + new_tokens / 2) * Hkv * D * 2 * n_layers
    w_bytes = (sum(p.numel() for p in model.parameters()) - model.model.embed_tokens.weight.numel()) * 2.0
    return {
        "workload": f"configs[4]: Llama-3-8B geometry, TOP_K=50 x 640-token docs, DOC_NEIGHBORS_K=2 (kNN over 50 random unit vectors, "
                    f"seed 42), S=48, 1 gap token per doc, 24 QA tokens, L={Lp}; prefill + {new_tokens} greedy decode steps, batch of 1",
        "L": Lp, "layers": n_layers, "allowed_pairs": hdr["pairs"], "new_tokens": new_tokens,
        "prefill_ms": prefill_ms, "prefill_tok_s": Lp / (prefill_ms / 1000.0),
        "decode_ms_per_token": decode_ms, "decode_graph": ctx["graph"] is not None,
        "decode_hbm": {"bytes_per_token": kv_bytes + w_bytes, "achieved_gbs": (kv_bytes + w_bytes) / decode_ms / 1e6,
                       "peak_gbs": peak_gbs, "frac": (kv_bytes + w_bytes) / decode_ms / 1e6 / peak_gbs},
        "attention_roofline": {"kernel": "sdag_prefill_kernel<128> (+ v_colsum, v_mean) per layer launch", "bound": "tensor",
                               "launch_ms": attn_ms, "flops_per_launch": attn_flops,
                               "achieved": attn_flops / (attn_ms / 1000.0) / 1e12, "peak": peak_tflops, "unit": "TFLOP/s",
                               "frac": attn_flops / (attn_ms / 1000.0) / 1e12 / peak_tflops},
    }


# ------------------------------------------------------------------------------------------------------
# configs 4 and 1: the chained pipeline
# ------------------------------------------------------------------------------------------------------
WORDS = ("alpha beta gamma delta river mountain ocean tower bridge city museum garden engine signal orbit planet crystal "
         "forest harbor market library station valley desert island canyon glacier meadow village castle").split()


def make_vocabulary(n: int = 5000, seed: int = 42) -> List[str]:
    """5 k pronounceable pseudo-words (template sentences over a 5 k-word vocabulary)."""
    rnd = random.Random(seed)
    syl = [c + v for c in "bcdfghjklmnprstvwz" for v in "aeiou"]
    words = set(WORDS)
    while len(words) < n:
        words.add("".join(rnd.choice(syl) for _ in range(rnd.randint(2, 4))))
    return sorted(words)


def make_corpus(n_docs: int, seed: int = 42, lo: int = 64, hi: int = 256) -> List[str]:
    """Template sentences, 64-256 bytes (= tokens of the byte-level tokenizer) per document."""
    vocab = make_vocabulary()
    rnd = random.Random(seed)
    templates = ("the {} of {} is near the {}", "{} was built by the {} of {}", "in {} the {} became a {}",
                 "a {} can hold {} and {}", "{} and {} share the {}")
    docs = []
    for _ in range(n_docs):
        target = rnd.randint(lo, hi)
        parts: List[str] = []
        while sum(len(p) + 2 for p in parts) < target:
            t = rnd.choice(templates)
            parts.append(t.format(*(rnd.choice(vocab) for _ in range(t.count("{}")))))
        docs.append(". ".join(parts)[:target])
    return docs


def load_nq_fixture() -> List[Dict]:
    with gzip.open(NQ_FIXTURE, "rt", encoding="utf-8") as f:
        return json.load(f)["queries"]


class StageTimer:
    """Wall time (with a device synchronise on both sides) of the named stages of the harness loop."""

    def __init__(self):
        self.seconds: Dict[str, float] = {}

    def wrap(self, name, fn):
        def inner(*a, **kw):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            try:
                return fn(*a, **kw)
            finally:
                torch.cuda.synchronize()
                self.seconds[name] = self.seconds.get(name, 0.0) + time.perf_counter() - t0
        return inner


@contextlib.contextmanager
def _patched(obj, **attrs):
    old = {k: getattr(obj, k) for k in attrs}
    for k, v in attrs.items():
        setattr(obj, k, v)
    try:
        yield
    finally:
        for k, v in old.items():
            setattr(obj, k, v)


class Pipeline:
    """Encoder + replicated corpus index + tokenizer, built once per process and shared by configs 1 and 4."""

    def __init__(self, device, rank: int, world: int, n_docs: int = 100_000, enc_layers: int = 24):
        import synth
        from transformers import BertConfig, BertModel

        from sdag_b200.encoder import B200Ranker, encode_corpus
        from sdag_b200.retrieval import DenseRetriever, FlatIPIndex

        self.device, self.rank, self.world = device, rank, world
        self.tok = synth.make_tokenizer()
        self.docs = make_corpus(n_docs)
        self.meta = [{"id": f"d{i}", "text": d} for i, d in enumerate(self.docs)]
        torch.manual_seed(42)
        with torch.device(device):
            bert = BertModel(BertConfig(vocab_size=30522, hidden_size=1024, num_hidden_layers=enc_layers, num_attention_heads=16,
                                        intermediate_size=4096, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
                             add_pooling_layer=False).to(torch.bfloat16)
        self.ranker = B200Ranker(bert, self.tok, device=device)
        # "encode corpus": every rank encodes a contiguous block of documents, one all-gather replicates the index
        lo, hi = rank * n_docs // world, (rank + 1) * n_docs // world
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        local = encode_corpus(self.ranker, self.docs[lo:hi])
        torch.cuda.synchronize()
        self.encode_seconds = time.perf_counter() - t0
        self.encode_docs_per_s_per_gpu = (hi - lo) / self.encode_seconds
        if world > 1:
            import torch.distributed as dist

            counts = [(r + 1) * n_docs // world - r * n_docs // world for r in range(world)]
            f32 = [torch.empty((c, 1024), dtype=torch.float32, device=device) for c in counts]
            b16 = [torch.empty((c, 1024), dtype=torch.bfloat16, device=device) for c in counts]
            dist.all_gather(f32, local._f32[: local.ntotal].contiguous())
            dist.all_gather(b16, local._bf16[: local.ntotal].contiguous())
            self.index = FlatIPIndex.from_encoded(torch.cat(f32), torch.cat(b16))
        else:
            self.index = local
        self.retriever = DenseRetriever(self.ranker, self.index, self.meta)

    def run(self, llm, queries: List[Dict], my_slice: List[int], top_k: int, strategy: str, new_tokens: int) -> Dict:
        """One pass of ``harness.main`` over this rank's queries; returns stage seconds and the pair metrics."""
        import sdag_b200.generation as generation
        import sdag_b200.harness as harness
        import sdag_b200.sdag as sdag
        import sdag_b200.selection as selection
        from sdag_b200 import config

        timer = StageTimer()
        retr = self.retriever

        def select_with_index_reuse(ranker, retrieved_docs_batch_full, malicious_doc_groups_batch, strategy, max_docs, **kw):
            return orig_select(ranker, retrieved_docs_batch_full, malicious_doc_groups_batch, strategy, max_docs,
                               device=self.device, **kw)

        orig_select = selection.select_malicious_docs_for_batch
        out_dir = tempfile.mkdtemp(prefix="sdag_b200_cfg_")
        with _patched(config, TOP_K=[top_k], ADD_ATTACK_IN_RANK=[1], DOC_NEIGHBORS_K=0, MALICIOUS_DOC_SELECTION_STRATEGY=strategy,
                      MAX_MALICIOUS_DOCS_PER_QUERY=1, ORACLE=True, MAX_GEN_TOKENS_RAG=new_tokens, TEMPERATURE=0.0), \
                _patched(sdag, TEMPERATURE=0.0, MAX_GEN_TOKENS_RAG=new_tokens), \
                _patched(retr, retrieve_batch=timer.wrap("retrieve", retr.retrieve_batch)), \
                _patched(selection, select_malicious_docs_for_batch=timer.wrap("select", select_with_index_reuse)), \
                _patched(sdag, run_rag_with_doc_isolation_batch=timer.wrap("generate_sdag", sdag.run_rag_with_doc_isolation_batch)), \
                _patched(generation, generate_noiso_batch=timer.wrap("generate_carg", generation.generate_noiso_batch)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(sys.stderr):
                metrics = harness.main(
                    retriever=retr, llm_model=llm, tokenizer=self.tok, ranker=self.ranker,
                    queries=[q["question"] for q in queries], short_answers=[q["short_answers"] for q in queries],
                    query_ids=[q["qid"] for q in queries], false_answer_groups=[q["false_answers"] for q in queries],
                    malicious_doc_groups=[q["malicious_docs"] for q in queries],
                    output_base=os.path.join(out_dir, f"rank{self.rank}_{strategy}"), device=str(self.device), query_slice=my_slice)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
        key = (top_k, 1)
        m = metrics.get(key, {})
        return {"seconds": wall, "stage_seconds": {k: round(v, 3) for k, v in timer.seconds.items()},
                "rows": int(m.get("num_queries", len(my_slice))) if isinstance(m, dict) else len(my_slice)}


def build_qwen05(device):
    """Qwen2.5-0.5B geometry (24 layers, hidden 896, 14 q / 2 kv heads of 64, FFN 4864, vocab 151 936, tied), bf16, seed 42."""
    from transformers import Qwen2Config, Qwen2ForCausalLM

    cfg = Qwen2Config(vocab_size=151936, hidden_size=896, intermediate_size=4864, num_hidden_layers=24, num_attention_heads=14,
                      num_key_value_heads=2, max_position_embeddings=32768, rope_theta=1000000.0, tie_word_embeddings=True)
    torch.manual_seed(42)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = Qwen2ForCausalLM(cfg)
    finally:
        torch.set_default_dtype(prev)
    return model.eval()


def _max_over_ranks(x: float, device, world: int) -> float:
    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def config4(pipe: Pipeline, llm, n_queries: int, new_tokens: int = 32) -> Dict:
    """configs[3]: both selection strategies over the first ``n_queries`` NQ attack queries, round-robin over the ranks."""
    queries = load_nq_fixture()
    n = min(n_queries, len(queries))
    mine = list(range(pipe.rank, n, pipe.world))
    per = {}
    total = 0.0
    # untimed warm-up on the first queries: decode contexts (KV cache + captured graph per batch shape) are created, kernels loaded
    pipe.run(llm, queries, mine[: min(8, len(mine))], top_k=10, strategy="closest_to_centroid", new_tokens=new_tokens)
    for strategy in ("closest_to_centroid", "furthest_from_centroid"):
        r = pipe.run(llm, queries, mine, top_k=10, strategy=strategy, new_tokens=new_tokens)
        r["seconds"] = _max_over_ranks(r["seconds"], pipe.device, pipe.world)
        total += r["seconds"]
        per[strategy] = r
    return {
        "workload": f"configs[3]: dense top-10 over a 100k-doc synthetic corpus -> centroid selection (near, then far; 1 of 5 CSV "
                    f"candidates) -> SDAG generation + the CARG answers of the reference loop, first {n} of the 1000 NQ attack "
                    f"queries, round-robin over {pipe.world} GPU(s), weights and index replicated, no collective on the query path",
        "queries": n, "queries_total_in_set": len(queries), "new_tokens": new_tokens, "generator": "Llama-3-8B geometry bf16 random-init, greedy",
        "max_gen_tokens_note": f"MAX_GEN_TOKENS_RAG={new_tokens} instead of 500: random-init weights never emit EOS",
        "warmup": "first 8 queries of the rank's slice once, untimed (decode graphs captured)",
        "seconds": total, "queries_per_s": 2 * n / total, "queries_per_s_unit": "(query, strategy) results per second, each with an SDAG and a CARG answer",
        "per_strategy": per, "encode_corpus_seconds": pipe.encode_seconds,
        "encode_docs_per_s_per_gpu": pipe.encode_docs_per_s_per_gpu,
    }


def config1(pipe: Pipeline, device, new_tokens: int = 16) -> Dict:
    """configs[0]: 50 sampled NQ queries (first 50 of random.Random(42).sample), TOP_K = 5, SDAG strict, Qwen2.5-0.5B geometry."""
    import sdag_b200.sdag as sdag

    queries = load_nq_fixture()
    picked = random.Random(42).sample(range(len(queries)), len(queries))[:50]
    mine = picked[pipe.rank :: pipe.world]
    llm = sdag.prepare_model(build_qwen05(device))
    pipe.run(llm, queries, mine[: min(8, len(mine))], top_k=5, strategy="closest_to_centroid", new_tokens=new_tokens)  # untimed warm-up
    r = pipe.run(llm, queries, mine, top_k=5, strategy="closest_to_centroid", new_tokens=new_tokens)
    secs = _max_over_ranks(r["seconds"], pipe.device, pipe.world)
    del llm
    torch.cuda.empty_cache()
    return {
        "workload": "configs[0]: 50 sampled NQ queries, dense e5-large-v2 geometry (random-init) over 100k synthetic docs, TOP_K=5, "
                    "SDAG strict (DOC_NEIGHBORS_K=0), Qwen2.5-0.5B geometry random-init bf16, greedy",
        "queries": 50, "new_tokens": new_tokens, "warmup": "first 8 queries once, untimed (decode graphs captured)", "seconds": secs, "queries_per_s": 50 / secs, "stage_seconds": r["stage_seconds"],
        "encode_corpus_seconds": pipe.encode_seconds, "encode_docs_per_s_per_gpu": pipe.encode_docs_per_s_per_gpu,
    }
