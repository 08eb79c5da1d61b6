"""CPU baselines for bench.py (TEST INFRASTRUCTURE — only ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs call this; never the product path).

The reference itself is pure Python on top of Hugging Face / faiss and cannot travel to the GPU box
(``/root/reference`` is absent there), so these are ports ("kind": "port") of what the reference executes:

* SDAG prefill — ``generate_with_custom_mask`` (SDAG.py:161-181): build the dense bool mask row by row
  (SDAG.py:104-125 issues one slice write per row), turn it into an additive ``finfo.min`` bias, run the HF
  Llama forward with that bias over the whole prompt (SDPA with a float mask), compute logits for all L
  positions.  A bounded sample runs ``n_layers`` decoder layers of the full Llama-3-8B geometry and a
  ``1/vocab_div`` slice of the LM head; the per-prompt time is extrapolated linearly (layers are identical,
  the head is linear in vocab) and the extrapolation is spelled out in the returned ``sample`` string.
* dense top-k — ``search_index`` (dense.py:68-84) over ``IndexFlatIP``: ``oracle.flat_ip.FlatIPOracle`` on a
  row slice of the corpus; flat search is linear in N.
"""
from __future__ import annotations

import os
import time
from typing import Dict, Sequence, Tuple

import numpy as np
import torch


def reference_mask_builder():
    """(builder, provenance): the reference's own ``build_blocked_causal_mask_full`` (SDAG.py:68-127), imported unmodified
    when ``/root/reference`` is present (this container; ``sentence_transformers`` is stubbed — only its name is needed
    at import, SDAG.py:7); on the GPU box, where the reference cannot travel, the restatement below, which is held to the
    reference's masks bit for bit by tests/test_oracle_golden.py (``test_cpu_reference_mask_equals_reference_builder``)."""
    import sys
    import types

    root = "/root/reference"
    if os.path.isdir(os.path.join(root, "src", "pipeline")):
        try:
            if "sentence_transformers" not in sys.modules:
                stub = types.ModuleType("sentence_transformers")
                stub.SentenceTransformer = object
                sys.modules["sentence_transformers"] = stub
            if root not in sys.path:
                sys.path.insert(0, root)
            from src.pipeline.sparse_attention_RAG.SDAG import build_blocked_causal_mask_full as ref_build

            def build(L, S, spans, qa_start):
                return ref_build(L, 0, S, list(spans), qa_start, "cpu", None)

            return build, "reference function (SDAG.py:68-127), imported"
        except Exception:  # any import problem of the reference tree: use the restatement
            pass
    return reference_style_mask, "restatement of SDAG.py:104-125 (row-by-row slice writes), pinned to the reference's masks"


def reference_style_mask(L: int, S: int, spans: Sequence[Tuple[int, int]], qa_start: int) -> torch.Tensor:
    """The reference's algorithmic shape: O(L) separate row-slice writes into a dense bool [L, L]."""
    m = torch.zeros(L, L, dtype=torch.bool)
    owner = {}
    for (a, b) in spans:
        for i in range(a, b):
            owner[i] = a
    for i in range(L):
        row = m[i]
        if i < S or i >= qa_start:
            row[: i + 1] = True
        elif i in owner:
            row[:S] = True
            row[owner[i] : i + 1] = True
    return m


class SdagPrefillCpu:
    """Bounded sample of the reference SDAG prefill on the host: ``step()`` times one pass, ``result()`` extrapolates."""

    def __init__(self, L: int, S: int, spans, qa_start: int, n_layers: int = 1, vocab_div: int = 8, threads: int = 0,
                 full_layers: int = 32, seed: int = 42):
        from transformers import LlamaConfig
        from transformers.models.llama.modeling_llama import LlamaModel

        self.L, self.S, self.spans, self.qa_start = L, S, spans, qa_start
        self.n_layers, self.vocab_div, self.full_layers = n_layers, vocab_div, full_layers
        self.build_mask, self.mask_source = reference_mask_builder()
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        vocab = 128256
        cfg = LlamaConfig(vocab_size=1024, hidden_size=4096, intermediate_size=14336, num_hidden_layers=n_layers,
                          num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192, rope_theta=500000.0)
        cfg._attn_implementation = "sdpa"
        torch.manual_seed(seed)
        self.model = LlamaModel(cfg).eval()
        self.head = torch.randn(vocab // vocab_div, 4096) * 0.02
        self.ids = torch.randint(0, 1024, (1, L), generator=torch.Generator().manual_seed(seed))

    def step(self) -> Dict:
        """One sample pass; returns its parts in seconds and the extrapolated per-prompt time."""
        t0 = time.perf_counter()
        mask = self.build_mask(self.L, self.S, self.spans, self.qa_start)
        neg = torch.finfo(torch.float32).min
        bias = torch.zeros(self.L, self.L, dtype=torch.float32).masked_fill(~mask, neg)[None, None]
        t_mask = time.perf_counter() - t0
        with torch.no_grad():
            t0 = time.perf_counter()
            hidden = self.model(input_ids=self.ids, attention_mask=bias, use_cache=True).last_hidden_state
            t_layers = time.perf_counter() - t0
            t0 = time.perf_counter()
            _ = hidden[0] @ self.head.T
            t_head = time.perf_counter() - t0
        per_prompt = t_mask + t_layers * (self.full_layers / self.n_layers) + t_head * self.vocab_div
        return dict(t_mask=t_mask, t_layers=t_layers, t_head=t_head, seconds_sample=t_mask + t_layers + t_head,
                    seconds_per_prompt=per_prompt)

    def describe(self, r: Dict) -> str:
        return (f"1 prompt, L={self.L}: row-by-row mask build {r['t_mask']:.2f}s + {self.n_layers}/{self.full_layers} "
                f"Llama-3-8B-geometry fp32 layers (SDPA + additive finfo.min mask) {r['t_layers']:.2f}s + 1/{self.vocab_div} of "
                f"the L x vocab head {r['t_head']:.2f}s on {self.threads} threads; per-prompt time extrapolated linearly to "
                f"{self.full_layers} layers and the full head = {r['seconds_per_prompt']:.1f}s")


def sdag_prefill_cpu(L: int, S: int, spans, qa_start: int, **kw) -> Dict:
    job = SdagPrefillCpu(L, S, spans, qa_start, **kw)
    r = job.step()
    r.update(tok_per_s=L / r["seconds_per_prompt"], threads=job.threads, sample=job.describe(r), mask_builder=job.mask_source)
    return r


def flat_ip_cpu(n_rows: int, d: int, Q: int, k: int, full_rows: int, threads: int = 0, seed: int = 42) -> Dict:
    """Time the restated IndexFlatIP search on a row slice; returns seconds and QPS extrapolated to full_rows."""
    from .flat_ip import FlatIPOracle

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n_rows, d, generator=g)
    x /= x.norm(dim=1, keepdim=True)
    q = torch.randn(Q, d, generator=g)
    q /= q.norm(dim=1, keepdim=True)
    idx = FlatIPOracle(d)
    idx.add(x.numpy())
    t0 = time.perf_counter()
    idx.search(q.numpy(), k)
    t = time.perf_counter() - t0
    full = t * (full_rows / n_rows)
    return dict(seconds_sample=t, seconds_full=full, qps=Q / full, threads=threads,
                sample=(f"{Q} queries x {n_rows} of {full_rows} corpus rows (d={d}, k={k}) in {t:.2f}s on {threads} threads; "
                        f"flat search is linear in N -> {full:.1f}s per batch"))
