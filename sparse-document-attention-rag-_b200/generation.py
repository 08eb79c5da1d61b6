"""CARG (no document isolation) batched generation through the same kernels as the SDAG path.

Mirrors ``generate_batch_seq2seq`` (src/pipeline/attack/content_generation.py:44-94) and ``generate_noiso_batch``
(src/pipeline/main.py:308-378).  The reference left-pads the batch and lets ``model.generate`` mask the pad columns and
shift the position ids; here the prompts are packed right-padded instead: every prompt gets a causal block map of its
own length, its own RoPE positions (starting at 0, exactly what the shifted position ids give) and its own KV-cache
length in the decode kernel, so no pad column is ever computed on or masked.  Token-for-token the two are the same
computation.

Llama-family bf16 models run on ``llama_runner.FusedPrefill`` (ragged prefill + one CUDA-graph decode step per token
for the whole batch); other models fall back to one prompt at a time through the Hugging Face forward with the
``sdag_b200`` attention hook.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import _lib, config, prompts
from .ranked_list import apply_ranked_list_order, inject_malicious_docs_into_ranked_list


def _split_rows(enc) -> List[torch.Tensor]:
    """Token ids of every prompt without its padding, whichever side the tokenizer padded on."""
    ids, mask = enc["input_ids"], enc.get("attention_mask")
    if mask is None:
        return [row for row in ids]
    return [row[m.bool()] for row, m in zip(ids, mask)]


def _eos_ids(tokenizer, model) -> List[int]:
    """Every id ``model.generate`` stops on: ``generation_config.eos_token_id`` — an int or a list (Llama-3.1-Instruct
    ships [128001, 128008, 128009]); content_generation.py:79-87 passes no ``eos_token_id``, so HF takes the generation
    config's, and generates to ``max_new_tokens`` when that is unset.  Models without a generation config fall back to the
    tokenizer's id."""
    gc = getattr(model, "generation_config", None)
    src = getattr(gc, "eos_token_id", None) if gc is not None else getattr(tokenizer, "eos_token_id", None)
    ids: List[int] = []
    for v in (src if isinstance(src, (list, tuple)) else [src]):
        if v is not None and int(v) not in ids:
            ids.append(int(v))
    return ids


def _cut_at_eos(ids: List[int], eos) -> List[int]:
    """``model.generate`` finishes a sequence at its first EOS (the first new token included) and pads after it."""
    stops = set(eos if isinstance(eos, (list, tuple, set)) else ([] if eos is None else [eos]))
    for pos, t in enumerate(ids):
        if t in stops:
            return ids[: pos + 1]
    return ids


@torch.no_grad()
def _generate_one_via_hook(model, row: torch.Tensor, max_tokens: int, eos, temperature: float, top_p: Optional[float]):
    from . import sdag
    from .llama_runner import sample_tokens

    dev = row.device
    t = torch.tensor([temperature if temperature > 0 else 1.0], dtype=torch.float32, device=dev)
    p = torch.tensor([top_p], dtype=torch.float32, device=dev) if (top_p is not None and top_p < 1.0) else None

    def pick(lg):
        return sample_tokens(lg, t, p) if temperature > 0 else torch.argmax(lg, dim=-1, keepdim=True)

    logits, past = sdag.prefill_forward(model, row.unsqueeze(0), None)
    nxt = pick(logits)
    out = [int(nxt)]
    stops = set(eos if isinstance(eos, (list, tuple, set)) else ([] if eos is None else [eos]))
    while len(out) < max_tokens and out[-1] not in stops:
        step = model(input_ids=nxt, past_key_values=past, use_cache=True)
        past = step.past_key_values
        nxt = pick(step.logits[:, -1, :])
        out.append(int(nxt))
    return out


@torch.no_grad()
def generate_batch_seq2seq(tokenizer, model, prompts_list: Sequence[str], max_tokens: int = 128, temperature: float = 0.7,
                           top_p: float = 0.9) -> List[str]:
    """Same contract as the reference function: greedy when ``temperature == 0``, otherwise temperature + nucleus
    sampling; returns the decoded new tokens of every prompt (special tokens removed, stripped)."""
    from . import sdag
    from .llama_runner import get_runner

    prompts_list = list(prompts_list)
    if not prompts_list:
        return []
    device = next(model.parameters()).device
    if device.type != "cuda":
        raise _lib.SdagB200Error("sdag_b200 runs on CUDA devices only (no CPU fallback)")
    enc = tokenizer(prompts_list, return_tensors="pt", padding=True, truncation=True)
    rows = [r.to(device) for r in _split_rows(enc)]
    eos = _eos_ids(tokenizer, model)
    sdag.prepare_model(model)
    runner = get_runner(model) if sdag.USE_FUSED_PREFILL else None
    if runner is None:
        new_ids = [_generate_one_via_hook(model, r, max_tokens, eos, temperature, top_p) for r in rows]
    else:
        lengths = torch.tensor([int(r.numel()) for r in rows], dtype=torch.int64, device=device)
        L = int(lengths.max())
        pad = tokenizer.eos_token_id if tokenizer.eos_token_id is not None else 0
        ids = torch.full((len(rows), L), pad, dtype=torch.int64, device=device)
        for b, r in enumerate(rows):
            ids[b, : r.numel()] = r
        layouts = [sdag.SdagLayout(L=int(n), S=int(n), spans=[], qa_start=int(n)) for n in lengths.tolist()]
        bm = sdag.build_blockmap(layouts, device)
        bm.no_uniform_rows = True  # plain causal prompts
        # the prefill yields the first new token, the decode loop the remaining max_tokens - 1
        gen = runner.generate(ids, bm, max(max_tokens - 1, 0), eos, temperature=float(temperature), lengths=lengths,
                              top_p=top_p, eos_from_first=True)
        new_ids = [row[:max_tokens] for row in gen.tolist()]
    return [tokenizer.decode(_cut_at_eos(ids_b, eos), skip_special_tokens=True).strip() for ids_b in new_ids]


def generate_noiso_batch(tokenizer, llm_model, queries: List[str], defended_docs_batch: List[List[str]],
                         malicious_docs_survived_batch: List[List[str]], attacker_pos: int, order_mode: str) -> List[str]:
    """RAG answers without document isolation (main.py:308-378): inject the surviving adversarial documents at the
    attack rank (``config.ORACLE``), apply the ranked-list order, build the chat prompt, generate in sub-batches of
    ``config.LLM_BATCH_SIZE`` with ``config.MAX_GEN_TOKENS_RAG`` / ``TEMPERATURE`` / ``TOP_P``."""
    rag_prompts: List[str] = []
    for q, docs_ranked, mals in zip(queries, defended_docs_batch, malicious_docs_survived_batch):
        if config.ORACLE:
            ranked = inject_malicious_docs_into_ranked_list(base_docs=list(docs_ranked), malicious_docs=list(mals),
                                                            attack_pos=attacker_pos)
        else:
            ranked = list(docs_ranked)
        ranked = apply_ranked_list_order(ranked, order_mode)
        docs_text = "\n\n".join(f"- {d.strip()}" for d in ranked if d and d.strip())
        chat = tokenizer.apply_chat_template(
            [{"role": "system", "content": prompts.SYSTEM_PROMPT_RAG},
             {"role": "user", "content": prompts.USER_RAG_PROMPT.format(query=q, docs_text=docs_text)}],
            tokenize=False, add_generation_prompt=True)
        rag_prompts.append(chat)
    answers: List[str] = []
    step = max(1, int(config.LLM_BATCH_SIZE))
    for j in range(0, len(rag_prompts), step):
        answers.extend(generate_batch_seq2seq(tokenizer, llm_model, rag_prompts[j : j + step],
                                              max_tokens=config.MAX_GEN_TOKENS_RAG, temperature=config.TEMPERATURE,
                                              top_p=config.TOP_P))
    return answers
