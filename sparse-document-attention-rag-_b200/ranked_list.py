"""Where the adversarial documents land in the ranked list (host logic, path A1b).

Same contract as ``inject_malicious_docs_into_ranked_list`` (src/pipeline/utils/ranked_list.py:31-110):
an int position is a 1-indexed rank (clamped to the list) at which the documents are inserted consecutively,
-1 draws ``random.randint(0, len)`` per document from the global RNG, 0 or any other negative leaves the
list alone; a sequence gives one position per document (missing entries mean -1, extra ones are dropped),
positive ones are placed first from the highest rank down, then the -1 ones; ``None`` / other values place
nothing.
"""
from __future__ import annotations

import random
from typing import List, Sequence, Union


def _rank_to_index(rank: int, n: int) -> int:
    return min(max(rank - 1, 0), n)


def inject_malicious_docs_into_ranked_list(
    base_docs: List[str], malicious_docs: List[str], attack_pos: Union[int, Sequence[int]]
) -> List[str]:
    out = list(base_docs)
    if not malicious_docs:
        return out
    if isinstance(attack_pos, int):
        if attack_pos > 0:
            at = _rank_to_index(attack_pos, len(out))
            out[at:at] = list(malicious_docs)
        elif attack_pos == -1:
            for doc in malicious_docs:
                out.insert(random.randint(0, len(out)), doc)
        return out

    positions = list(attack_pos)[: len(malicious_docs)]
    positions += [-1] * (len(malicious_docs) - len(positions))
    ranked = [(p, doc) for doc, p in zip(malicious_docs, positions) if p is not None and p > 0]
    # stable descending sort keeps the reference's order among equal ranks
    for p, doc in sorted(ranked, key=lambda pd: pd[0], reverse=True):
        out.insert(_rank_to_index(p, len(out)), doc)
    for doc, p in zip(malicious_docs, positions):
        if p == -1:
            out.insert(random.randint(0, len(out)), doc)
    return out


def apply_ranked_list_order(ranked_docs: List[str], order_mode: str) -> List[str]:
    """Presentation order of the ranked list in the prompt (src/pipeline/utils/ranked_list.py:113-139):
    ``"bottom_up"`` reverses, ``"random"`` shuffles a copy with the global ``random`` state, anything else
    (``"top_down"``) returns the list object itself."""
    if order_mode == "bottom_up":
        return ranked_docs[::-1]
    if order_mode == "random":
        shuffled = list(ranked_docs)
        random.shuffle(shuffled)
        return shuffled
    return ranked_docs
