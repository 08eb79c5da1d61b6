"""sdag_b200 — B200-native (sm_100a) hot paths of Sparse-Document-Attention-RAG.

Import as ``sdag_b200`` (the repo-root ``sdag_b200.py`` aliases this directory, whose name is not a Python
identifier).  Sub-modules mirror the reference's modules on the two hot paths:

* ``sdag``       — ``src/pipeline/sparse_attention_RAG/SDAG.py``
* ``retrieval``  — ``src/pipeline/retrieval/dense.py`` + ``retriever.py``
* ``selection``  — ``src/pipeline/attack/malicious_selection.py``
* ``generation`` — ``generate_batch_seq2seq`` (attack/content_generation.py) / ``generate_noiso_batch`` (main.py): CARG
* ``hybrid``     — ``src/pipeline/retrieval/hybrid.py`` (RRF fusion, host only)
* ``harness``    — the query loop of ``src/pipeline/main.py`` with its answer scoring, per-pair CSV and JSON metrics
                   (``utils/normalization.py``, ``parsing.py``, ``save_results.py``, ``metrics.py``; host only)
* ``ranked_list``, ``prompts``, ``datamodels``, ``config`` — the small host-side pieces those call into

All arithmetic runs in ``libsdag_b200.so`` (hand-written CUDA, C ABI in ``include/sdag_b200.h``).
"""
__version__ = "0.1.0"
