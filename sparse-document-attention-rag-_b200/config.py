"""Knob names of the reference's ``src/pipeline/config.py`` that touch the two hot paths.

Only the names are API surface; values are the reference defaults (config.py:20-87).  Unlike the reference
(which copies ``TEMPERATURE`` etc. at import time, SURVEY.md §5 "Hazard"), the hot-path modules here read
``sdag.TEMPERATURE`` / ``sdag.MAX_GEN_TOKENS_RAG`` module attributes that callers may patch the same way.
"""
SEED = 42
TOP_K = [5]
ADD_ATTACK_IN_RANK = [1]
BATCH_SIZE_EMBED_Q = 32
LLM_BATCH_SIZE = 4
RANKER_MODEL_NAME = "intfloat/e5-large-v2"
MAX_GEN_TOKENS_RAG = 500
TEMPERATURE = 0.1
TOP_P = 1.0
DOC_NEIGHBORS_K = 0
MALICIOUS_DOC_SELECTION_STRATEGY = "random"
MAX_MALICIOUS_DOCS_PER_QUERY = 1
ORACLE = True
RETRIEVER_BACKEND = "dense"
FAISS_INDEX_PATH = "faiss.index"
META_JSONL_PATH = "docs_meta.jsonl"
