"""Prompt text of the RAG generation path.

The SDAG span arithmetic depends on this exact layout (documents as ``- <doc>`` joined by blank lines, then
a ``- Question:`` marker), so the strings must equal the reference's ``SYSTEM_PROMPT_RAG`` /
``USER_RAG_PROMPT`` (src/pipeline/utils/prompts.py:15, 32-41) character for character; the golden
``chat_sha256`` values in tests/golden/generation.json check that.
"""
SYSTEM_PROMPT_RAG = "You are a helpful assistant, below is a query from a user and some relevant contexts."

USER_RAG_PROMPT = (
    "Answer the question concisely, based on the following passages.\n"
    "Keep the answer concise.\n"
    "\n"
    "passages:\n"
    "{docs_text}\n"
    "\n"
    "- Question: {query}\n"
    "\n"
    "- Answer:\n"
)

QUESTION_MARKER = "- Question:"
