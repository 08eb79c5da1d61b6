"""CPU oracle for the two hot paths (TEST INFRASTRUCTURE — not product code).

Everything under ``oracle/`` is a restatement, in numpy / CPU torch, of what the reference computes on
the paths SURVEY.md §8 names.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker or the timed CPU
baseline.  The product package never imports this module: it calls the CUDA library and fails loudly
when the library is missing.

Pinning status:

* path A (SDAG mask / masked attention / generation loop): pinned.  ``tests/golden/make_golden.py``
  imports the reference's own ``SDAG.py`` from ``/root/reference`` (with a stub ``sentence_transformers``
  module) and stores its outputs on seeded inputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
  checks this oracle against those files bit-for-bit (masks, kNN lists, greedy token ids) and to 1e-5
  (attention outputs, fp32).
* path B selector: pinned the same way against ``attack/malicious_selection.py``.
* path B search: the arithmetic lives in ``faiss-cpu`` (un-vendored, unpinned — README.md:78), which is not
  installed and cannot be; ``oracle/flat_ip.py`` restates ``IndexFlatIP.search`` from its published
  semantics and is pinned only at the reference's own call sites (``dense.py:68-130`` run unmodified on top
  of it).  **The faiss arithmetic itself is "parity unpinned".**
"""
