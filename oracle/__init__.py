"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's dense per-anchor path.

Nothing under ``oracle/`` is part of the product.  It may be imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` — always as the checker (or the timed CPU baseline), never as
a fallback for the CUDA path.  ``pytorch_retinanet_b200`` never imports it.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md §4), so the oracle is pinned against outputs of the unmodified
reference itself, imported in the build container through ``oracle/ref_shim.py``;
the resulting vectors are committed under ``tests/golden/`` together with the
script that made them (``tests/golden/make_golden.py``).
"""
