"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as
the checker or the CPU baseline; the product path (``spectralclustersupertree_b200``) never
does and fails loudly when its CUDA library is missing.
"""
