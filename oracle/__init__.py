"""CPU oracle for the chromosight hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU arm -- never as a fallback of ``chromosight_b200``.

Parity status: PINNED.  Every function here is checked against outputs of the
unmodified reference (koszullab/chromosight @ ecb32c5, imported from
/root/reference in the build container) through the committed fixtures in
``tests/golden/`` (generator: ``tests/golden/make_golden.py``) and against the
reference's own known-answer tests (``tests/test_detection.py``,
``tests/test_preprocessing.py`` of the reference).
"""
