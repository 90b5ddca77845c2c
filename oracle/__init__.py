"""CPU oracle for the frame-rate analysis hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diffsptk_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  It is the checker, never the
thing shipped or measured as the product.

Parity status: PINNED.  ``oracle/np_oracle.py`` is a numpy restatement of the
reference's algorithms (each function cites the reference file:line it
follows).  It is pinned two ways:

* against the 13 doctest known answers stored in the reference's own sources
  (``tests/test_oracle_doctest_vectors.py``), and
* against outputs of the reference itself, produced in the build container by
  importing ``/root/reference`` (``tests/golden/make_golden.py`` writes
  ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks them, and
  ``tests/test_oracle_vs_reference.py`` re-runs the live comparison whenever
  ``/root/reference`` is present).
"""
