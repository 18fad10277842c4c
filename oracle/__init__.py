"""CPU oracle for the sypha Mehrotra IPM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``sypha_b200/`` may import this package:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs are allowed to, and there only as the checker or as
the CPU side of a timing comparison.

The oracle is a NumPy restatement of the algorithm executed by the reference's
``solver_sparse_mehrotra_run`` (/root/reference/src/sypha_solver.cpp:42-886) and
its Python prototype (/root/reference/python/interior_point.py:60-194).

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the reference's
own ``interior_point.mehrotra_linopt_dense`` in the build container and stores
its iterates; ``tests/test_oracle.py`` checks the oracle against those fixtures
and against the reference's known-answer tables
(python/sypha_unit_tests.py:21-77, benchmark/results/benchmark_results_with_ip.csv).
"""
