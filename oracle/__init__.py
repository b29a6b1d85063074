"""CPU oracle for the dynsight hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()``, ``tests/golden/make_golden.py`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package.  The product
package ``dynsight_b200`` never does (``tests/test_boundary.py`` greps for it).
"""
