"""TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's hot-path algorithms.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package; the product path (``monocularsfm_b200``) never does.
"""
