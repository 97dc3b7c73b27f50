"""TEST INFRASTRUCTURE ONLY.

``oracle/`` holds the CPU restatement of LUCI's per-spaxel fitting path (numpy + SciPy SLSQP) and
the import shims that let the *real* reference run in the build container.  Nothing under
``luci_b200/`` may import from here; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and only as the checker.
"""
