"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's OAK Gram / SGPR-statistics / Sobol hot path.
Nothing in the product package may import this; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and there only as the checker / the timed CPU comparator.
"""
