"""CPU oracle of the ReReVST hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here, and only as the checker (or as the
timed CPU baseline), never as part of the product path.
"""
