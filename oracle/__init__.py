"""CPU oracle for the mcframework hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``mcframework_b200`` never does.

Two layers:

* :mod:`oracle.c_oracle`  -- ctypes front-end of ``mcf_oracle.c``: a plain-C restatement of
  the NumPy algorithms the reference calls (SeedSequence, PCG64, Philox, uniform, ziggurat
  normal, Lemire bounded integers, pairwise ``np.sum``) and of the built-in simulations.
* :mod:`oracle.np_port`   -- a NumPy-level restatement of the reference's replication loop
  (``core.py:583-752``), simulations and ``StatsEngine`` metrics.  It has the same
  performance character as the reference (interpreter loop around small NumPy calls) and is
  what ``bench.py`` times as ``cpu_baseline`` (kind ``"port"``).

Pinned against: NumPy's own known-answer files, live NumPy draws, and fixtures produced
by running the unmodified reference (``tools/gen_golden.py`` -> ``tests/golden/``).
"""
