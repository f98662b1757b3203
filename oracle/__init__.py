"""CPU oracle for the pb_bss EM / beamforming hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``pb_bss_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` do, and only as the checker or as
the timed CPU baseline -- never as the product path.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified
reference from ``/root/reference`` (possible only in the build container),
runs it on seeded inputs and stores its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every oracle function against those
fixtures and against the reference's own doctest / unit-test known answers.
"""
