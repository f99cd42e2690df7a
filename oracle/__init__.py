"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MonoPort occupancy hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only
as the checker / reported CPU baseline -- never as the thing measured or shipped.
The product path (``monoport_b200``) raises if its CUDA library is missing.

Parity status:
  * query() restatement (``oracle.spec.query_ref``)        -- PINNED against the reference's own
    ``MonoPortNet.query`` run in the build container (``tests/golden/query_*.npz``,
    generator ``tests/golden/make_golden.py``).
  * forward_vertices restatement                           -- PINNED the same way (``fv_*.npz``).
  * Seg3dLossless / Seg3dTopk restatement                  -- PARITY UNPINNED (third-party
    ``implicit-seg`` is un-vendored and unpinned in the reference: requirements.txt:15).
  * marching cubes                                         -- PARITY UNPINNED (absent from the
    reference).
"""
