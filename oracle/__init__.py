"""CPU oracle for the geometric message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker or the reported CPU baseline.

What it is: a pure-PyTorch (CPU, fp32/fp64) restatement of

* the reference's own layer code (``oracle/ref_layers.py``; every function
  cites the reference file:line it follows), and
* the third-party libraries the reference calls but which are neither vendored
  in ``/root/reference`` nor installable here (``oracle/thirdparty``:
  ``torch_scatter``, the PyG ``MessagePassing``/``SchNet`` pieces, ``e3nn``
  ``o3``/``nn``, ``opt_einsum.contract``, ``torch_cluster.radius_graph``).

Parity status
-------------
* Reference-owned code: PINNED.  ``tests/golden/make_golden.py`` imports the
  unmodified reference modules from ``/root/reference`` (with
  ``oracle/shims`` standing in for the missing wheels) and stores their
  outputs; ``tests/test_oracle_golden.py`` checks ``oracle/ref_layers.py``
  against those vectors.
* Third-party semantics (e3nn 0.4.4/0.5.1, PyG 2.3.1, torch_scatter,
  torch_cluster): PARITY UNPINNED.  The wheels cannot be run here, so these
  are restated from the published algorithms and checked by closed-form and
  property tests (``tests/test_oracle_thirdparty.py``): SH orthonormality,
  ``wigner_3j`` known values and O(3) invariance, ``normalize2mom``
  constants, equivariance of every layer.
"""
