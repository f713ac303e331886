"""CPU oracle for the VLN-CE policy hot path.  TEST INFRASTRUCTURE ONLY.

This package is the parity checker for ``vlnce_b200``: a plain PyTorch-CPU fp32
restatement of (i) the arithmetic the reference borrows from the un-vendored
dependency habitat-lab / habitat_baselines v0.1.7 (``oracle/hb017.py``) and
(ii) the reference's own model / loss code for the path
(``oracle/vlnce.py``, each function citing the reference file:line it follows).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  Nothing under ``vlnce_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

Pinning status
--------------
* ``oracle/vlnce.py`` is pinned against the UNMODIFIED reference files
  (``/root/reference/vlnce_baselines/models/**``, ``common/aux_losses.py``)
  imported in the build container through ``oracle/reference_import.py``; the
  pinning test is ``tests/test_oracle_vs_reference.py`` and the golden vectors
  it freezes live in ``tests/golden/`` (generator: ``oracle/make_golden.py``).
* ``oracle/hb017.py`` restates habitat_baselines v0.1.7 from its published
  source; that package is absent from ``/root/reference`` and from this image and
  the reference holds no test vectors for it, so this part is **parity
  unpinned** at the dependency boundary.  The reference's own
  call sites are exercised through it.
"""
