"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement (torch-CPU fp32 / numpy) of the GSR-Net hot path of
denisfouchard/fMRI-connectivity-super-resolution.  It exists so that the CUDA
path can be checked against the reference's arithmetic on a box where
``/root/reference`` is not mounted.

Rules (enforced by tests/test_abi_and_host.py::test_product_package_never_imports_the_oracle):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``fmri-connectivity-super-resolution_b200/`` imports it: the
    product path has no CPU fallback and fails loudly without its CUDA library.

Pinning status: the reference ships NO tests / golden vectors for this path
(SURVEY.md section 4), so the pin is made here: ``oracle/make_golden.py``
imports the UNMODIFIED reference from ``/root/reference`` in the build
container, runs it on seeded inputs and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` holds the restatement to
those vectors on every run.  ``oracle/mlp_oracle.py`` (the MLP baseline,
models/mlp.py) is pinned the same way by ``oracle/make_golden_mlp.py`` ->
``tests/golden/mlp.npz`` -> ``tests/test_mlp.py``.
"""
