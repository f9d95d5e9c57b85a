"""CPU oracle for the NUTS/HMC hot path of numpyro 0.21.0 -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm (numpyro/infer/hmc_util.py,
hmc.py, util.py, mcmc.py plus the Threefry PRNG that lives in the un-vendored dependency
jax>=0.7).  It exists to *check* the CUDA engine; nothing under ``numpyro_b200/`` may import
it.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl
reference legs use it.

Parity pin status (see DESIGN.md "Oracle"):
  * Threefry-2x32 is pinned by the three Random123 known-answer vectors and by the values
    JAX's documentation prints for split(key(0)) / normal(key(42)).
  * adaptation windows, checkpoint index tables, iterative U-turn truth table, warm-up
    script, diagnostics known answers are pinned by the reference's own tests
    (test/infer/test_hmc_util.py, test/test_diagnostics.py) restated in tests/.
  * No NUTS sample vector exists anywhere in the reference and JAX cannot be imported in
    this image, so sampled *values* are "parity unpinned" against real numpyro; they are
    pinned statistically (README posterior table) and structurally (reference invariants).

Numeric convention ("det-f32"): every float operation of the tree/adaptation bookkeeping is
IEEE binary32, round-to-nearest, one rounding per written operation, never contracted into
FMA; reductions over the latent dimension use the lane-strided + butterfly order of
``oracle.detmath.lane_sum``; exp/log/log1p are the explicit algorithms in
``oracle.detmath``.  The CUDA engine follows the same convention so integer bookkeeping
(tree depth, number of leapfrogs, directions, divergences, proposal choices) is bit-exact.
"""
