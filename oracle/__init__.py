"""CPU ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain torch-on-CPU restatement of the reference algorithm for the SE(3) reverse-diffusion
sampling path of ddrichman/SE3Diff (vendored BioEmu 0.1.12).  Each function cites the reference
file:line it follows (paths relative to /root/reference/bioemu/src/bioemu/ unless stated).

Who may import this package: ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker or as the timed CPU
baseline, never as the thing shipped.  ``se3diff_b200`` never imports it; the product path raises
when its CUDA library is missing.

Pinning: the restatement is checked (tests/test_oracle_golden.py) against
  * the reference's own golden vector for the score model (bioemu/tests/expected.npz, regenerated
    through the unmodified reference into tests/golden/score_model_tiny.npz),
  * known-answer tests mirroring bioemu/tests/test_so3_utils.py (scipy Rotation, matrix_exp),
  * fixtures produced by importing the unmodified reference in the build container
    (oracle/gen_golden.py -> tests/golden/*.npz), covering every SO(3)/IGSO3/R3 function, the lookup
    tables, and full dpm / Euler-Maruyama / Heun trajectories.
Why torch and not numpy/C: the reference itself is fp32 torch-on-CPU; using the same ATen scalar
kernels is the only way to restate it bit-for-bit (the fixtures above agree to 0 ulp for the
SDE algebra and <= 2e-7 for the network).
"""
