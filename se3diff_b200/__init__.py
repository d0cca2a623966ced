"""se3diff_b200 -- B200 (sm_100a) implementation of the batched SE(3) reverse-diffusion sampling
step of ddrichman/SE3Diff's vendored BioEmu denoiser, behind BioEmu's own call surface.

Public surface mirrors `bioemu.shortcuts` (reference: bioemu/src/bioemu/shortcuts.py:4-16):
    se3diff_b200.shortcuts.{dpm_solver, euler_maruyama_predictor, euler_maruyama_predictor_finetune,
                            heun_denoiser, heun_denoiser_finetune, DiGConditionalScoreModel,
                            DiGSO3SDE, CosineVPSDE}
All device work goes through hand-written CUDA kernels in libse3diff_b200.so (C ABI declared in
include/se3diff_b200.h).  There is no CPU fallback: calling an op without the library or without a
CUDA tensor raises.
"""
__version__ = "0.1.0"
