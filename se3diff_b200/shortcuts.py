"""`_target_` names for hydra-style configs -- the plug-in point of the reference
(bioemu/src/bioemu/shortcuts.py:4-16; used by sample.py:120-138 and finetune.py:150-188).
Pointing a denoiser / model YAML at `se3diff_b200.shortcuts.<Name>` selects this implementation."""
from .denoiser import (  # noqa: F401
    dpm_solver,
    euler_maruyama_predictor,
    euler_maruyama_predictor_finetune,
    heun_denoiser,
    heun_denoiser_finetune,
    sde_dpm_solver_finetune,
)
from .models import DiGConditionalScoreModel  # noqa: F401
from .sdes import CosineVPSDE, DiGSO3SDE  # noqa: F401
