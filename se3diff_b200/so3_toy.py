"""The SO(3)-only toy layer of the reference -- host mirror of `se3diff/models.py`, `se3diff/train.py` and
`se3diff/finetune.py` (same class / function names, arguments and return values) on top of this library's kernels.

Per reverse step the reference runs `rotmat_to_rotvec` (~35 ATen launches), the MLP, a table gather and the
Euler-Maruyama update (~40 launches).  Here a step is: one log-map kernel (se3_so3_log), the MLP (cuBLAS), and ONE
fused kernel (se3_so3_update_em) that converts the network output to a score, forms the reverse drift, draws nothing
(the normals come in) and applies both exponential maps; the per-step quantities g(t), score scaling, dt are kernel
arguments computed once per call.  Losses stay torch expressions (they need autograd through the small MLP)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, pathwise, schedule
from . import sdes as S
from ._lib import EmScalars
from .denoiser import EulerMaruyamaPredictor
from .models import SinusoidalPositionEmbedder


class ScoreNet(nn.Module):
    """3-vector score from a rotation and a time (se3diff/models.py:9-61).  Same sub-module names and registration order,
    hence the same state_dict and the same seeded initialisation (the reference's Xavier loop iterates over parameters,
    never matches nn.Linear, and so leaves PyTorch's default init in place)."""

    def __init__(self, rot_embed_dim: int = 32, time_embed_dim: int = 32, hidden_dim: int = 128):
        super().__init__()
        self.rot_embed = nn.Sequential(nn.Linear(3, rot_embed_dim), nn.LayerNorm(rot_embed_dim), nn.ReLU())
        self.time_embed = SinusoidalPositionEmbedder(time_embed_dim)
        self.net = nn.Sequential(nn.Linear(rot_embed_dim + time_embed_dim, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, hidden_dim),
                                 nn.ReLU(), nn.Linear(hidden_dim, 3))

    def forward(self, rot_mat: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        rot_vec = S.rotmat_to_rotvec(rot_mat)                     # se3_so3_log; an input, not differentiated
        x = torch.cat(torch.broadcast_tensors(self.rot_embed(rot_vec), self.time_embed(t)), dim=-1)
        return self.net(x)


class DiGMixSO3SDE(S.DiGSO3SDE):
    """se3diff/models.py:64-89."""

    def sample_multiple_igso3(self, mus, sigmas, weights, num_samples: int, device=None) -> torch.Tensor:
        dev = self.igso3.cdf_igso3.device if device is None else torch.device(device)
        mus, sigmas, weights = mus.to(dev), sigmas.to(dev), weights.to(dev)
        k = S.noise_multinomial(weights, num_samples)             # mixture component per sample
        return self.igso3.sample(sigmas[k], 1, left=mus[k].contiguous()).squeeze(-3)   # mu_k . r in the sampling kernel


def _get_so3_score(x_t, sde, score_model, t):
    """se3diff/train.py:19-38."""
    return score_model(x_t, t) * sde.get_score_scaling(t).unsqueeze(-1)


def _so3_em_scalars(sde, ts: torch.Tensor, dts: torch.Tensor) -> list[EmScalars]:
    """Per-step constants of EulerMaruyamaPredictor(noise_weight=1, mcf=1) on a bare SO(3) state."""
    t = ts[:-1].clone()
    g, sc, sdt = schedule.so3_g(sde, t), schedule.so3_score_scaling(sde, t), torch.sqrt(dts.abs())
    return [EmScalars(float(dts[i]), float(sdt[i]), 1.0, 1.0, float(g[i]), float(sc[i]), 0.0, 0.0, 1.0, float(sde.tol))
            for i in range(t.numel())]


def _rollout(sde, score_model, finetune_model, device, batch_size, num_steps):
    if device is None or torch.device(device).type != "cuda":
        raise RuntimeError("se3diff_b200 samplers run on a CUDA device only (there is no CPU fallback)")
    device = torch.device(device)
    sde.to(device)
    for m in (score_model, finetune_model):
        if isinstance(m, nn.Module):
            m.to(device)
    x_t = sde.prior_sampling((batch_size, 3, 3), device=device)
    ts = torch.linspace(1.0, 0.0, num_steps + 1)
    dts = torch.diff(ts)
    consts = _so3_em_scalars(sde, ts, dts)
    xs, us, dWs = [x_t], [], []
    for i in range(num_steps):
        t = torch.full((batch_size,), float(ts[i]), device=device)
        m_rot = score_model(x_t, t)                                # raw network output; the scaling happens in the kernel
        u = finetune_model(x_t, t) if finetune_model is not None else None
        z = S.noise_randn((batch_size, 3), device)
        x_t, dW = ops.so3_update_em(x_t, m_rot, z, consts[i], u_rot=u, want_dw=finetune_model is not None)
        xs.append(x_t)
        us.append(u)
        dWs.append(dW)
    return xs, ts.to(device), us, dWs


@torch.no_grad()
def reverse_diffusion(sde, score_model, *, device=None, batch_size: int = 4096, num_steps: int = 200):
    """se3diff/train.py:41-77 -> (xs [T+1,B,3,3], timesteps [T+1])."""
    xs, ts, _, _ = _rollout(sde, score_model, None, device, batch_size, num_steps)
    return torch.stack(xs, dim=0), ts


@torch.no_grad()
def reverse_finetune_diffusion(sde, score_model, finetune_model, *, device=None, batch_size: int = 4096, num_steps: int = 200):
    """se3diff/finetune.py:17-66 -> (xs [T+1,B,3,3], timesteps, us [T,B,3], dWs [T,B,3])."""
    xs, ts, us, dWs = _rollout(sde, score_model, finetune_model, device, batch_size, num_steps)
    return torch.stack(xs, dim=0), ts, torch.stack(us, dim=0), torch.stack(dWs, dim=0)


def igso3_mixture_marginal_pdf(mus, sigmas, weights, l_max: int = 1000, num_points: int = 1000, tol: float = 1e-7):
    """Angle density of the IGSO(3) mixture (se3diff/train.py:80-110): one se3_igso3_marginal_pdf launch over [K, points]."""
    dev = mus.device
    omega = torch.linspace(0, math.pi, num_points, device=dev)
    omega_0 = S.angle_from_rotmat(mus)[0]
    pdfs = S.igso3_marginal_pdf(omega.unsqueeze(0), omega_0.unsqueeze(1), sigmas.unsqueeze(1), torch.arange(l_max), tol=tol)
    return omega, torch.clamp((weights.unsqueeze(-1) * pdfs).sum(dim=0), min=0.0)


def compute_train_loss(sde, score_model, mus, sigmas, weights, device=None, batch_size: int = 4096, tol: float = 1e-7):
    """Denoising score matching against the mixture (se3diff/train.py:113-143)."""
    x_0 = sde.sample_multiple_igso3(mus, sigmas, weights, batch_size, device=device)
    dev = x_0.device
    t = S.noise_rand((batch_size,), dev)
    x_t = sde.sample_marginal(x_0, t)
    q_t = ops.so3_rel_log(x_0, x_t)                                # Log(x_0^T x_t), one kernel
    true_score = sde.compute_score(q_t, t)
    pred_score = score_model(x_t, t)
    return F.mse_loss(pred_score, true_score / (sde.get_score_scaling(t).unsqueeze(-1) + tol))


def assign_igso3(x_0, mus, sigmas, weights, l_max: int = 1000, tol: float = 1e-7):
    """Responsibility of each mixture component for each sample (se3diff/finetune.py:69-93) -> [B, K]."""
    B, K = x_0.shape[0], mus.shape[0]
    rel = ops.so3_matmul(mus.unsqueeze(0).expand(B, K, 3, 3).contiguous(), x_0.unsqueeze(1).expand(B, K, 3, 3).contiguous(),
                         transpose_a=True)                        # mu_k^T x_b
    ang = S.angle_from_rotmat(rel.view(B, K, 3, 3))[0]
    pdf = S.igso3_expansion(ang, sigmas.to(ang.device).expand(B, K), torch.arange(l_max), tol=tol) * weights.to(ang.device)
    return pdf / (pdf.sum(dim=-1, keepdim=True) + tol)


def compute_finetune_loss(sde, score_model, finetune_model, mus, sigmas, h_stars, device=None, lambda_: float = 0.1,
                          batch_size: int = 4096, num_steps: int = 200, l_max: int = 1000, tol: float = 1e-7):
    """se3diff/finetune.py:96-143: no-grad rollout, then the controls are re-evaluated with gradients on the stored states."""
    xs, timesteps, us_sg, dWs = reverse_finetune_diffusion(sde, score_model, finetune_model, device=device, batch_size=batch_size,
                                                           num_steps=num_steps)
    ts_host = timesteps.cpu()
    us = torch.stack([finetune_model(xs[i], torch.full((batch_size,), float(ts_host[i]), device=xs.device))
                      for i in range(num_steps)], dim=0)
    hs = assign_igso3(xs[-1], mus.to(xs.device), sigmas.to(xs.device), h_stars.to(xs.device), l_max=l_max, tol=tol)
    dts = torch.diff(timesteps)
    int_u_u_dt = pathwise.compute_int_u_u_dt(us=us, dts=dts)
    int_u_u_dt_sg = pathwise.compute_int_u_u_dt(us=us_sg, dts=dts)
    int_dws = pathwise.compute_int_dws(us=us, dWs=dWs)
    loss_ev = pathwise.compute_ev_loss(ws=int_dws, hs=hs, h_stars=h_stars.to(xs.device), tol=tol)
    loss_kl = pathwise.compute_kl_loss(ws=int_dws, int_u_u_dt=int_u_u_dt, int_u_u_dt_sg=int_u_u_dt_sg)
    return loss_ev + lambda_ * loss_kl


__all__ = ["ScoreNet", "DiGMixSO3SDE", "EulerMaruyamaPredictor", "reverse_diffusion", "reverse_finetune_diffusion",
           "igso3_mixture_marginal_pdf", "compute_train_loss", "assign_igso3", "compute_finetune_loss"]
