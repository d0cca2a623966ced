"""TEST INFRASTRUCTURE -- CPU restatement of the SO(3)-only toy layer `se3diff/{models,train,finetune}.py` and of the
path functionals of `bioemu/src/bioemu/ppft.py` it uses.  torch on the CPU, same arithmetic order and the same order
of random draws as the reference, so that `tests/golden/toy.npz` (minted from the unmodified reference by
`oracle/gen_golden.py`) is reproduced exactly.  Nothing outside tests/, smoke() and bench.py's CPU arm may import this.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import so3 as O
from .samplers import EM
from .score_model import sinusoid


class ScoreNetOracle:
    """se3diff/models.py:9-61 evaluated from a state_dict (rot_embed.{0,1}, net.{0,2,4}; time_embed.dummy is empty)."""

    def __init__(self, state_dict: dict, time_embed_dim: int = 32):
        self.p = {k: v.clone().requires_grad_(v.is_floating_point() and v.numel() > 0) for k, v in state_dict.items()}
        self.time_embed_dim = time_embed_dim

    def parameters(self):
        return [v for v in self.p.values() if v.requires_grad]

    def __call__(self, rot_mat: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        p = self.p
        v = O.rotmat_to_rotvec(rot_mat)                                                     # models.py:47
        e = F.relu(F.layer_norm(F.linear(v, p["rot_embed.0.weight"], p["rot_embed.0.bias"]), (p["rot_embed.1.weight"].numel(),),
                                p["rot_embed.1.weight"], p["rot_embed.1.bias"]))            # models.py:49
        te = sinusoid(t, self.time_embed_dim)                                               # models.py:51 (max_input 1000 => t unscaled)
        x = torch.cat(torch.broadcast_tensors(e, te), dim=-1)                               # models.py:53-55
        x = F.relu(F.linear(x, p["net.0.weight"], p["net.0.bias"]))
        x = F.relu(F.linear(x, p["net.2.weight"], p["net.2.bias"]))
        return F.linear(x, p["net.4.weight"], p["net.4.bias"])                              # models.py:57


def sample_multiple_igso3(tab: O.SO3Tables, mus, sigmas, weights, num_samples: int):
    """DiGMixSO3SDE.sample_multiple_igso3 (se3diff/models.py:64-89): multinomial, then randn(n,1,3), rand(n,1)."""
    k = torch.multinomial(weights, num_samples, replacement=True)
    r = tab.sample_igso3(sigmas[k]).squeeze(-3)
    return mus[k] @ r


def so3_score(x_t, tab: O.SO3Tables, score_model, t):
    """_get_so3_score (se3diff/train.py:19-38)."""
    return score_model(x_t, t) * tab.score_scaling_at(t).unsqueeze(-1)


def _em_step(em: EM, tab: O.SO3Tables, x, t, dt, score, u=None):
    """EulerMaruyamaPredictor.update_given_score on bare rotations: batch_idx None (denoiser.py:54-116)."""
    w = 0.5 * em.mcf * (1 + em.nw**2)
    g = tab.beta(t)[:, None]                                       # maybe_expand(beta(t), None, drift)
    drift = torch.zeros_like(x[..., 0]) - g**2 * score * w
    if u is not None:
        drift = drift + g * u * w
    return em.update(x, dt, drift, g)


@torch.no_grad()
def reverse_diffusion(tab: O.SO3Tables, score_model, batch_size: int, num_steps: int, finetune_model=None):
    """reverse_diffusion (se3diff/train.py:41-77) / reverse_finetune_diffusion (se3diff/finetune.py:17-66)."""
    x = tab.prior(batch_size)
    em = EM("rot", None, tab)
    ts = torch.linspace(1.0, 0.0, num_steps + 1)
    dts = torch.diff(ts)
    xs, us, dws = [x], [], []
    for i in range(num_steps):
        t = torch.full((batch_size,), ts[i].item())
        score = so3_score(x, tab, score_model, t)
        u = finetune_model(x, t) if finetune_model is not None else None
        x, _, dw = _em_step(em, tab, x, t, dts[i], score, u)
        xs.append(x)
        us.append(u)
        dws.append(dw)
    if finetune_model is None:
        return torch.stack(xs), ts
    return torch.stack(xs), ts, torch.stack(us), torch.stack(dws)


def igso3_mixture_marginal_pdf(mus, sigmas, weights, l_max: int = 1000, num_points: int = 1000, tol: float = 1e-7):
    """se3diff/train.py:80-110 (l = 0..l_max-1)."""
    omega = torch.linspace(0, math.pi, num_points)
    omega_0 = O.angle_from_rotmat(mus)[0]
    pdfs = O.igso3_marginal_pdf(omega.unsqueeze(0), omega_0.unsqueeze(1), sigmas.unsqueeze(1), torch.arange(l_max), tol=tol)
    return omega, torch.clamp((weights.unsqueeze(-1) * pdfs).sum(dim=0), min=0.0)


def compute_train_loss(tab: O.SO3Tables, score_model, mus, sigmas, weights, batch_size: int, tol: float = 1e-7):
    """Denoising score matching on the mixture (se3diff/train.py:113-143).  Draw order: multinomial, IGSO3 normals +
    uniforms of x_0, rand(t), IGSO3 normals + uniforms of x_t."""
    x_0 = sample_multiple_igso3(tab, mus, sigmas, weights, batch_size)
    t = torch.rand(batch_size)
    x_t = tab.sample_marginal(x_0, t)
    q_t = O.rotmat_to_rotvec(torch.einsum("...ki,...kj->...ij", x_0, x_t))
    true = tab.compute_score(q_t, t)
    pred = score_model(x_t, t)
    return F.mse_loss(pred, true / (tab.score_scaling_at(t).unsqueeze(-1) + tol))


def assign_igso3(x_0, mus, sigmas, weights, l_max: int = 1000, tol: float = 1e-7):
    """Mixture responsibilities (se3diff/finetune.py:69-93)."""
    rel = torch.einsum("k...ij,b...il->bk...jl", mus, x_0)
    ang = O.angle_from_rotmat(rel)[0]
    pdf = O.igso3_expansion(ang, sigmas, torch.arange(l_max), tol=tol) * weights
    return pdf / (torch.sum(pdf, dim=-1, keepdim=True) + tol)


# ---- path functionals (ppft.py:4-194) ------------------------------------------------------------------------
def ito_integral(fs, dWs):                      # ppft.py:4-14
    return torch.einsum("tb...i,tb...i->b...", fs, dWs)


def quadratic_covariation(fs, gs, dts):         # ppft.py:17-29
    return torch.einsum("tb...i,tb...i,t->b...", fs, gs, dts)


def rloo_baseline(fs):                          # ppft.py:32-43
    return (fs.sum(dim=0, keepdim=True) - fs) / (fs.shape[0] - 1)


def compute_ev_loss(ws, hs, h_stars, tol=1e-7):  # ppft.py:81-139, from_int_dws=True, use_stab=True
    B = ws.shape[0]
    w = ws.unsqueeze(1)
    dhs = hs - h_stars
    if B > 1:
        pbar = torch.mean(hs, dim=0)
        stab = torch.sum(pbar, dim=0) / (pbar + tol)
        stab = stab / torch.mean(stab)
    else:
        stab = torch.tensor(1.0)
    s1, s2, s3 = torch.sum(w * dhs, dim=0), torch.sum(dhs, dim=0), torch.sum(w * dhs**2, dim=0)
    return torch.sum(2 * (s1 * s2 - s3) * stab / (B * (B - 1)))


def compute_kl_loss(ws, int_u_u_dt, int_u_u_dt_sg):   # ppft.py:155-194, from_int_dws=True, use_rloo=True
    baseline, baseline_sg = rloo_baseline(int_u_u_dt.detach()), rloo_baseline(int_u_u_dt_sg)
    return torch.mean(int_u_u_dt - baseline + (int_u_u_dt_sg - baseline_sg) * ws) / 2


def compute_finetune_loss(tab: O.SO3Tables, score_model, finetune_model, mus, sigmas, h_stars, lambda_: float = 0.1,
                          batch_size: int = 4096, num_steps: int = 200, l_max: int = 1000, tol: float = 1e-7):
    """se3diff/finetune.py:96-143."""
    xs, ts, us_sg, dWs = reverse_diffusion(tab, score_model, batch_size, num_steps, finetune_model)
    us = torch.stack([finetune_model(xs[i], torch.full((batch_size,), ts[i].item())) for i in range(num_steps)])
    hs = assign_igso3(xs[-1], mus, sigmas, h_stars, l_max=l_max, tol=tol)
    dts = torch.diff(ts)
    int_u_u_dt = quadratic_covariation(us, us, -dts)
    int_u_u_dt_sg = quadratic_covariation(us_sg, us_sg, -dts)
    int_dws = ito_integral(us, -dWs)
    return compute_ev_loss(int_dws, hs, h_stars, tol) + lambda_ * compute_kl_loss(int_dws, int_u_u_dt, int_u_u_dt_sg)
