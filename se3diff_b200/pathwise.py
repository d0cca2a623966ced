"""Path functionals of the fine-tuning objective -- the host mirror of `bioemu/src/bioemu/ppft.py:4-194` (same names,
argument meaning and defaults).  These are reductions over stored controls `us [T,B,...,D]` and Brownian increments
`dWs [T,B,...,D]` that must stay differentiable with respect to `us`, so they are torch expressions running on the
device the tensors live on; the sampling path that produces `us`/`dWs` is where the CUDA kernels are."""
from __future__ import annotations

import torch


def riemannian_ito_integral(fs: torch.Tensor, dWs: torch.Tensor) -> torch.Tensor:
    """sum_t <f_t, dW_t>  -> [B, ...] (ppft.py:4-14)."""
    return (fs * dWs).sum(dim=-1).sum(dim=0)


def riemannian_quadratic_covariation(fs: torch.Tensor, gs: torch.Tensor, dts: torch.Tensor) -> torch.Tensor:
    """sum_t <f_t, g_t> dt_t -> [B, ...] (ppft.py:17-29)."""
    inner = (fs * gs).sum(dim=-1)
    return (inner * dts.reshape((-1,) + (1,) * (inner.dim() - 1))).sum(dim=0)


def rloo_baseline(fs: torch.Tensor) -> torch.Tensor:
    """Leave-one-out mean of the other samples (ppft.py:32-43)."""
    return (fs.sum(dim=0, keepdim=True) - fs) / (fs.shape[0] - 1)


def compute_ws(*, us, dWs, dts):
    """Importance weights exp(int <u - sg(u), dW> - 1/2 int |u - sg(u)|^2 dt), reverse time (ppft.py:46-63)."""
    diff = us - us.detach()
    return torch.exp(riemannian_ito_integral(diff, -dWs) - riemannian_quadratic_covariation(diff, diff, -dts) / 2)


def compute_int_dws(*, us, dWs):
    """int <u, dW> in reverse time: its gradient equals the gradient of the importance weight (ppft.py:66-78)."""
    return riemannian_ito_integral(us, -dWs)


def compute_int_u_u_dt(*, us, dts):
    """int |u|^2 dt in reverse time (ppft.py:142-152)."""
    return riemannian_quadratic_covariation(us, us, -dts)


def compute_ev_loss(*, ws, hs, h_stars, from_int_dws: bool = True, use_stab: bool = True, tol: float = 1e-7):
    """U-statistic estimator of sum_k (E[h_k] - h*_k)^2 from importance weights or their integrated gradients
    (ppft.py:81-139; the reference's debugging print of its arguments is not reproduced)."""
    B = ws.shape[0]
    w = ws.unsqueeze(1)
    dhs = hs - h_stars
    if use_stab and B > 1:
        pbar = hs.mean(dim=0)
        stab = pbar.sum(dim=0) / (pbar + tol)
        stab = stab / stab.mean()
    else:
        stab = torch.tensor(1.0, device=ws.device)
    if from_int_dws:
        s1, s2, s3 = (w * dhs).sum(dim=0), dhs.sum(dim=0), (w * dhs**2).sum(dim=0)
        per_k = 2 * (s1 * s2 - s3) * stab / (B * (B - 1))
    else:
        wd = w * dhs
        per_k = (wd.sum(dim=0) ** 2 - (wd**2).sum(dim=0)) * stab / (B * (B - 1))
    return per_k.sum()


def compute_kl_loss(*, ws, int_u_u_dt, int_u_u_dt_sg, from_int_dws: bool = True, use_rloo: bool = True):
    """1/2 E[ int |u|^2 dt ] with a REINFORCE leave-one-out baseline (ppft.py:155-194)."""
    if use_rloo:
        baseline, baseline_sg = rloo_baseline(int_u_u_dt.detach()), rloo_baseline(int_u_u_dt_sg)
    else:
        baseline, baseline_sg = torch.zeros_like(int_u_u_dt), torch.zeros_like(int_u_u_dt_sg)
    if from_int_dws:
        integrand = int_u_u_dt - baseline + (int_u_u_dt_sg - baseline_sg) * ws
    else:
        integrand = (int_u_u_dt - baseline) * ws
    return integrand.mean() / 2
