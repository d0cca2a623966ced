"""Host-side evaluation of the per-step schedule scalars.

The reference evaluates alpha(t), sigma(t), lambda, h, t_lambda, beta(t), g(t) and the score scaling
on `[B]`/`[N,3]` device tensors every step and pulls two of them back with `.item()`
(denoiser.py:669, 677-695).  All graphs of a batch share one t per step, so these are *scalars*:
here they are computed once per sampler call, for all steps, with the SDE objects' own torch
expressions on CPU fp32 tensors (bit-identical to the reference's CPU path), and handed to the
fused kernels by value -- no per-step host sync, no gathers, CUDA-graph capturable.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from ._lib import DpmScalars, EmScalars, HeunScalars


def _f(x) -> float:
    return float(x.item() if torch.is_tensor(x) else x)


def _so3_tables_cpu(so3):
    """CPU copies of (sigma_grid, score_scaling) of a DiGSO3SDE-like object, cached on the object."""
    sf = getattr(so3, "score_function", None)
    if sf is None or not hasattr(sf, "score_scaling"):
        return None
    key = (sf.sigma_grid.data_ptr(), sf.score_scaling.data_ptr())
    cached = getattr(so3, "_se3_cpu_tables", None)
    if cached is None or cached[0] != key:
        cached = (key, sf.sigma_grid.detach().cpu(), sf.score_scaling.detach().cpu())
        object.__setattr__(so3, "_se3_cpu_tables", cached)
    return cached[1], cached[2]


def _each(fn, t: torch.Tensor) -> torch.Tensor:
    """Evaluate `fn` element by element.  ATen's vectorised fp32 `pow` (used by the geometric sigma
    schedule, so3_sde.py:378) differs from its scalar path by 1 ulp on some inputs; the reference sees the
    scalar path for small batches, so the schedule is evaluated that way too."""
    return torch.cat([fn(t[i:i + 1]) for i in range(t.numel())]) if t.numel() else t.clone()


def so3_sigma(so3, t):
    return _each(so3._marginal_std, t)


def so3_g(so3, t):
    return _each(so3.beta, t)


def so3_score_scaling(so3, t: torch.Tensor) -> torch.Tensor:
    """get_score_scaling(t) (so3_sde.py:142-161, 1610-1635) for CPU t."""
    tabs = _so3_tables_cpu(so3)
    if tabs is None:  # foreign SO3SDE: go through its public method on its own device
        return _each(lambda x: so3.get_score_scaling(x).detach().cpu(), t)
    grid, scaling = tabs
    return scaling[torch.bucketize(so3_sigma(so3, t), grid)]


def r3_alpha(r3, t):
    return r3._marginal_mean_coeff(t)


def r3_std(r3, t):
    return torch.sqrt(1.0 - r3._marginal_mean_coeff(t) ** 2)


def t_from_lambda(r3, lam: torch.Tensor) -> torch.Tensor:
    """Inverse of lambda(t) for the cosine schedule (denoiser.py:623-631)."""
    f = -1 / 2 * torch.log(torch.exp(-2 * lam) + 1)
    e = f + torch.log(torch.cos(torch.tensor(np.pi * r3.s / 2 / (1 + r3.s))))
    return 2 * (1 + r3.s) / np.pi * torch.acos(torch.exp(e)) - r3.s


def timesteps(max_t: float, min_t: float, num_steps: int):
    ts = torch.linspace(max_t, min_t, num_steps + 1)  # denoiser.py:231, 663 (fp32: dt is not constant)
    return ts, torch.diff(ts)


@dataclass
class DpmStep:
    t: float
    t_lambda: float
    scalars: DpmScalars


def dpm_schedule(r3, so3, num_steps: int, max_t: float, min_t: float) -> list[DpmStep]:
    """Per-step constants of dpm_solver (denoiser.py:668-762)."""
    ts, dts = timesteps(max_t, min_t, num_steps)
    t = ts[:-1].clone()
    t_next = t + dts
    a_t, s_t = r3_alpha(r3, t), r3_std(r3, t)
    a_n, s_n = r3_alpha(r3, t_next), r3_std(r3, t_next)
    lam, lam_n = torch.log(a_t / s_t), torch.log(a_n / s_n)
    h = lam_n - lam
    t_lam = t_from_lambda(r3, (lam + lam_n) / 2)
    a_l, s_l = r3_alpha(r3, t_lam), r3_std(r3, t_lam)
    c_x_mid = a_l / a_t
    c_s_mid = s_l * s_t * (torch.exp(h / 2) - 1)
    c_x_fin = a_n / a_t
    c_s_fin = s_n * s_l * (torch.exp(h) - 1)
    sc_t, sc_l = so3_score_scaling(so3, t), so3_score_scaling(so3, t_lam)
    g_t, g_l = so3_g(so3, t), so3_g(so3, t_lam)
    dt_mid = t_lam - t
    out = []
    for i in range(num_steps):
        out.append(DpmStep(_f(t[i]), _f(t_lam[i]), DpmScalars(
            _f(s_t[i]), _f(c_x_mid[i]), _f(c_s_mid[i]), _f(s_l[i]), _f(c_x_fin[i]), _f(c_s_fin[i]), _f(sc_t[i]), _f(sc_l[i]),
            _f(g_t[i]), _f(g_l[i]), _f(dt_mid[i]), _f(dts[i]), float(so3.tol))))
    return out


@dataclass
class EmStep:
    t: float
    scalars: EmScalars


def em_scalars(r3, so3, t: torch.Tensor, dt: torch.Tensor, noise_weight: float = 1.0, mcf: float = 1.0) -> list[EmScalars]:
    """EulerMaruyamaPredictor constants (denoiser.py:54-97) for vectors of (t, dt)."""
    w = 0.5 * mcf * (1 + noise_weight**2)
    beta = r3.beta(t)
    sq = torch.sqrt(beta)
    std = r3_std(r3, t)
    g = so3_g(so3, t)
    sc = so3_score_scaling(so3, t)
    sdt = torch.sqrt(dt.abs())
    return [EmScalars(_f(dt[i]), _f(sdt[i]), float(noise_weight), float(w), _f(g[i]), _f(sc[i]), _f(beta[i]), _f(sq[i]),
                      _f(std[i]), float(so3.tol)) for i in range(t.numel())]


def em_schedule(r3, so3, num_steps: int, max_t: float, min_t: float) -> list[EmStep]:
    ts, dts = timesteps(max_t, min_t, num_steps)
    sc = em_scalars(r3, so3, ts[:-1].clone(), dts)
    return [EmStep(_f(ts[i]), sc[i]) for i in range(num_steps)]


@dataclass
class HeunStep:
    t: float
    t_hat: float
    t_next: float
    correct: bool
    scalars: HeunScalars
    em_at_t: EmScalars  # used by the fine-tune variant's trace-back (denoiser.py:598-607)


def heun_schedule(r3, so3, num_steps: int, max_t: float, min_t: float, noise: float) -> list[HeunStep]:
    """Per-step constants of heun_denoiser (denoiser.py:401-459)."""
    ts, dts = timesteps(max_t, min_t, num_steps)
    out = []
    for i in range(num_steps):
        t = torch.full((1,), ts[i].item())
        t_next = t + dts[i]
        t_hat = t - noise * dts[i] if (i > 0 and 0.0 < t[0] < 1.0) else t
        churn_dt = (t_hat - t)[0]
        step_dt = (t_next - t_hat)[0]

        def at(tt):
            b = r3.beta(tt)
            return so3_g(so3, tt), so3_score_scaling(so3, tt), b, torch.sqrt(b), r3_std(r3, tt)

        g0, _, b0, q0, _ = at(t)
        gh, sh, bh, qh, stdh = at(t_hat)
        gn, sn, bn, qn, stdn = at(t_next)
        sc = HeunScalars(_f(churn_dt), _f(torch.sqrt(churn_dt.abs())), _f(g0), _f(b0), _f(q0), _f(step_dt), _f(gh), _f(sh),
                         _f(bh), _f(qh), _f(stdh), _f(gn), _f(sn), _f(bn), _f(qn), _f(stdn), float(so3.tol))
        em = em_scalars(r3, so3, t, dts[i].reshape(1))[0]
        out.append(HeunStep(_f(t), _f(t_hat), _f(t_next), bool(t_next[0] > 0.0), sc, em))
    return out
