"""CPU ORACLE (test infrastructure) -- R3 VP-SDE, score conversion and the three samplers.

Restates bioemu/src/bioemu/sde_lib.py and bioemu/src/bioemu/denoiser.py on flat ``[N, .]`` state
(N = total residues, graphs concatenated) with a per-residue graph index ``batch_idx`` -- the same
sparse layout the reference works in, but without the ChemGraph container.  RNG consumption order
is the reference's: prior pos ``randn(N,3)``, prior rot ``randn(N,1,3)`` then
``rand(N,1)``, then per step per field (in ``fields`` order) one ``randn(N,3)`` -- drawn even
when it is multiplied by zero.

``score_fn(pos [N,3], rot [N,3,3], t [B]) -> (pos_out [N,3], rot_out [N,3])`` is the *raw* network
output; `get_score` converts it exactly like denoiser.py:169-203.
"""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, NamedTuple

import numpy as np
import torch

from .so3 import SO3Tables, apply_rotvec_to_rotmat, rotmat_to_rotvec


# --------------------------------------------------------------------------------------------
# Cosine VP-SDE                                                         sde_lib.py:105-167
# --------------------------------------------------------------------------------------------
class CosineVP:
    def __init__(self, s: float = 0.008):
        self.s = s
        self.c = np.cos(s / (1 + s) * np.pi / 2)

    def beta(self, t):  # sde_lib.py:160-162
        return torch.tan((t + self.s) / (1 + self.s) * np.pi / 2) * np.pi / (1 + self.s)

    def alpha(self, t):  # sde_lib.py:164-167
        return torch.clip(torch.cos((t + self.s) / (1 + self.s) * np.pi / 2) / self.c, 0, 1)

    def std(self, t):  # sde_lib.py:128
        return torch.sqrt(1.0 - self.alpha(t) ** 2)

    def t_from_lambda(self, lam):  # denoiser.py:623-631
        f = -1 / 2 * torch.log(torch.exp(-2 * lam) + 1)
        e = f + torch.log(torch.cos(torch.tensor(np.pi * self.s / 2 / (1 + self.s))))
        return 2 * (1 + self.s) / np.pi * torch.acos(torch.exp(e)) - self.s


class Path(NamedTuple):
    """DenoisedSDEPath (denoiser.py:23-27) with dense [T, B, L, 3] controls / increments."""

    pos: list
    rot: list
    timesteps: torch.Tensor
    us: dict
    dWs: dict


def _col(x, batch_idx, like):
    """maybe_expand (sde_lib.py:26-47): per-graph [B] -> per-residue, broadcast over trailing dims."""
    x = x[(...,) + (None,) * (like.ndim - x.ndim)]
    return x[batch_idx]


def _dense(x, batch_idx, lengths):
    """to_dense_batch for the stored controls (denoiser.py:334-335)."""
    b, lmax = len(lengths), max(lengths)
    ptr = np.concatenate([[0], np.cumsum(lengths)])
    out = x.new_zeros((b, lmax) + tuple(x.shape[1:]))
    for g in range(b):
        out[g, : lengths[g]] = x[ptr[g] : ptr[g + 1]]
    return out


def get_score(score_fn, pos, rot, t, batch_idx, r3: CosineVP, so3: SO3Tables):
    """denoiser.py:169-203."""
    p, r = score_fn(pos, rot, t)
    rot_score = r * so3.score_scaling_at(t)[batch_idx].unsqueeze(-1)
    pos_std = _col(torch.sqrt(1.0 - r3.alpha(t) ** 2), batch_idx, p)
    return p / pos_std, rot_score


class EM:
    """EulerMaruyamaPredictor (denoiser.py:30-166) for field kind 'pos' | 'rot'."""

    def __init__(self, kind, r3, so3, noise_weight=1.0, mcf=1.0):
        self.kind, self.r3, self.so3, self.nw, self.mcf = kind, r3, so3, noise_weight, mcf

    def sde(self, x, t, bi):
        if self.kind == "pos":  # sde_lib.py:140-150
            beta = self.r3.beta(t)
            return -0.5 * _col(beta, bi, x) * x, _col(torch.sqrt(beta), bi, x)
        drift = torch.zeros_like(x[..., 0])  # so3_sde.py:173-194
        return drift, _col(self.so3.beta(t), bi, drift)

    def drift_diffusion(self, x, t, score, bi, u=None):  # denoiser.py:54-70
        w = 0.5 * self.mcf * (1 + self.nw**2)
        drift, g = self.sde(x, t, bi)
        drift = drift - g**2 * score * w
        if u is not None:
            drift = drift + g * u * w
        return drift, g

    def update(self, x, dt, drift, g):  # denoiser.py:72-97
        z = torch.randn_like(drift)
        dW = self.nw * torch.sqrt(dt.abs()) * z
        if self.kind == "rot":
            mean = apply_rotvec_to_rotmat(x, drift * dt, tol=self.so3.tol)
            return apply_rotvec_to_rotmat(mean, g * dW, tol=self.so3.tol), mean, dW
        mean = x + drift * dt
        return mean + g * dW, mean, dW

    def step(self, x, t, dt, score, bi, u=None):  # denoiser.py:99-116
        d, g = self.drift_diffusion(x, t, score, bi, u)
        return self.update(x, dt, d, g)

    def forward_step(self, x, t, dt, bi):  # denoiser.py:118-131
        d, g = self.sde(x, t, bi)
        return self.update(x, dt, d, g)

    def traceback(self, x_next, x, t, dt, score, bi, u=None):  # denoiser.py:133-166
        d, g = self.drift_diffusion(x, t, score, bi, u)
        mean = self.update(x, dt, d, 0.0)[1]
        if self.kind == "rot":
            return rotmat_to_rotvec(torch.einsum("...ji,...jk->...ik", mean, x_next)) / g
        return (x_next - mean) / g


def _prior(n, r3, so3, fields):
    """denoiser.py:224-229: keyword evaluation order is pos first, then node_orientations."""
    pos = torch.randn(n, 3)
    rot = so3.prior(n)
    return pos, rot


FIELDS_YAML = ("node_orientations", "pos")  # config.yaml:23-38 key order -> per-step RNG order


def euler_maruyama(score_fn, lengths, r3, so3, num_steps, max_t, min_t, fields=FIELDS_YAML,
                   finetune_fn: Callable | None = None, init=None):
    """denoiser.py:206-348.  With ``finetune_fn`` returns a Path (the _finetune variant)."""
    b = len(lengths)
    bi = torch.repeat_interleave(torch.arange(b), torch.tensor(lengths))
    pos, rot = _prior(int(sum(lengths)), r3, so3, fields) if init is None else init
    ts = torch.linspace(max_t, min_t, num_steps + 1)
    dts = torch.diff(ts)
    em = {"pos": EM("pos", r3, so3, 1.0), "node_orientations": EM("rot", r3, so3, 1.0)}
    path_pos, path_rot, us, dWs = [pos], [rot], defaultdict(list), defaultdict(list)
    for i in range(num_steps):
        t = torch.full((b,), ts[i].item())
        s_pos, s_rot = get_score(score_fn, pos, rot, t, bi, r3, so3)
        score = {"pos": s_pos, "node_orientations": s_rot}
        x = {"pos": pos, "node_orientations": rot}
        u = None
        if finetune_fn is not None:
            u_pos, u_rot = finetune_fn(pos, rot, t)
            u = {"pos": u_pos, "node_orientations": u_rot}
        new = {}
        for f in fields:
            new[f], _, dW = em[f].step(x[f], t, dts[i], score[f], bi, None if u is None else u[f])
            if u is not None:
                us[f].append(_dense(u[f], bi, lengths))
                dWs[f].append(_dense(dW, bi, lengths))
        pos, rot = new["pos"], new["node_orientations"]
        path_pos.append(pos)
        path_rot.append(rot)
    if finetune_fn is None:
        return pos, rot
    return Path(path_pos, path_rot, ts, {f: torch.stack(us[f]) for f in fields},
                {f: torch.stack(dWs[f]) for f in fields})


def heun(score_fn, lengths, r3, so3, num_steps, max_t, min_t, noise, fields=FIELDS_YAML, init=None):
    """denoiser.py:351-461."""
    b = len(lengths)
    bi = torch.repeat_interleave(torch.arange(b), torch.tensor(lengths))
    pos, rot = _prior(int(sum(lengths)), r3, so3, fields) if init is None else init
    ts = torch.linspace(max_t, min_t, num_steps + 1)
    dts = torch.diff(ts)
    kinds = {"pos": "pos", "node_orientations": "rot"}
    pred = {f: EM(kinds[f], r3, so3, 0.0) for f in fields}
    nois = {f: EM(kinds[f], r3, so3, 1.0) for f in fields}
    x = {"pos": pos, "node_orientations": rot}
    for i in range(num_steps):
        t = torch.full((b,), ts[i].item())
        t_next = t + dts[i]
        t_hat = t - noise * dts[i] if (i > 0 and 0.0 < t[0] < 1.0) else t
        xh = {f: nois[f].forward_step(x[f], t, (t_hat - t)[0], bi)[0] for f in fields}
        s_pos, s_rot = get_score(score_fn, xh["pos"], xh["node_orientations"], t_hat, bi, r3, so3)
        sc = {"pos": s_pos, "node_orientations": s_rot}
        dh = {f: pred[f].drift_diffusion(xh[f], t_hat, sc[f], bi)[0] for f in fields}
        x = dict(x)
        for f in fields:
            x[f] = pred[f].update(xh[f], (t_next - t_hat)[0], dh[f], 0.0)[1]
        if t_next[0] > 0.0:
            s_pos, s_rot = get_score(score_fn, x["pos"], x["node_orientations"], t_next, bi, r3, so3)
            sn = {"pos": s_pos, "node_orientations": s_rot}
            avg = {}
            for f in fields:
                dn = pred[f].drift_diffusion(x[f], t_next, sn[f], bi)[0]
                avg[f] = (dn + dh[f]) / 2
            for f in fields:
                x[f] = pred[f].update(xh[f], (t_next - t_hat)[0], avg[f], 0.0)[1]
    return x["pos"], x["node_orientations"]


def heun_finetune(score_fn, finetune_fn, lengths, r3, so3, num_steps, max_t, min_t, noise, fields=FIELDS_YAML):
    """heun_denoiser_finetune (denoiser.py:462-620): the Heun step with the control in every drift, plus the Brownian
    increment of the equivalent Euler-Maruyama step traced back per field.  The reference appends ONE in-place-mutated
    batch object to `batches` every step (denoiser.py:518,564,588,596), so all its stored batches alias the final state;
    this restatement returns per-step snapshots and the tests compare the last one."""
    b = len(lengths)
    bi = torch.repeat_interleave(torch.arange(b), torch.tensor(lengths))
    pos, rot = _prior(int(sum(lengths)), r3, so3, fields)
    ts = torch.linspace(max_t, min_t, num_steps + 1)
    dts = torch.diff(ts)
    kinds = {"pos": "pos", "node_orientations": "rot"}
    pred = {f: EM(kinds[f], r3, so3, 0.0) for f in fields}
    nois = {f: EM(kinds[f], r3, so3, 1.0) for f in fields}
    x = {"pos": pos, "node_orientations": rot}
    path_pos, path_rot, us, dWs = [pos], [rot], defaultdict(list), defaultdict(list)

    def both(fn, xx, t):
        a, c = fn(xx["pos"], xx["node_orientations"], t)
        return {"pos": a, "node_orientations": c}

    for i in range(num_steps):
        t = torch.full((b,), ts[i].item())
        t_next = t + dts[i]
        churn = i > 0 and 0.0 < t[0] < 1.0
        t_hat = t - noise * dts[i] if churn else t
        xh = {f: nois[f].forward_step(x[f], t, (t_hat - t)[0], bi)[0] for f in fields}
        sc_h = both(lambda p, r, tt: get_score(score_fn, p, r, tt, bi, r3, so3), xh, t_hat)
        u_h = both(finetune_fn, xh, t_hat)
        x_prev = dict(x)
        if churn:
            sc = both(lambda p, r, tt: get_score(score_fn, p, r, tt, bi, r3, so3), x, t)
            u = both(finetune_fn, x, t)
        else:
            sc, u = sc_h, u_h
        dh = {f: pred[f].drift_diffusion(xh[f], t_hat, sc_h[f], bi, u_h[f])[0] for f in fields}
        x = {}
        for f in fields:
            x[f] = pred[f].update(xh[f], (t_next - t_hat)[0], dh[f], 0.0)[1]
        if t_next[0] > 0.0:
            sc_n = both(lambda p, r, tt: get_score(score_fn, p, r, tt, bi, r3, so3), x, t_next)
            u_n = both(finetune_fn, x, t_next)
            avg = {}
            for f in fields:
                dn = pred[f].drift_diffusion(x[f], t_next, sc_n[f], bi, u_n[f])[0]
                avg[f] = (dn + dh[f]) / 2
            for f in fields:
                x[f] = pred[f].update(xh[f], (t_next - t_hat)[0], avg[f], 0.0)[1]
        path_pos.append(x["pos"])
        path_rot.append(x["node_orientations"])
        for f in fields:
            dW = nois[f].traceback(x[f], x_prev[f], t, dts[i], sc[f], bi, u[f])
            us[f].append(_dense(u[f], bi, lengths))
            dWs[f].append(_dense(dW, bi, lengths))
    return Path(path_pos, path_rot, ts, {f: torch.stack(us[f]) for f in fields}, {f: torch.stack(dWs[f]) for f in fields})


def dpm_solver(score_fn, lengths, r3, so3, num_steps, max_t, min_t, init=None, trace=None):
    """denoiser.py:634-764 (DPM-Solver-2 on pos, midpoint/extrapolated exp-map step on rot)."""
    assert max_t < 1.0
    b = len(lengths)
    bi = torch.repeat_interleave(torch.arange(b), torch.tensor(lengths))
    pos, rot = _prior(int(sum(lengths)), r3, so3, None) if init is None else init
    ts = torch.linspace(max_t, min_t, num_steps + 1)
    dts = torch.diff(ts)
    so3p = EM("rot", r3, so3, 0.0)
    for i in range(num_steps):
        t = torch.full((b,), ts[i].item())
        t_next = t + dts[i]
        s_pos, s_rot = get_score(score_fn, pos, rot, t, bi, r3, so3)

        a_t, sg_t = _col(r3.alpha(t), bi, pos), _col(r3.std(t), bi, pos)
        lam = torch.log(a_t / sg_t)
        a_n, sg_n = _col(r3.alpha(t_next), bi, pos), _col(r3.std(t_next), bi, pos)
        lam_n = torch.log(a_n / sg_n)
        h = lam_n - lam
        t_lam = r3.t_from_lambda((lam + lam_n) / 2)
        t_lam = torch.full((b,), t_lam[0][0].item())
        a_l, sg_l = _col(r3.alpha(t_lam), bi, pos), _col(r3.std(t_lam), bi, pos)

        u = a_l / a_t * pos + sg_l * sg_t * (torch.exp(h / 2) - 1) * s_pos
        drift, _ = so3p.drift_diffusion(rot, t, s_rot, bi)
        rot_u = so3p.update(rot, (t_lam - t)[0], drift, 0.0)[1]

        su_pos, su_rot = get_score(score_fn, u, rot_u, t_lam, bi, r3, so3)
        pos_next = a_n / a_t * pos + sg_n * sg_l * (torch.exp(h) - 1) * su_pos
        node_score = su_rot + 0.5 * (su_rot - s_rot) / (t_lam - t)[0] * dts[i]
        drift, _ = so3p.drift_diffusion(rot_u, t_lam, node_score, bi)
        rot_next = so3p.update(rot, dts[i], drift, 0.0)[1]
        if trace is not None:
            trace.append(dict(t=t[0].item(), t_next=t_next[0].item(), t_lam=t_lam[0].item(),
                              alpha_t=a_t[0, 0].item(), sigma_t=sg_t[0, 0].item(),
                              lam=lam[0, 0].item(), h=h[0, 0].item(),
                              alpha_l=a_l[0, 0].item(), sigma_l=sg_l[0, 0].item(),
                              alpha_n=a_n[0, 0].item(), sigma_n=sg_n[0, 0].item(),
                              u=u, rot_u=rot_u, pos=pos_next, rot=rot_next))
        pos, rot = pos_next, rot_next
    return pos, rot
