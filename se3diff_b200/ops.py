"""torch-tensor front end of the C ABI (include/se3diff_b200.h).

torch is used here only for device memory and the current CUDA stream; every function hands raw
device pointers to libse3diff_b200.so.  CPU tensors are rejected: there is no CPU implementation.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

__all__ = [
    "so3_exp", "so3_log", "so3_angle", "so3_compose_rotvec", "so3_matmul", "so3_rel_log", "so3_geodesic",
    "so3_from_quat", "frame_update_em", "frame_update_dpm_mid", "frame_update_dpm_final", "frame_heun_churn",
    "frame_heun_predict", "frame_heun_correct", "frame_traceback", "igso3_series", "igso3_score",
    "igso3_marginal_pdf", "igso3_build_cdf", "igso3_build_score_scaling", "igso3_sample", "ipa_attention_fwd",
    "launch_count", "launch_count_reset",
]


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _dev(t: torch.Tensor, dtype=torch.float32, name="tensor") -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise L.Se3LibraryError(f"{name}: se3diff_b200 ops run on CUDA tensors only (got {t.device}); "
                                "there is no CPU fallback")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _guard(t):
    return torch.cuda.device(t.device)


_replayed_launches = 0


def launch_count() -> int:
    """Kernels of libse3diff_b200 launched by this thread (direct launches + launches replayed through CUDA graphs)."""
    return int(L.lib().se3_launch_count()) + _replayed_launches


def launch_count_reset() -> None:
    global _replayed_launches
    _replayed_launches = 0
    L.lib().se3_launch_count_reset()


def count_replayed_launches(n: int) -> None:
    global _replayed_launches
    _replayed_launches += int(n)


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
def so3_exp(rotvec: torch.Tensor, tol: float = 1e-7) -> torch.Tensor:
    """rotvec_to_rotmat (so3_sde.py:533-554).  [...,3] -> [...,3,3]; fp32 or fp64."""
    dt = torch.float64 if rotvec.dtype == torch.float64 else torch.float32
    v = _dev(rotvec, dt, "rotvec")
    out = torch.empty(v.shape + (3,), dtype=dt, device=v.device)
    n = v.numel() // 3
    with _guard(v):
        fn = L.lib().se3_so3_exp_f64 if dt == torch.float64 else L.lib().se3_so3_exp
        L.check(fn(_p(v), _p(out), n, tol, _stream(v)), "se3_so3_exp")
    return out


def so3_log(rotmat: torch.Tensor) -> torch.Tensor:
    """rotmat_to_rotvec (so3_sde.py:557-648).  [...,3,3] -> [...,3]."""
    dt = torch.float64 if rotmat.dtype == torch.float64 else torch.float32
    r = _dev(rotmat, dt, "rotmat")
    out = torch.empty(r.shape[:-1], dtype=dt, device=r.device)
    with _guard(r):
        fn = L.lib().se3_so3_log_f64 if dt == torch.float64 else L.lib().se3_so3_log
        L.check(fn(_p(r), _p(out), r.numel() // 9, _stream(r)), "se3_so3_log")
    return out


def so3_angle(rotmat: torch.Tensor):
    """angle_from_rotmat (so3_sde.py:651-676) -> (angle, sin, cos)."""
    r = _dev(rotmat, name="rotmat")
    a, s, c = (torch.empty(r.shape[:-2], dtype=torch.float32, device=r.device) for _ in range(3))
    with _guard(r):
        L.check(L.lib().se3_so3_angle(_p(r), _p(a), _p(s), _p(c), r.numel() // 9, _stream(r)), "se3_so3_angle")
    return a, s, c


def so3_compose_rotvec(rotmat: torch.Tensor, rotvec: torch.Tensor, tol: float = 1e-7, out=None) -> torch.Tensor:
    """apply_rotvec_to_rotmat (so3_sde.py:782-802): R . Exp(v)."""
    r, v = _dev(rotmat, name="rotmat"), _dev(rotvec, name="rotvec")
    if r.shape[:-2] != v.shape[:-1]:
        raise ValueError(f"shape mismatch {tuple(r.shape)} vs {tuple(v.shape)}")
    out = torch.empty_like(r) if out is None else out
    with _guard(r):
        L.check(L.lib().se3_so3_compose_rotvec(_p(r), _p(v), _p(out), r.numel() // 9, tol, _stream(r)),
                "se3_so3_compose_rotvec")
    return out


def so3_matmul(a: torch.Tensor, b: torch.Tensor, transpose_a: bool = False) -> torch.Tensor:
    """rot_mult / rot_transpose (so3_sde.py:870-877)."""
    a, b = _dev(a, name="a"), _dev(b, name="b")
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    out = torch.empty_like(a)
    with _guard(a):
        L.check(L.lib().se3_so3_matmul(_p(a), _p(b), _p(out), a.numel() // 9, int(transpose_a), _stream(a)),
                "se3_so3_matmul")
    return out


def so3_rel_log(base: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """rot_vf (so3_sde.py:880-891): Log(base^T target)."""
    a, b = _dev(base, name="base"), _dev(target, name="target")
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    out = torch.empty(a.shape[:-1], dtype=torch.float32, device=a.device)
    with _guard(a):
        L.check(L.lib().se3_so3_rel_log(_p(a), _p(b), _p(out), a.numel() // 9, _stream(a)), "se3_so3_rel_log")
    return out


def so3_geodesic(base: torch.Tensor, target: torch.Tensor, t: float, tol: float = 1e-7) -> torch.Tensor:
    """geodesic_t (so3_sde.py:894-911): base . Exp(t Log(base^T target))."""
    a, b = _dev(base, name="base"), _dev(target, name="target")
    if a.shape != b.shape:
        raise ValueError(f"Incompatible shapes: base_mat={tuple(a.shape)}, mat_t={tuple(b.shape)}")
    out = torch.empty_like(a)
    with _guard(a):
        L.check(L.lib().se3_so3_geodesic(_p(a), _p(b), float(t), _p(out), a.numel() // 9, tol, _stream(a)),
                "se3_so3_geodesic")
    return out


def so3_from_quat(quat: torch.Tensor, want_rotvec=True, want_rotmat=True, tol: float = 1e-7):
    """rotquat_to_rotvec / rotquat_to_rotmat (so3_sde.py:725-779); quaternion [r,i,j,k]."""
    q = _dev(quat, name="quat")
    rv = torch.empty(q.shape[:-1] + (3,), dtype=torch.float32, device=q.device) if want_rotvec else None
    rm = torch.empty(q.shape[:-1] + (3, 3), dtype=torch.float32, device=q.device) if want_rotmat else None
    with _guard(q):
        L.check(L.lib().se3_so3_from_quat(_p(q), _p(rv), _p(rm), q.numel() // 4, tol, _stream(q)), "se3_so3_from_quat")
    return rv, rm


# ------------------------------------------------------------------------------------------------
# K3 + frame update
# ------------------------------------------------------------------------------------------------
def _chk_frames(rot, pos):
    rot, pos = _dev(rot, name="rot"), _dev(pos, name="pos")
    n = pos.numel() // 3
    if rot.numel() != 9 * n:
        raise ValueError(f"rot {tuple(rot.shape)} and pos {tuple(pos.shape)} disagree on the residue count")
    return rot, pos, n


def _vec(t, n, name):
    if t is None:
        return None
    t = _dev(t, name=name)
    if t.numel() != 3 * n:
        raise ValueError(f"{name}: expected {n}x3 values, got {tuple(t.shape)}")
    return t


def frame_update_em(rot, pos, m_rot, m_pos, z_rot, z_pos, scalars: L.EmScalars, u_rot=None, u_pos=None,
                    want_dw=False, rot_out=None, pos_out=None):
    """One Euler-Maruyama reverse step on both fields (denoiser.py:54-116)."""
    rot, pos, n = _chk_frames(rot, pos)
    m_rot, m_pos, z_rot, z_pos = (_vec(t, n, k) for t, k in ((m_rot, "m_rot"), (m_pos, "m_pos"), (z_rot, "z_rot"), (z_pos, "z_pos")))
    u_rot, u_pos = _vec(u_rot, n, "u_rot"), _vec(u_pos, n, "u_pos")
    rot_out = torch.empty_like(rot) if rot_out is None else rot_out
    pos_out = torch.empty_like(pos) if pos_out is None else pos_out
    dw_rot = torch.empty_like(pos) if want_dw else None
    dw_pos = torch.empty_like(pos) if want_dw else None
    with _guard(rot):
        L.check(L.lib().se3_frame_update_em(_p(rot), _p(pos), _p(m_rot), _p(m_pos), _p(u_rot), _p(u_pos), _p(z_rot),
                                            _p(z_pos), _p(rot_out), _p(pos_out), _p(dw_rot), _p(dw_pos), n,
                                            C.byref(scalars), _stream(rot)), "se3_frame_update_em")
    return rot_out, pos_out, dw_rot, dw_pos


def so3_update_em(rot, m_rot, z_rot, scalars: L.EmScalars, u_rot=None, want_dw=False, rot_out=None):
    """The rotation half of the Euler-Maruyama step alone, on bare [n,3,3] rotations (se3diff/train.py:54-70)."""
    rot = _dev(rot, name="rot")
    n = rot.numel() // 9
    m_rot, z_rot, u_rot = _vec(m_rot, n, "m_rot"), _vec(z_rot, n, "z_rot"), _vec(u_rot, n, "u_rot")
    rot_out = torch.empty_like(rot) if rot_out is None else rot_out
    dw = torch.empty(n, 3, dtype=torch.float32, device=rot.device) if want_dw else None
    with _guard(rot):
        L.check(L.lib().se3_so3_update_em(_p(rot), _p(m_rot), _p(u_rot), _p(z_rot), _p(rot_out), _p(dw), n, C.byref(scalars),
                                          _stream(rot)), "se3_so3_update_em")
    return rot_out, dw


def _pos(x, name="pos"):
    x = _dev(x, name=name)
    if x.shape[-1] != 3:
        raise ValueError(f"{name}: expected [..., 3], got {tuple(x.shape)}")
    return x, x.numel() // 3


def r3_update_em(pos, m_pos, z_pos, scalars: L.EmScalars, u_pos=None, want_dw=False, pos_out=None):
    """The translation half of the Euler-Maruyama step alone (EulerMaruyamaPredictor with a CosineVPSDE corruption,
    denoiser.py:54-97): returns (pos_out, dW or None).  Bit-identical to the pos outputs of `frame_update_em`."""
    pos, n = _pos(pos)
    m_pos, z_pos, u_pos = _vec(m_pos, n, "m_pos"), _vec(z_pos, n, "z_pos"), _vec(u_pos, n, "u_pos")
    pos_out = torch.empty_like(pos) if pos_out is None else pos_out
    dw = torch.empty_like(pos) if want_dw else None
    with _guard(pos):
        L.check(L.lib().se3_r3_update_em(_p(pos), _p(m_pos), _p(u_pos), _p(z_pos), _p(pos_out), _p(dw), n, C.byref(scalars),
                                         _stream(pos)), "se3_r3_update_em")
    return pos_out, dw


def r3_update_dpm(pos, m_pos, scalars: L.DpmScalars, final_half: bool, pos_out=None):
    """Position line of a DPM-Solver-2 half step (denoiser.py:699-701 / 733-735)."""
    pos, n = _pos(pos)
    m_pos = _vec(m_pos, n, "m_pos")
    pos_out = torch.empty_like(pos) if pos_out is None else pos_out
    with _guard(pos):
        L.check(L.lib().se3_r3_update_dpm(_p(pos), _p(m_pos), _p(pos_out), n, C.byref(scalars), int(bool(final_half)), _stream(pos)),
                "se3_r3_update_dpm")
    return pos_out


def r3_heun_churn(pos, z_pos, scalars: L.HeunScalars, pos_out=None):
    pos, n = _pos(pos)
    z_pos = _vec(z_pos, n, "z_pos")
    pos_out = torch.empty_like(pos) if pos_out is None else pos_out
    with _guard(pos):
        L.check(L.lib().se3_r3_heun_churn(_p(pos), _p(z_pos), _p(pos_out), n, C.byref(scalars), _stream(pos)), "se3_r3_heun_churn")
    return pos_out


def r3_heun_step(pos_hat, m_pos_hat, scalars: L.HeunScalars, pos_pred=None, m_pos_next=None, pos_out=None):
    """Heun position step from pos_hat: first order, or corrected when (pos_pred, m_pos_next) are given (denoiser.py:423-459)."""
    pos_hat, n = _pos(pos_hat, "pos_hat")
    m_pos_hat = _vec(m_pos_hat, n, "m_pos_hat")
    pos_pred, m_pos_next = _vec(pos_pred, n, "pos_pred"), _vec(m_pos_next, n, "m_pos_next")
    pos_out = torch.empty_like(pos_hat) if pos_out is None else pos_out
    with _guard(pos_hat):
        L.check(L.lib().se3_r3_heun_step(_p(pos_hat), _p(m_pos_hat), _p(pos_pred), _p(m_pos_next), _p(pos_out), n, C.byref(scalars),
                                         _stream(pos_hat)), "se3_r3_heun_step")
    return pos_out


def frame_update_dpm_mid(rot, pos, m_rot, m_pos, scalars: L.DpmScalars, rot_out=None, pos_out=None):
    rot, pos, n = _chk_frames(rot, pos)
    m_rot, m_pos = _vec(m_rot, n, "m_rot"), _vec(m_pos, n, "m_pos")
    rot_out = torch.empty_like(rot) if rot_out is None else rot_out
    pos_out = torch.empty_like(pos) if pos_out is None else pos_out
    with _guard(rot):
        L.check(L.lib().se3_frame_update_dpm_mid(_p(rot), _p(pos), _p(m_rot), _p(m_pos), _p(rot_out), _p(pos_out), n,
                                                 C.byref(scalars), _stream(rot)), "se3_frame_update_dpm_mid")
    return rot_out, pos_out


def frame_update_dpm_final(rot, pos, m_rot_t, m_rot_lam, m_pos_lam, scalars: L.DpmScalars, rot_out=None, pos_out=None):
    rot, pos, n = _chk_frames(rot, pos)
    m_rot_t, m_rot_lam, m_pos_lam = _vec(m_rot_t, n, "m_rot_t"), _vec(m_rot_lam, n, "m_rot_lam"), _vec(m_pos_lam, n, "m_pos_lam")
    rot_out = torch.empty_like(rot) if rot_out is None else rot_out
    pos_out = torch.empty_like(pos) if pos_out is None else pos_out
    with _guard(rot):
        L.check(L.lib().se3_frame_update_dpm_final(_p(rot), _p(pos), _p(m_rot_t), _p(m_rot_lam), _p(m_pos_lam),
                                                   _p(rot_out), _p(pos_out), n, C.byref(scalars), _stream(rot)),
                "se3_frame_update_dpm_final")
    return rot_out, pos_out


def frame_heun_churn(rot, pos, z_rot, z_pos, scalars: L.HeunScalars):
    rot, pos, n = _chk_frames(rot, pos)
    z_rot, z_pos = _vec(z_rot, n, "z_rot"), _vec(z_pos, n, "z_pos")
    rot_out, pos_out = torch.empty_like(rot), torch.empty_like(pos)
    with _guard(rot):
        L.check(L.lib().se3_frame_heun_churn(_p(rot), _p(pos), _p(z_rot), _p(z_pos), _p(rot_out), _p(pos_out), n,
                                             C.byref(scalars), _stream(rot)), "se3_frame_heun_churn")
    return rot_out, pos_out


def frame_heun_predict(rot_hat, pos_hat, m_rot_hat, m_pos_hat, scalars: L.HeunScalars):
    rot_hat, pos_hat, n = _chk_frames(rot_hat, pos_hat)
    m_rot_hat, m_pos_hat = _vec(m_rot_hat, n, "m_rot_hat"), _vec(m_pos_hat, n, "m_pos_hat")
    rot_out, pos_out = torch.empty_like(rot_hat), torch.empty_like(pos_hat)
    with _guard(rot_hat):
        L.check(L.lib().se3_frame_heun_predict(_p(rot_hat), _p(pos_hat), _p(m_rot_hat), _p(m_pos_hat), _p(rot_out),
                                               _p(pos_out), n, C.byref(scalars), _stream(rot_hat)), "se3_frame_heun_predict")
    return rot_out, pos_out


def frame_heun_correct(rot_hat, pos_hat, m_rot_hat, m_pos_hat, pos_pred, m_rot_next, m_pos_next, scalars: L.HeunScalars):
    rot_hat, pos_hat, n = _chk_frames(rot_hat, pos_hat)
    m_rot_hat, m_pos_hat, pos_pred, m_rot_next, m_pos_next = (
        _vec(t, n, k) for t, k in ((m_rot_hat, "m_rot_hat"), (m_pos_hat, "m_pos_hat"), (pos_pred, "pos_pred"),
                                   (m_rot_next, "m_rot_next"), (m_pos_next, "m_pos_next")))
    rot_out, pos_out = torch.empty_like(rot_hat), torch.empty_like(pos_hat)
    with _guard(rot_hat):
        L.check(L.lib().se3_frame_heun_correct(_p(rot_hat), _p(pos_hat), _p(m_rot_hat), _p(m_pos_hat), _p(pos_pred),
                                               _p(m_rot_next), _p(m_pos_next), _p(rot_out), _p(pos_out), n,
                                               C.byref(scalars), _stream(rot_hat)), "se3_frame_heun_correct")
    return rot_out, pos_out


def frame_traceback(rot, pos, rot_next, pos_next, m_rot, m_pos, scalars: L.EmScalars, u_rot=None, u_pos=None):
    """traceback_brownian_motion (denoiser.py:133-166) for both fields -> (dW_rot, dW_pos)."""
    rot, pos, n = _chk_frames(rot, pos)
    rot_next, pos_next, _ = _chk_frames(rot_next, pos_next)
    m_rot, m_pos = _vec(m_rot, n, "m_rot"), _vec(m_pos, n, "m_pos")
    u_rot, u_pos = _vec(u_rot, n, "u_rot"), _vec(u_pos, n, "u_pos")
    dw_rot, dw_pos = torch.empty_like(pos), torch.empty_like(pos)
    with _guard(rot):
        L.check(L.lib().se3_frame_traceback(_p(rot), _p(pos), _p(rot_next), _p(pos_next), _p(m_rot), _p(m_pos),
                                            _p(u_rot), _p(u_pos), _p(dw_rot), _p(dw_pos), n, C.byref(scalars),
                                            _stream(rot)), "se3_frame_traceback")
    return dw_rot, dw_pos


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
def igso3_series(omega: torch.Tensor, sigma: torch.Tensor, l_max: int, tol: float = 1e-7, want=("f", "df", "dlog")):
    """igso3_expansion / digso3_expansion / dlog_igso3_expansion (so3_sde.py:1731-1940), l = 0..l_max."""
    dt = torch.float64 if omega.dtype == torch.float64 else torch.float32
    omega, sigma = torch.broadcast_tensors(omega, sigma)
    o, s = _dev(omega, dt, "omega"), _dev(sigma, dt, "sigma")
    outs = {k: (torch.empty_like(o) if k in want else None) for k in ("f", "df", "dlog")}
    with _guard(o):
        fn = L.lib().se3_igso3_series_f64 if dt == torch.float64 else L.lib().se3_igso3_series_f32
        L.check(fn(_p(o), _p(s), _p(outs["f"]), _p(outs["df"]), _p(outs["dlog"]), o.numel(), int(l_max), tol, _stream(o)),
                "se3_igso3_series")
    return outs


def igso3_score(rotvec: torch.Tensor, sigma: torch.Tensor, l_max: int, tol: float = 1e-7) -> torch.Tensor:
    """ScoreSO3.forward (so3_sde.py:1698-1715)."""
    q = _dev(rotvec, name="rotvec")
    s = _dev(sigma.expand(q.shape[:-1]) if sigma.shape != q.shape[:-1] else sigma, name="sigma")
    out = torch.empty_like(q)
    with _guard(q):
        L.check(L.lib().se3_igso3_score(_p(q), _p(s), _p(out), q.numel() // 3, int(l_max), tol, _stream(q)), "se3_igso3_score")
    return out


def igso3_marginal_pdf(omega, omega_0, sigma, l_count: int, tol: float = 1e-7) -> torch.Tensor:
    """igso3_marginal_pdf (so3_sde.py:1795-1854) with l = 0..l_count-1; inputs broadcast together."""
    omega, omega_0, sigma = torch.broadcast_tensors(omega, omega_0, sigma)
    o, o0, s = _dev(omega, name="omega"), _dev(omega_0, name="omega_0"), _dev(sigma, name="sigma")
    out = torch.empty_like(o)
    with _guard(o):
        L.check(L.lib().se3_igso3_marginal_pdf(_p(o), _p(o0), _p(s), _p(out), o.numel(), int(l_count), tol, _stream(o)),
                "se3_igso3_marginal_pdf")
    return out


def igso3_build_cdf(sigma_grid: torch.Tensor, omega_pts: torch.Tensor, l_max: int, tol: float = 1e-7, uniform=False):
    sg = _dev(sigma_grid, name="sigma_grid")
    om = _dev(omega_pts, torch.float64, "omega_pts")
    rows = 1 if uniform else sg.numel()
    cdf = torch.empty(rows, om.numel() - 1, dtype=torch.float32, device=sg.device)
    with _guard(sg):
        L.check(L.lib().se3_igso3_build_cdf(_p(sg), sg.numel(), _p(om), om.numel(), int(l_max), tol, int(uniform), _p(cdf),
                                            _stream(sg)), "se3_igso3_build_cdf")
    return cdf


def igso3_build_score_scaling(sigma_grid: torch.Tensor, omega_pts: torch.Tensor, l_max: int, tol: float = 1e-7):
    sg = _dev(sigma_grid, name="sigma_grid")
    om = _dev(omega_pts, torch.float64, "omega_pts")
    out = torch.empty(sg.numel(), dtype=torch.float32, device=sg.device)
    with _guard(sg):
        L.check(L.lib().se3_igso3_build_score_scaling(_p(sg), sg.numel(), _p(om), om.numel(), int(l_max), tol, _p(out),
                                                      _stream(sg)), "se3_igso3_build_score_scaling")
    return out


def igso3_build_cdf_index(cdf: torch.Tensor) -> torch.Tensor:
    """Guide records over the CDF rows (one 32-byte record per (row, bin of [0,1)): a lookup is one L2 sector); flat fp32."""
    cdf = _dev(cdf, name="cdf")
    rows, n = cdf.shape
    count = int(L.lib().se3_igso3_cdf_index_floats(rows, n))
    if count < 0:
        raise ValueError(f"igso3_build_cdf_index: unsupported table shape {tuple(cdf.shape)}")
    out = torch.empty(count, dtype=torch.float32, device=cdf.device)
    with _guard(cdf):
        L.check(L.lib().se3_igso3_build_cdf_index(_p(cdf), rows, n, _p(out), _stream(cdf)), "se3_igso3_build_cdf_index")
    return out


def igso3_sample(cdf, omega_grid, n: int, sigma=None, sigma_grid=None, normals=None, u=None, seed: int = 0, x=None,
                 tol: float = 1e-7, want_angle=False, cdf_index=None):
    """BaseSampleSO3.sample with one sample per element (so3_sde.py:1189-1286) [+ x . r]."""
    cdf, og = _dev(cdf, name="cdf"), _dev(omega_grid, name="omega_grid")
    sigma = None if sigma is None else _dev(sigma, name="sigma")
    sigma_grid = None if sigma_grid is None else _dev(sigma_grid, name="sigma_grid")
    normals = None if normals is None else _dev(normals, name="normals")
    u = None if u is None else _dev(u, name="u")
    x = None if x is None else _dev(x, name="x")
    cdf_index = None if cdf_index is None else _dev(cdf_index, name="cdf_index")
    out = torch.empty(n, 3, 3, dtype=torch.float32, device=cdf.device)
    ang = torch.empty(n, dtype=torch.float32, device=cdf.device) if want_angle else None
    with _guard(cdf):
        L.check(L.lib().se3_igso3_sample(_p(sigma), _p(sigma_grid), 0 if sigma_grid is None else sigma_grid.numel(), _p(cdf),
                                         _p(og), og.numel(), _p(normals), _p(u), int(seed) & (2**64 - 1), _p(x), _p(out),
                                         _p(ang), n, tol, _p(cdf_index), _stream(cdf)), "se3_igso3_sample")
    return (out, ang) if want_angle else out


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------
IPA_EXACT, IPA_FAST_MATH = 0, 1


def ipa_attention_fwd(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight: float,
                      shape: L.IpaShape, flags: int = IPA_EXACT, out=None):
    """SAAttention.forward between the projections and fc_out (structure_module.py:131-216)."""
    proj, rot, trans = _dev(proj, name="proj"), _dev(rot, name="rot"), _dev(trans, name="trans")
    pair_bias, pair_value = _dev(pair_bias, name="pair_bias"), _dev(pair_value, name="pair_value")
    key_bias = None if key_bias is None else _dev(key_bias, name="key_bias")
    head_weight = _dev(head_weight, name="head_weight")
    width = shape.heads * (2 * shape.dk + 4 * shape.pv)
    if out is None:
        out = torch.empty(shape.batch * shape.len, width, dtype=torch.float32, device=proj.device)
    with _guard(proj):
        L.check(L.lib().se3_ipa_attention_fwd(_p(proj), _p(rot), _p(trans), _p(pair_bias), _p(pair_value), _p(key_bias),
                                              _p(head_weight), float(scalar_weight), _p(out), C.byref(shape), int(flags),
                                              _stream(proj)), "se3_ipa_attention_fwd")
    return out


def ipa_bwd_supported(shape: L.IpaShape) -> bool:
    """Shapes se3_ipa_attention_bwd takes: L <= 128 with everything of a (sample, head) in shared memory, or -- up to L = 512 --
    the tiled two-kernel edition, which parks the logits of 64 query rows in shared memory."""
    n, kw = shape.len, 2 * shape.dk + 36
    tiled = (64 * kw + 64 * (n | 1) + 2 * 64 * 65 + n) * 4      # the larger of the two key-chunk plans of k_ipa_bwd_rows
    return (shape.dk in (4, 8, 16, 32) and shape.pq == 4 and shape.pv == 8 and 0 < n <= 512 and 0 < shape.batch <= 65535
            and tiled <= 227 * 1024)


def ipa_attention_bwd(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight: float, out, d_out,
                      shape: L.IpaShape):
    """Gradient of `ipa_attention_fwd` (exact mode): returns (d_proj [B*L, stride], P [B,H,L,L], dS [B,H,L,L], d_hw_rows [B*L,H])."""
    proj, rot, trans = _dev(proj, name="proj"), _dev(rot, name="rot"), _dev(trans, name="trans")
    pair_bias, pair_value = _dev(pair_bias, name="pair_bias"), _dev(pair_value, name="pair_value")
    key_bias = None if key_bias is None else _dev(key_bias, name="key_bias")
    head_weight, out, d_out = _dev(head_weight, name="head_weight"), _dev(out, name="out"), _dev(d_out, name="d_out")
    B, n, H = shape.batch, shape.len, shape.heads
    d_proj = torch.empty_like(proj)
    p_ws = torch.empty(B, H, n, n, dtype=torch.float32, device=proj.device)
    ds_ws = torch.empty_like(p_ws)
    d_hw_rows = torch.empty(B * n, H, dtype=torch.float32, device=proj.device)
    with _guard(proj):
        L.check(L.lib().se3_ipa_attention_bwd(_p(proj), _p(rot), _p(trans), _p(pair_bias), _p(pair_value), _p(key_bias),
                                              _p(head_weight), float(scalar_weight), _p(out), _p(d_out), _p(d_proj), _p(p_ws),
                                              _p(ds_ws), _p(d_hw_rows), C.byref(shape), _stream(proj)), "se3_ipa_attention_bwd")
    return d_proj, p_ws, ds_ws, d_hw_rows


class IpaAttention(torch.autograd.Function):
    """`ipa_attention_fwd` (exact fp32) as a differentiable operator: gradients for the projection, the pair bias, the pair
    values and the head weights from `se3_ipa_attention_bwd`; the frames are constants (stored rollout states)."""

    @staticmethod
    def forward(ctx, proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, shape):
        if rot.requires_grad or trans.requires_grad:
            raise L.Se3LibraryError("IpaAttention: frames with requires_grad are not supported (no gradient is produced for them)")
        proj, pair_bias, pair_value = proj.contiguous(), pair_bias.contiguous(), pair_value.contiguous()
        out = ipa_attention_fwd(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, scalar_weight, shape, IPA_EXACT)
        ctx.save_for_backward(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, out)
        ctx.scalar_weight, ctx.shape = float(scalar_weight), shape
        return out

    @staticmethod
    def backward(ctx, d_out):
        proj, rot, trans, pair_bias, pair_value, key_bias, head_weight, out = ctx.saved_tensors
        sh = ctx.shape
        B, n, H, dk = sh.batch, sh.len, sh.heads, sh.dk
        d_out = d_out.contiguous()
        d_proj, P, dS, d_hw_rows = ipa_attention_bwd(proj, rot, trans, pair_bias, pair_value, key_bias, head_weight,
                                                      ctx.scalar_weight, out, d_out, sh)
        hd, c_z = H * dk, H * dk + 3 * H * sh.pv
        d_bias = d_pv = None
        if ctx.needs_input_grad[3]:
            d_bias = dS.sum(0, keepdim=True) if sh.pair_batch == 1 else dS
        if ctx.needs_input_grad[4]:
            gz = d_out[:, c_z:c_z + hd].reshape(B, n, H, dk)
            if sh.pair_batch == 1:                                      # GEMM over the samples, per (head, query)
                d_pv = torch.einsum("bhij,bihc->ijhc", P, gz).reshape(1, n, n, hd)
            else:
                d_pv = (P.permute(0, 2, 3, 1).unsqueeze(-1) * gz.unsqueeze(2)).reshape(B, n, n, hd)
        d_hw = d_hw_rows.sum(0) if ctx.needs_input_grad[6] else None
        return d_proj, None, None, d_bias, d_pv, None, d_hw, None, None


def debug_umma_gemm(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """tcgen05 self-test: a [128,K] bf16, b [N,K] bf16 -> a @ b.T in fp32 via UMMA/TMEM."""
    a, b = _dev(a, torch.bfloat16, "a"), _dev(b, torch.bfloat16, "b")
    out = torch.empty(128, b.shape[0], dtype=torch.float32, device=a.device)
    with _guard(a):
        L.check(L.lib().se3_debug_umma_gemm(_p(a), _p(b), _p(out), b.shape[0], a.shape[1], _stream(a)), "se3_debug_umma_gemm")
    return out


def ipa_head_major_perm(heads: int, dk: int, device=None) -> torch.Tensor:
    """Row permutation of the fused projection weight [q | k | v | q_pt | k_pt | v_pt] (block-major, the reference's
    parameter order) into head-major order: head h owns the contiguous record [q dk | k dk | v dk | qp 12 | kp 12 | vp 24]."""
    hd = heads * dk
    ar = lambda n: torch.arange(n)           # built on the host: 6*H tiny device launches otherwise
    idx = []
    for h in range(heads):
        idx += [h * dk + ar(dk), hd + h * dk + ar(dk), 2 * hd + h * dk + ar(dk), 3 * hd + h * 12 + ar(12),
                3 * hd + 12 * heads + h * 12 + ar(12), 3 * hd + 24 * heads + h * 24 + ar(24)]
    return torch.cat(idx).to(device) if device is not None else torch.cat(idx)


def ipa_shape(batch: int, length: int, heads: int, dk: int, pair_batch: int, head_major: bool) -> L.IpaShape:
    """Column map of the fused projection output for the two supported row layouts."""
    hd, rec = heads * dk, 3 * dk + 48
    if head_major:
        return L.IpaShape(batch, length, heads, dk, 4, 8, heads * rec, 0, dk, 2 * dk, 3 * dk, 3 * dk + 12, 3 * dk + 24, rec, rec, rec,
                          pair_batch)
    return L.IpaShape(batch, length, heads, dk, 4, 8, heads * rec, 0, hd, 2 * hd, 3 * hd, 3 * hd + 12 * heads, 3 * hd + 24 * heads,
                      dk, 12, 24, pair_batch)


def ipa_tc_supported(shape: L.IpaShape) -> bool:
    return shape.dk == 16 and shape.pq == 4 and shape.pv == 8 and shape.pair_batch == 1 and shape.len <= 512


def _packed_pair_sizes(length: int, heads: int):
    bb, vb = C.c_int64(0), C.c_int64(0)
    if L.lib().se3_ipa_tc_packed_pair_bytes(int(length), int(heads), C.byref(bb), C.byref(vb)) < 0:
        raise ValueError(f"ipa_tc pack: bad shape L={length} H={heads}")
    return bb.value, vb.value


def ipa_tc_bias_pitch(length: int, heads: int = 1) -> int:
    """Row pitch (bf16 elements) of the packed pair bias, as the library lays it out: query-major [H][L(i)][pitch(j)] for L <= 128,
    key-major [H][L(j)][pitch(i)] for longer chains (csrc/common.cuh: ipa_bias_pitch)."""
    return _packed_pair_sizes(length, heads)[0] // (2 * heads * length)


def ipa_tc_pack_pair_value(pair_value: torch.Tensor, heads: int) -> torch.Tensor:
    """[1, L, L, H*16] (pair_value(x2d), structure_module.py:209) -> bf16 [L][H][Lp/8][16][8], the UMMA K-major
    operand layout read by pass 2 of the tensor-core attention (se3_ipa_tc_pack_pair)."""
    Lq = pair_value.shape[1]
    if pair_value.numel() != Lq * Lq * heads * 16:
        raise ValueError(f"pair_value must be [1, L, L, H*16], got {tuple(pair_value.shape)} for H={heads}")
    pv = _dev(pair_value, name="pair_value")
    Lp = (Lq + 15) // 16 * 16
    out = torch.empty(Lq, heads, Lp // 8, 16, 8, dtype=torch.bfloat16, device=pv.device)
    assert out.numel() * 2 == _packed_pair_sizes(Lq, heads)[1]
    with _guard(pv):
        L.check(L.lib().se3_ipa_tc_pack_pair(None, _p(pv), None, _p(out), Lq, heads, _stream(pv)), "se3_ipa_tc_pack_pair")
    return out


def ipa_tc_pack_pair_bias(pair_bias: torch.Tensor) -> torch.Tensor:
    """[1, L(i), L(j), H] (= pair_weight * pair_bias(x2d), structure_module.py:179) -> the bf16 slab layout the tensor-core attention
    fetches with TMA (se3_ipa_tc_pack_pair): [H][L(i)][pitch(j)] for L <= 128, [H][L(j)][round_up(L,8)(i)] for longer chains."""
    Lq, heads = pair_bias.shape[1], pair_bias.shape[-1]
    if pair_bias.numel() != Lq * Lq * heads:
        raise ValueError(f"pair_bias must be [1, L, L, H], got {tuple(pair_bias.shape)}")
    pb = _dev(pair_bias, name="pair_bias")
    out = torch.empty(heads, Lq, ipa_tc_bias_pitch(Lq, heads), dtype=torch.bfloat16, device=pb.device)
    assert out.numel() * 2 == _packed_pair_sizes(Lq, heads)[0]
    with _guard(pb):
        L.check(L.lib().se3_ipa_tc_pack_pair(_p(pb), None, _p(out), None, Lq, heads, _stream(pb)), "se3_ipa_tc_pack_pair")
    return out


def pair_precompute_supported(dim_embed: int, dim_pair: int) -> bool:
    return dim_embed % 32 == 0 and dim_pair % 32 == 0


def pair_embed(pair_dense, ln_weight, ln_bias, ln_eps: float, w_x2d, relpos_table, bucket) -> torch.Tensor:
    """x2d = x2d_proj(pair) + relative-position bias (models.py:243-293): pair_dense [Bp, L, L, de] -> [Bp, L, L, dp] (se3_pair_embed)."""
    pd = _dev(pair_dense, name="pair_dense")
    Bp, Lq, de = pd.shape[0], pd.shape[1], pd.shape[-1]
    w, rp = _dev(w_x2d, name="w_x2d"), _dev(relpos_table, name="relpos_table")
    bk = _dev(bucket.reshape(-1), torch.int32, "bucket")
    g, b = _dev(ln_weight, name="ln_weight"), _dev(ln_bias, name="ln_bias")
    dp = w.shape[0]
    out = torch.empty(Bp, Lq, Lq, dp, dtype=torch.float32, device=pd.device)
    stats = torch.empty(Bp * Lq * Lq * 2, dtype=torch.float32, device=pd.device)
    with _guard(pd):
        L.check(L.lib().se3_pair_embed(_p(pd), _p(g), _p(b), float(ln_eps), _p(w), _p(rp), _p(bk), _p(out), _p(stats), Bp, Lq, de, dp, _stream(pd)),
                "se3_pair_embed")
    return out


def pair_project(x2d, w_bias, w_value, pair_weight: float, heads: int, dk: int, packed: bool):
    """One layer's pair tensors from x2d [Bp, L, L, dp] (structure_module.py:179, 209; se3_pair_project): fp32 ([Bp, H, L, L], [Bp, L, L, H*dk])
    or, packed, the bf16 operands of the tensor-core attention ([H][L][ipa_tc_bias_pitch(L)], [L][H][Lp/8][16][8])."""
    x = _dev(x2d, name="x2d")
    Bp, Lq, dp = x.shape[0], x.shape[1], x.shape[-1]
    w = torch.cat([_dev(w_bias, name="w_bias"), _dev(w_value, name="w_value")], dim=0).contiguous()
    if packed:
        bias = torch.empty(heads, Lq, ipa_tc_bias_pitch(Lq, heads), dtype=torch.bfloat16, device=x.device)
        value = torch.empty(Lq, heads, (Lq + 15) // 16 * 2, 16, 8, dtype=torch.bfloat16, device=x.device)
    else:
        bias = torch.empty(Bp, heads, Lq, Lq, dtype=torch.float32, device=x.device)
        value = torch.empty(Bp, Lq, Lq, heads * dk, dtype=torch.float32, device=x.device)
    with _guard(x):
        L.check(L.lib().se3_pair_project(_p(x), _p(w), float(pair_weight), _p(bias), _p(value), int(packed), Bp, Lq, heads, dk, dp, _stream(x)),
                "se3_pair_project")
    return bias, value


def ipa_tc_workspace(shape: L.IpaShape, device) -> tuple[torch.Tensor, torch.Tensor]:
    pb, ib = C.c_int64(0), C.c_int64(0)
    L.lib().se3_ipa_tc_workspace_bytes(C.byref(shape), C.byref(pb), C.byref(ib))
    # the row-sum buffer ends in 64 bytes of work-queue state that must be zero before the first call (include/se3diff_b200.h)
    return (torch.empty(pb.value // 2, dtype=torch.bfloat16, device=device), torch.zeros(ib.value // 4, dtype=torch.float32, device=device))


def ipa_split_perms(heads: int, dk: int = 16):
    """Row index sets of the fused projection weight [q | k | v | q_pt | k_pt | v_pt] (structure_module.py:131-135) for the
    split layout of se3_ipa_attention_tc_fwd: (scalar rows, point rows, positions of the q rows inside the first), from the
    library's host function se3_ipa_split_perm."""
    sc, pt, qpos = (C.c_int32 * (heads * 3 * dk))(), (C.c_int32 * (heads * 48))(), (C.c_int32 * (heads * dk))()
    L.check(L.lib().se3_ipa_split_perm(int(heads), int(dk), sc, pt, qpos), "se3_ipa_split_perm")
    return (torch.tensor(list(sc), dtype=torch.int64), torch.tensor(list(pt), dtype=torch.int64), torch.tensor(list(qpos), dtype=torch.int64))


def _rows_view(t, dtype, name):
    """A 2-D tensor whose rows are contiguous (any row pitch): accepted as is; anything else is made contiguous."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        return _dev(t, dtype, name)
    if t.dtype == dtype and t.dim() == 2 and t.stride(1) == 1:
        return t
    return _dev(t, dtype, name)


def ipa_attention_tc_fwd(scalars, points, rot, trans, pair_bias_packed, pair_value_packed, key_bias, head_weight,
                               shape: L.IpaShape, workspace, out_dtype=torch.bfloat16, out=None):
    """Tensor-core edition (tcgen05/TMEM) of SAAttention.forward between the projections and fc_out, fed by head-major
    records: bf16 scalar records with pre-scaled q, and fp32 or bf16 point records (see ipa_split_perms).  Both may be column
    slices of one projection output (row pitch = that matrix's width)."""
    scalars = _rows_view(scalars, torch.bfloat16, "scalars")
    points = _rows_view(points, torch.bfloat16 if points.dtype == torch.bfloat16 else torch.float32, "points")
    rot, trans = _dev(rot, name="rot"), _dev(trans, name="trans")
    pair_bias = _dev(pair_bias_packed, torch.bfloat16, "pair_bias_packed")
    pvp = _dev(pair_value_packed, torch.bfloat16, "pair_value_packed")
    key_bias = None if key_bias is None else _dev(key_bias, name="key_bias")
    head_weight = _dev(head_weight, name="head_weight")
    if out is None:
        out = torch.empty(shape.batch * shape.len, shape.heads * (2 * shape.dk + 4 * shape.pv), dtype=out_dtype, device=points.device)
    pws, iws = workspace
    with _guard(points):
        L.check(L.lib().se3_ipa_attention_tc_fwd(_p(scalars), scalars.stride(0), _p(points), int(points.dtype == torch.bfloat16), points.stride(0),
                                                       _p(rot), _p(trans), _p(pair_bias), _p(pvp), _p(key_bias), _p(head_weight), _p(out),
                                                       int(out.dtype == torch.bfloat16), _p(pws), _p(iws), C.byref(shape), _stream(points)),
                "se3_ipa_attention_tc_fwd")
    return out


def folded_proportion(coords: torch.Tensor, ref_coords: torch.Tensor, k: float = -24.0, d_0: float = 0.4, tol: float = 1e-7,
                      want_drmsd: bool = False):
    """compute_folded_proportion (observables/folding_stability.py:52-81): coords [B, L, 3], ref_coords [L, 3] -> p [B]."""
    c, r = _dev(coords, name="coords"), _dev(ref_coords, name="ref_coords")
    if c.dim() != 3 or c.shape[-1] != 3 or tuple(r.shape) != (c.shape[1], 3):
        raise ValueError(f"coords [B, L, 3] and ref_coords [L, 3] expected, got {tuple(c.shape)} and {tuple(r.shape)}")
    p = torch.empty(c.shape[0], dtype=torch.float32, device=c.device)
    d = torch.empty_like(p) if want_drmsd else None
    with _guard(c):
        L.check(L.lib().se3_folded_proportion(_p(c), _p(r), _p(p), _p(d), c.shape[0], c.shape[1], float(k), float(d_0), float(tol),
                                              _stream(c)), "se3_folded_proportion")
    return (p, d) if want_drmsd else p


def backbone_atoms(pos: torch.Tensor, rot: torch.Tensor, aatype: torch.Tensor, pos_is_known: torch.Tensor | None = None) -> torch.Tensor:
    """pos [B, L, 3] (Angstrom), rot [B, L, 3, 3], aatype [L] -> atoms [B, L, 5, 3] = N, CA, C, CB, O (convert_chemgraph.py:139-293)."""
    pos, rot = _dev(pos, name="pos"), _dev(rot, name="rot")
    if pos.dim() != 3 or tuple(rot.shape) != tuple(pos.shape[:2]) + (3, 3):
        raise ValueError(f"pos [B, L, 3] and rot [B, L, 3, 3] expected, got {tuple(pos.shape)} and {tuple(rot.shape)}")
    aa = _dev(aatype, torch.int32, "aatype")
    if aa.numel() != pos.shape[1] or int(aa.max()) > 19 or int(aa.min()) < 0:
        raise ValueError("aatype must hold one residue type in [0, 19] per residue")
    known = None if pos_is_known is None else _dev(pos_is_known.to(torch.uint8), torch.uint8, "pos_is_known")
    out = torch.empty(pos.shape[0], pos.shape[1], 5, 3, dtype=torch.float32, device=pos.device)
    with _guard(pos):
        L.check(L.lib().se3_backbone_atoms(_p(pos), _p(rot), _p(aa), _p(known), _p(out), pos.shape[0], pos.shape[1], _stream(pos)),
                "se3_backbone_atoms")
    return out


def physicality(atoms: torch.Tensor, aatype: torch.Tensor) -> torch.Tensor:
    """atoms [B, L, 5, 3] -> [B, 3]: max sequential CA-CA, max sequential C-N, min heavy-atom distance between residues >= 3 apart."""
    atoms = _dev(atoms, name="atoms")
    aa = _dev(aatype, torch.int32, "aatype")
    out = torch.empty(atoms.shape[0], 3, dtype=torch.float32, device=atoms.device)
    with _guard(atoms):
        L.check(L.lib().se3_physicality(_p(atoms), _p(aa), _p(out), atoms.shape[0], atoms.shape[1], _stream(atoms)), "se3_physicality")
    return out


def residual_layernorm(x, y, bias, gamma, beta, eps: float, out_dtype=torch.bfloat16):
    """x += y + bias (in place, skipped when y is None); returns LayerNorm(x) in `out_dtype`."""
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("x must be a contiguous fp32 tensor (it is updated in place)")
    x = _dev(x, name="x")
    y = None if y is None else _dev(y, torch.bfloat16 if y.dtype == torch.bfloat16 else torch.float32, "y")
    bias = None if bias is None else _dev(bias, name="bias")
    gamma, beta = _dev(gamma, name="gamma"), _dev(beta, name="beta")
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    dim = x.shape[-1]
    with _guard(x):
        L.check(L.lib().se3_residual_layernorm(_p(x), _p(y), int(y is not None and y.dtype == torch.bfloat16), _p(bias), _p(gamma), _p(beta), float(eps), _p(out),
                                               int(out_dtype == torch.bfloat16), x.numel() // dim, dim, _stream(x)),
                "se3_residual_layernorm")
    return out


def bias_relu_project3(y, b1, w3, b3, rot=None):
    """Tail of a diffusion head: relu(y + b1) @ w3.T + b3 with w3 [3, D] (structure_module.py:12-22), one pass over y;
    `rot` [rows, 3, 3]: additionally rotate each 3-vector by its residue's frame (models.py:305)."""
    y = _dev(y, name="y")
    rot = None if rot is None else _dev(rot, name="rot")
    if rot is not None and rot.numel() != y.shape[0] * 9:
        raise ValueError("bias_relu_project3: rot must hold one 3x3 matrix per row")
    b1, w3, b3 = _dev(b1, name="b1"), _dev(w3, name="w3"), _dev(b3, name="b3")
    rows, dim = y.shape
    if w3.shape != (3, dim) or b1.shape != (dim,) or b3.shape != (3,):
        raise ValueError(f"bias_relu_project3: shapes {tuple(y.shape)}, {tuple(b1.shape)}, {tuple(w3.shape)}, {tuple(b3.shape)}")
    out = torch.empty(rows, 3, dtype=torch.float32, device=y.device)
    with _guard(y):
        L.check(L.lib().se3_bias_relu_project3(_p(y), _p(b1), _p(w3), _p(b3), _p(rot), _p(out), rows, dim, _stream(y)), "se3_bias_relu_project3")
    return out


def gelu_bf16_(x):
    """Exact (erf) GELU in place on a contiguous bf16 tensor."""
    x = _dev(x, torch.bfloat16, "x")
    if x.numel() % 8:
        raise ValueError("gelu_bf16_: element count must be a multiple of 8")
    with _guard(x):
        L.check(L.lib().se3_gelu_bf16(_p(x), _p(x), x.numel(), _stream(x)), "se3_gelu_bf16")
    return x
