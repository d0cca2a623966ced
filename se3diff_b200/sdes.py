"""SDE objects at the drop-in boundary: `CosineVPSDE` (R3) and `DiGSO3SDE` (SO(3)), plus the SO(3)
free functions, all backed by the CUDA kernels of libse3diff_b200.

Reference surface mirrored (names, arguments, shapes, error behaviour):
  bioemu/src/bioemu/sde_lib.py:26-167   maybe_expand, SDE, BaseVPSDE, CosineVPSDE
  bioemu/src/bioemu/so3_sde.py:20-403   SO3SDE / DiGSO3SDE (nn.Module; .igso3 / .uso3 / .score_function
                                        sub-modules holding the non-persistent lookup buffers)
  bioemu/src/bioemu/so3_sde.py:406-911  rotvec_to_rotmat, rotmat_to_rotvec, apply_rotvec_to_rotmat, ...
  bioemu/src/bioemu/so3_sde.py:914-990  npz lookup cache, same file names and keys, so that tables are
                                        interchangeable with the reference's ~/.cache/bioemu/so3

The per-graph methods (`sde`, `marginal_prob`, `beta`, ...) are thin torch expressions kept for API
compatibility; the samplers evaluate the schedule once per step on the host (schedule.py) instead.
"""
from __future__ import annotations

import logging
import math
import os

import numpy as np
import torch
from torch import nn

from . import ops

logger = logging.getLogger(__name__)


# ------------------------------------------------------------------------------------------------
# shared helpers                                                              sde_lib.py:17-47
# ------------------------------------------------------------------------------------------------
def _broadcast_like(x, like):
    return x if like is None else x[(...,) + (None,) * (like.ndim - x.ndim)]


def maybe_expand(x, batch_idx=None, like=None):
    x = _broadcast_like(x, like)
    if batch_idx is None:
        return x
    if x.shape[0] == batch_idx.shape[0]:
        logging.warning("Warning: batch shape is == x shape, are you trying to expand something that is already expanded?")
    return x[batch_idx]


class SDE:
    """Interface of sde_lib.py:50-102."""

    T = 1.0

    def sde(self, x, t, batch_idx=None):
        raise NotImplementedError

    def marginal_prob(self, x, t, batch_idx=None):
        raise NotImplementedError

    def prior_sampling(self, shape, device=None):
        raise NotImplementedError

    def mean_coeff_and_std(self, x, t, batch_idx=None):
        return self.marginal_prob(torch.ones_like(x), t, batch_idx)

    def sample_marginal(self, x, t, batch_idx=None):
        mean, std = self.marginal_prob(x=x, t=t, batch_idx=batch_idx)
        return mean + std * torch.randn_like(x)


class CosineVPSDE(SDE):
    """dx = -1/2 beta(t) x dt + sqrt(beta(t)) dW with the cosine schedule (sde_lib.py:105-167)."""

    def __init__(self, s: float = 0.008):
        self.s = s
        self.c = np.cos(s / (1 + s) * np.pi / 2)

    def _phase(self, t):
        return (t + self.s) / (1 + self.s) * np.pi / 2

    def beta(self, t):
        return torch.tan(self._phase(t)) * np.pi / (1 + self.s)

    def _marginal_mean_coeff(self, t):
        return torch.clip(torch.cos(self._phase(t)) / self.c, 0, 1)

    def marginal_prob(self, x, t, batch_idx=None):
        a = self._marginal_mean_coeff(t)
        return maybe_expand(a, batch_idx, x) * x, maybe_expand(torch.sqrt(1.0 - a**2), batch_idx, x)

    def sde(self, x, t, batch_idx=None):
        b = self.beta(t)
        return -0.5 * maybe_expand(b, batch_idx, x) * x, maybe_expand(torch.sqrt(b), batch_idx, x)

    def prior_sampling(self, shape, device=None):
        return torch.randn(*shape, device=device)


BaseVPSDE = CosineVPSDE  # the reference's abstract parent; only the cosine schedule exists


# ------------------------------------------------------------------------------------------------
# SO(3) free functions (so3_sde.py:406-911) -> CUDA
# ------------------------------------------------------------------------------------------------
def rotvec_to_rotmat(rotation_vectors, tol: float = 1e-7):
    return ops.so3_exp(rotation_vectors, tol)


def rotmat_to_rotvec(rotation_matrices):
    return ops.so3_log(rotation_matrices)


def angle_from_rotmat(rotation_matrices):
    return ops.so3_angle(rotation_matrices)


def apply_rotvec_to_rotmat(rotation_matrices, rotation_vectors, tol: float = 1e-7):
    return ops.so3_compose_rotvec(rotation_matrices, rotation_vectors, tol)


def rot_transpose(mat):
    return torch.transpose(mat, -1, -2)


def rot_mult(mat_1, mat_2):
    return ops.so3_matmul(mat_1, mat_2)


def rot_vf(mat_t, mat_1):
    return ops.so3_rel_log(mat_t, mat_1)


def geodesic_t(t: float, mat, base_mat):
    return ops.so3_geodesic(base_mat, mat, t)


def scale_rotmat(rotation_matrix, scalar, tol: float = 1e-7):
    assert rotation_matrix.ndim - 1 == scalar.ndim
    return ops.so3_exp(ops.so3_log(rotation_matrix) * scalar, tol)


def rotquat_to_rotvec(rotation_quaternions):
    return ops.so3_from_quat(rotation_quaternions, True, False)[0]


def rotquat_to_rotmat(rotation_quaternions):
    return ops.so3_from_quat(rotation_quaternions, False, True)[1]


def vector_to_skew_matrix(v):
    o = torch.zeros_like(v[..., 0])
    return torch.stack([torch.stack([o, -v[..., 2], v[..., 1]], -1), torch.stack([v[..., 2], o, -v[..., 0]], -1),
                        torch.stack([-v[..., 1], v[..., 0], o], -1)], -2)


def skew_matrix_to_vector(m):
    return torch.stack([m[..., 2, 1], m[..., 0, 2], m[..., 1, 0]], -1)


def geodesic_dist(mat_1, mat_2):
    # |Log(m1^T m2)| * sqrt(2): tr(K K^T) = 2 |v|^2 for K = hat(v)  (so3_sde.py:854-867)
    return torch.sqrt(2.0 * ops.so3_rel_log(mat_1, mat_2).square().sum(-1))


def igso3_expansion(omega, sigma, l_grid, tol=1e-7):
    return ops.igso3_series(omega, sigma, int(l_grid.numel()) - 1, tol, want=("f",))["f"]


def digso3_expansion(omega, sigma, l_grid, tol=1e-7):
    return ops.igso3_series(omega, sigma, int(l_grid.numel()) - 1, tol, want=("df",))["df"]


def dlog_igso3_expansion(omega, sigma, l_grid, tol=1e-7):
    return ops.igso3_series(omega, sigma, int(l_grid.numel()) - 1, tol, want=("dlog",))["dlog"]


def igso3_marginal_pdf(omega, omega_0, sigma, l_grid, tol: float = 1e-7):
    return ops.igso3_marginal_pdf(omega, omega_0, sigma, int(l_grid.numel()), tol)


# ------------------------------------------------------------------------------------------------
# lookup tables + sampler + score sub-modules                                so3_sde.py:914-1715
# ------------------------------------------------------------------------------------------------
def _cache_tag(sigma_grid: torch.Tensor) -> str:
    return f"s{sigma_grid.min().item():04.3f}-{sigma_grid.max().item():04.3f}-{sigma_grid.shape[0]:d}"


def _build_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "se3diff_b200: building IGSO(3) lookup tables needs a CUDA device (the series kernels are CUDA-only, there "
            "is no CPU fallback); pass cache_dir pointing at existing npz tables or run on the GPU box")
    return torch.device("cuda", torch.cuda.current_device())


def _omega_points(n: int, exponent: int) -> torch.Tensor:
    """pi * linspace(0,1,n)^k with the fp32 linspace -> fp64 promotion of so3_sde.py:1165-1170, 1667-1671."""
    om = torch.linspace(0.0, 1, n).to(torch.float64)
    return (om**exponent) * np.pi


class BaseSampleSO3(nn.Module):
    """Inverse-transform sampler over a CDF lookup table (so3_sde.py:993-1286)."""

    so3_type = "base"

    def __init__(self, num_omega, sigma_grid, omega_exponent=3, tol=1e-7, interpolate=True, cache_dir=None,
                 overwrite_cache=False):
        super().__init__()
        if not interpolate:
            raise NotImplementedError("interpolate=False is not implemented (unused by every reference call site)")
        self.num_omega, self.omega_exponent, self.tol, self.interpolate = num_omega, omega_exponent, tol, interpolate
        self.register_buffer("sigma_grid", sigma_grid, persistent=False)
        omega_grid, cdf = self._setup_lookup(sigma_grid, cache_dir, overwrite_cache)
        self.register_buffer("omega_grid", omega_grid, persistent=False)
        self.register_buffer("cdf_igso3", cdf, persistent=False)

    _uniform = False
    l_max = 0

    def _get_cache_name(self) -> str:
        return f"cache_{self.so3_type}_{_cache_tag(self.sigma_grid)}_o{self.num_omega:d}-{self.omega_exponent:d}.npz"

    def _generate_lookup(self, sigma_grid):
        dev = _build_device()
        pts = _omega_points(self.num_omega + 1, self.omega_exponent)
        cdf = ops.igso3_build_cdf(sigma_grid.to(dev, torch.float32), pts.to(dev), self.l_max, self.tol, uniform=self._uniform)
        return pts[1:].to(sigma_grid.dtype), cdf.to(sigma_grid.device, sigma_grid.dtype)

    def _setup_lookup(self, sigma_grid, cache_dir, overwrite_cache):
        path = None if cache_dir is None else os.path.join(os.path.expanduser(cache_dir), self._get_cache_name())
        if path is not None and os.path.exists(path) and not overwrite_cache:
            z = np.load(path)
            return torch.from_numpy(z["omega_grid"]).to(sigma_grid.dtype), torch.from_numpy(z["cdf_igso3"]).to(sigma_grid.dtype)
        omega_grid, cdf = self._generate_lookup(sigma_grid)
        if path is not None:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            np.savez(path, omega_grid=omega_grid.cpu().numpy(), cdf_igso3=cdf.cpu().numpy())
        return omega_grid, cdf

    def get_sigma_idx(self, sigma):
        return torch.bucketize(sigma, self.sigma_grid)

    def sample(self, sigma: torch.Tensor, num_samples: int, normals=None, u=None, left=None) -> torch.Tensor:
        """[n] std devs -> [n, num_samples, 3, 3].  RNG order of the reference: axis normals
        `randn(n, m, 3)` then uniforms `rand(n, m)` (so3_sde.py:1204-1205, 1240, 1262)."""
        dev = self.cdf_igso3.device
        n, m = sigma.shape[0], num_samples
        if (normals is None) != (u is None):
            raise ValueError("BaseSampleSO3.sample: pass both `normals` [n, m, 3] and `u` [n, m], or neither")
        if normals is None:
            normals = noise_randn((n, m, 3), dev)
            u = noise_rand((n, m), dev)
        sig = None if self._uniform else sigma.to(dev, torch.float32).reshape(n, 1).expand(n, m).reshape(-1)
        # search index over the CDF rows, rebuilt whenever the table buffer is replaced, moved or written to
        key = (self.cdf_igso3.data_ptr(), self.cdf_igso3._version)
        if getattr(self, "_index_key", None) != key:
            self._cdf_index = ops.igso3_build_cdf_index(self.cdf_igso3) if self.cdf_igso3.shape[1] <= 65535 else None
            self._index_key = key
        out = ops.igso3_sample(self.cdf_igso3, self.omega_grid, n * m, sigma=sig, cdf_index=self._cdf_index,
                               sigma_grid=None if self._uniform else self.sigma_grid, normals=normals.reshape(-1, 3),
                               u=u.reshape(-1), x=left, tol=self.tol)
        return out.view(n, m, 3, 3)


class SampleIGSO3(BaseSampleSO3):
    so3_type = "igso3"

    def __init__(self, num_omega, sigma_grid, omega_exponent=3, tol=1e-7, interpolate=True, l_max=1000, cache_dir=None,
                 overwrite_cache=False):
        self.l_max = l_max
        super().__init__(num_omega, sigma_grid, omega_exponent, tol, interpolate, cache_dir, overwrite_cache)

    def _get_cache_name(self) -> str:
        return (f"cache_{self.so3_type}_{_cache_tag(self.sigma_grid)}_l{self.l_max:d}_o{self.num_omega:d}-"
                f"{self.omega_exponent:d}.npz")


class SampleUSO3(BaseSampleSO3):
    so3_type = "uso3"
    _uniform = True

    def get_sigma_idx(self, sigma):
        return torch.zeros_like(sigma).long()

    def sample_shape(self, num_sigma: int, num_samples: int) -> torch.Tensor:
        return self.sample(torch.zeros(num_sigma, device=self.sigma_grid.device), num_samples)


class ScoreSO3(nn.Module):
    """IGSO(3) score and its tabulated RMS scaling (so3_sde.py:1495-1715)."""

    def __init__(self, num_omega, sigma_grid, omega_exponent=3, l_max=1000, tol=1e-7, cache_dir=None, overwrite_cache=False):
        super().__init__()
        self.l_max, self.tol, self.num_omega, self.omega_exponent = l_max, tol, num_omega, omega_exponent
        self.register_buffer("l_grid", torch.arange(l_max + 1), persistent=False)
        self.register_buffer("sigma_grid", sigma_grid, persistent=False)
        path = None if cache_dir is None else os.path.join(
            os.path.expanduser(cache_dir),
            f"cache_score-scaling_{_cache_tag(sigma_grid)}_l{l_max + 1:d}_o{num_omega:d}-{omega_exponent:d}.npz")
        if path is not None and os.path.exists(path) and not overwrite_cache:
            scaling = torch.from_numpy(np.load(path)["score_scaling"]).to(sigma_grid.dtype)
        else:
            scaling = self._compute_score_scaling(sigma_grid)
            if path is not None:
                os.makedirs(os.path.dirname(path), exist_ok=True)
                np.savez(path, score_scaling=scaling.cpu().numpy())
        self.register_buffer("score_scaling", scaling, persistent=False)

    def _compute_score_scaling(self, sigma_grid):
        dev = _build_device()
        pts = _omega_points(self.num_omega, self.omega_exponent)
        out = ops.igso3_build_score_scaling(sigma_grid.to(dev, torch.float32), pts.to(dev), self.l_max, self.tol)
        return out.to(sigma_grid.device, sigma_grid.dtype)

    def get_sigma_idx(self, sigma):
        return torch.bucketize(sigma, self.sigma_grid)

    def get_score_scaling(self, sigma):
        return self.score_scaling[self.get_sigma_idx(sigma)].detach()

    def forward(self, sigma, rotation_vectors):
        return ops.igso3_score(rotation_vectors, sigma, self.l_max, self.tol)


class SO3SDE(SDE, nn.Module):
    """Driftless SO(3) SDE dR = g(t) dB (so3_sde.py:20-288)."""

    def __init__(self, eps_t=1e-4, num_sigma=1000, num_omega=1000, omega_exponent=3, l_max=1000, tol=1e-7, cache_dir=None,
                 overwrite_cache=False):
        nn.Module.__init__(self)
        sigma_grid = self._marginal_std(torch.linspace(eps_t, self.T, num_sigma))
        self.tol = tol
        if "AMLT_EXPERIMENT_NAME" in os.environ and ("RANK" in os.environ or "OMPI_COMM_WORLD_RANK" in os.environ):
            cache_dir = None  # so3_sde.py:80-84
        kw = dict(num_omega=num_omega, sigma_grid=sigma_grid, omega_exponent=omega_exponent, tol=tol, cache_dir=cache_dir,
                  overwrite_cache=overwrite_cache)
        self.igso3 = SampleIGSO3(l_max=l_max, **kw)
        self.uso3 = SampleUSO3(**kw)
        self.score_function = ScoreSO3(l_max=l_max, **kw)

    # schedule ------------------------------------------------------------------------------------
    def beta(self, t):
        raise NotImplementedError

    def _marginal_std(self, t):
        raise NotImplementedError

    def sde(self, x, t, batch_idx=None):
        drift = torch.zeros_like(x[..., 0])
        return drift, maybe_expand(self.beta(t), batch_idx, drift)

    def marginal_prob(self, x, t, batch_idx=None):
        return x, maybe_expand(self._marginal_std(t), batch_idx)

    def mean_coeff_and_std(self, x, t, batch_idx=None):
        mean, std = self.marginal_prob(torch.ones_like(x[..., 0]), t, batch_idx)
        return mean, _broadcast_like(std, mean)

    # score -----------------------------------------------------------------------------------------
    def compute_score(self, rotation_vectors, t, batch_idx=None):
        sigma = maybe_expand(self._marginal_std(t), batch_idx)
        return self.score_function(sigma, rotation_vectors)

    def get_score_scaling(self, t, batch_idx=None):
        return maybe_expand(self.score_function.get_score_scaling(self._marginal_std(t)), batch_idx)

    # sampling ----------------------------------------------------------------------------------------
    def prior_sampling(self, shape, device=None):
        """Uniform SO(3) (so3_sde.py:206-247).  `shape` is [..., 3, 3] or the leading shape."""
        shape = tuple(shape)
        if len(shape) > 2:
            assert shape[-2:] == (3, 3)
            shape = shape[:-2]
        assert len(shape) <= 2
        squeeze = len(shape) == 1
        if squeeze:
            shape = shape + (1,)
        samples = self.uso3.sample_shape(*shape)
        if device is not None:
            samples = samples.to(device)
        return samples.squeeze(-3) if squeeze else samples

    @torch.no_grad()
    def sample_marginal(self, x, t, batch_idx=None):
        """x(t) | x(0) = x . r, r ~ IGSO3(sigma(t)) (so3_sde.py:249-288); fused into one kernel."""
        _, std = self.marginal_prob(x=x, t=t, batch_idx=batch_idx)
        lead = x.shape[:-2]
        dev = self.igso3.cdf_igso3.device
        if tuple(std.shape) == tuple(lead):            # one rotation per matrix (the sparse [N, 3, 3] layout): sample and compose in one kernel
            out = self.igso3.sample(std.reshape(-1), 1, left=x.reshape(-1, 3, 3).to(dev))
            return out.view(*lead, 3, 3)
        if std.dim() == 1 and len(lead) > 1 and std.shape[0] == lead[0]:
            # dense [B, L, 3, 3] with per-graph t: the reference draws ONE rotation per graph and applies it to every frame of
            # that graph (einsum "b...j,b...sjk->b...sk" with r [B, 1, 3, 3], so3_sde.py:274-283)
            r = self.igso3.sample(std, 1)[:, 0]                                   # [B, 3, 3]
            return torch.matmul(x.to(dev), r.view(lead[0], *([1] * (len(lead) - 1)), 3, 3))
        raise ValueError(f"sample_marginal: std of shape {tuple(std.shape)} does not match x of shape {tuple(x.shape)}")


class DiGSO3SDE(SO3SDE):
    """Geometric sigma schedule of the DiG paper (so3_sde.py:291-403)."""

    def __init__(self, eps_t=1e-4, num_sigma=1000, num_omega=2000, omega_exponent=3, l_max=2000, sigma_min=0.02,
                 sigma_max=1.65, tol=1e-7, cache_dir=None, overwrite_cache=False):
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        super().__init__(eps_t=eps_t, num_sigma=num_sigma, num_omega=num_omega, omega_exponent=omega_exponent, l_max=l_max,
                         tol=tol, cache_dir=cache_dir, overwrite_cache=overwrite_cache)

    def _marginal_std(self, t):
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def beta(self, t):
        return self._marginal_std(t) * np.sqrt(2.0 * np.log(self.sigma_max / self.sigma_min))


# ------------------------------------------------------------------------------------------------
# noise source.  Default: torch's generator of the target device (what the reference does on a GPU).
# `host_noise()` draws every variate on the CPU from torch's global generator in the reference's
# order and ships it to the device -- this is how the parity tests feed the CUDA path and the CPU
# oracle identical noise streams.
# ------------------------------------------------------------------------------------------------
_HOST_NOISE = False


class host_noise:
    def __enter__(self):
        global _HOST_NOISE
        self._prev, _HOST_NOISE = _HOST_NOISE, True
        return self

    def __exit__(self, *exc):
        global _HOST_NOISE
        _HOST_NOISE = self._prev
        return False


def noise_randn(shape, device):
    if _HOST_NOISE:
        return torch.randn(*shape).to(device)
    return torch.randn(*shape, device=device)


def noise_rand(shape, device):
    if _HOST_NOISE:
        return torch.rand(*shape).to(device)
    return torch.rand(*shape, device=device)


def noise_multinomial(weights: torch.Tensor, num_samples: int) -> torch.Tensor:
    """torch.multinomial(weights, n, replacement=True) on the noise source of the current mode."""
    if _HOST_NOISE:
        return torch.multinomial(weights.cpu(), num_samples, replacement=True).to(weights.device)
    return torch.multinomial(weights, num_samples, replacement=True)
