"""CPU ORACLE (test infrastructure) -- SO(3) Lie-group maps and the IGSO(3) series/tables/sampler.

Restates bioemu/src/bioemu/so3_sde.py.  All functions are pure torch-on-CPU and follow the input
dtype exactly as the reference does (fp32 on the sampling path, fp64 for table construction).
"""
from __future__ import annotations

import math

import numpy as np
import torch

PI = np.pi  # the reference multiplies by np.pi (python float -> rounded to the tensor dtype)


# --------------------------------------------------------------------------------------------
# exp / log / composition                                              so3_sde.py:406-911
# --------------------------------------------------------------------------------------------
def hat(v: torch.Tensor) -> torch.Tensor:
    """so(3) vector -> skew matrix [[0,-z,y],[z,0,-x],[-y,x,0]] (so3_sde.py:679-705)."""
    x, y, z = v[..., 0], v[..., 1], v[..., 2]
    o = torch.zeros_like(x)
    return torch.stack(
        [torch.stack([o, -z, y], -1), torch.stack([z, o, -x], -1), torch.stack([-y, x, o], -1)], -2
    )


def vee(m: torch.Tensor) -> torch.Tensor:
    """skew matrix -> vector (m21, m02, m10) (so3_sde.py:708-722).  Filled into zeros_like(m[..., 0]) as the reference does:
    for permuted inputs that keeps the input's stride order, which decides the summation order of the norm taken next."""
    v = torch.zeros_like(m[..., 0])
    v[..., 0], v[..., 1], v[..., 2] = m[..., 2, 1], m[..., 0, 2], m[..., 1, 0]
    return v


def rotvec_to_rotmat(v: torch.Tensor, tol: float = 1e-7) -> torch.Tensor:
    """Rodrigues exp map with Taylor coefficients below ``tol`` (so3_sde.py:478-554)."""
    theta = torch.norm(v, dim=-1)[..., None, None]
    k = hat(v)
    th2 = theta.square()
    a = torch.sin(theta) / theta
    b = (1.0 - torch.cos(theta)) / th2
    small = torch.abs(theta) < tol
    a = torch.where(small, 1.0 - th2 / 6.0, a)
    b = torch.where(small, 0.5 - th2 / 24.0, b)
    eye = torch.eye(3, dtype=v.dtype, device=v.device).expand(k.shape)
    return eye + a * k + b * torch.einsum("...ik,...kj->...ij", k, k)


def angle_from_rotmat(r: torch.Tensor):
    """theta = atan2(|vee(R-R^T)|/2, (tr R - 1)/2) (so3_sde.py:651-676)."""
    sv = vee(r - r.transpose(-2, -1))
    s = torch.norm(sv, dim=-1) / 2.0
    c = (torch.einsum("...ii", r) - 1.0) / 2.0
    return torch.atan2(s, c), s, c


def rotmat_to_rotvec(r: torch.Tensor) -> torch.Tensor:
    """Three-regime log map with the reference's isclose thresholds (so3_sde.py:557-648)."""
    theta, s, _ = angle_from_rotmat(r)
    w = vee(r - r.transpose(-2, -1))
    m0 = torch.isclose(theta, torch.zeros_like(theta)).to(theta.dtype)
    mpi = torch.isclose(theta, torch.full_like(theta, PI), atol=1e-2).to(theta.dtype)
    me = (1 - m0) * (1 - mpi)
    num = m0 / 2.0 + theta * me
    den = (1.0 - theta**2 / 6.0) * m0 + 2.0 * s * me + mpi
    w = w * (num / den)[..., None]

    eye = torch.eye(3, dtype=r.dtype, device=r.device).expand(r.shape)
    outer = (eye + r) / 2.0
    outer = outer + (torch.relu(outer) - outer) * eye
    axis = torch.sqrt(torch.clamp(torch.diagonal(outer, dim1=-2, dim2=-1), min=1e-8))
    row = torch.argmax(torch.norm(outer, dim=-1), dim=-1).long()
    line = torch.take_along_dim(outer, dim=-2, indices=row[..., None, None]).squeeze(-2)
    wpi = axis * theta[..., None] * torch.sign(line)
    return w + wpi * mpi[..., None]


def rot_mult(a, b):  # so3_sde.py:875-877
    return torch.einsum("...ij,...jk->...ik", a, b)


def rot_transpose(a):  # so3_sde.py:870-872
    return a.transpose(-1, -2)


def apply_rotvec_to_rotmat(r: torch.Tensor, v: torch.Tensor, tol: float = 1e-7) -> torch.Tensor:
    """R . Exp(v) (so3_sde.py:782-802)."""
    return rot_mult(r, rotvec_to_rotmat(v, tol=tol))


def rot_vf(base: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Log(base^T target) (so3_sde.py:880-891)."""
    return rotmat_to_rotvec(rot_mult(rot_transpose(base), target))


def geodesic_t(t: float, mat: torch.Tensor, base: torch.Tensor) -> torch.Tensor:
    """base . Exp(t Log(base^T mat)) (so3_sde.py:894-911)."""
    return rot_mult(base, rotvec_to_rotmat(t * rot_vf(base, mat)))


def scale_rotmat(r: torch.Tensor, scalar: torch.Tensor, tol: float = 1e-7) -> torch.Tensor:
    """Exp(scalar . Log(R)) (so3_sde.py:406-425)."""
    return rotvec_to_rotmat(rotmat_to_rotvec(r) * scalar, tol=tol)


def geodesic_dist(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """sqrt(tr(A A^T)), A = hat(Log(a^T b)) (so3_sde.py:854-867)."""
    k = hat(rot_vf(a, b))
    return torch.sqrt(torch.einsum("...ii->...", rot_mult(k, rot_transpose(k))))


def rotquat_to_axis_angle(q: torch.Tensor, tol: float = 1e-7):
    """[r,i,j,k] unit quaternion -> (angle, axis) (so3_sde.py:725-748)."""
    ax = q[..., 1:]
    n = torch.norm(ax, dim=-1)
    ang = 2.0 * torch.atan2(n, q[..., 0])
    return ang, ax / (n[:, None] + tol)


def rotquat_to_rotvec(q: torch.Tensor) -> torch.Tensor:  # so3_sde.py:751-763
    ang, ax = rotquat_to_axis_angle(q)
    return ax * ang[..., None]


def rotquat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """so3_sde.py:766-779 -- note: the exp map is evaluated with angle `ang` and skew(ax*ang)."""
    ang, ax = rotquat_to_axis_angle(q)
    k = hat(ax * ang[..., None])
    theta = ang[..., None, None]
    th2 = theta.square()
    a = torch.sin(theta) / theta
    b = (1.0 - torch.cos(theta)) / th2
    small = torch.abs(theta) < 1e-7
    a = torch.where(small, 1.0 - th2 / 6.0, a)
    b = torch.where(small, 0.5 - th2 / 24.0, b)
    eye = torch.eye(3, dtype=q.dtype).expand(k.shape)
    return eye + a * k + b * torch.einsum("...ik,...kj->...ij", k, k)


# --------------------------------------------------------------------------------------------
# IGSO(3) truncated series                                            so3_sde.py:1731-1940
# --------------------------------------------------------------------------------------------
def _clean(x: torch.Tensor) -> torch.Tensor:
    return torch.where(torch.logical_or(torch.isinf(x), torch.isnan(x)), torch.zeros_like(x), x)


def igso3_expansion(omega, sigma, l_grid, tol=1e-7):
    """f(omega, sigma) = sum_l (2l+1) e^{-l(l+1)sigma^2/2} sin((l+1/2)omega) / (sin(omega/2)+tol)
    with the omega<=tol limit sum (2l+1)^2 a_l, NaN/Inf -> 0, clamp >= 0 (so3_sde.py:1731-1792)."""
    den = torch.sin(0.5 * omega)
    f1 = 2.0 * l_grid + 1.0
    f2 = -l_grid * (l_grid + 1.0)
    num = torch.sin((l_grid + 1 / 2) * omega.unsqueeze(-1))
    e = f1 * torch.exp(f2 * sigma.unsqueeze(-1) ** 2 / 2)
    f = torch.sum(e * num, dim=-1)
    flim = torch.sum(e * f1, dim=-1)
    f = f / (den + tol)
    f = torch.where(omega <= tol, flim, f)
    return torch.clamp(_clean(f), min=0.0)


def digso3_expansion(omega, sigma, l_grid, tol=1e-7):
    """d/d omega of the above, closed form per term (so3_sde.py:1857-1913)."""
    den = 1.0 - torch.cos(omega)
    f1 = 2.0 * l_grid + 1.0
    f2 = l_grid + 1.0
    f3 = -l_grid * f2
    num = l_grid * torch.sin(f2 * omega.unsqueeze(-1)) - f2 * torch.sin(l_grid * omega.unsqueeze(-1))
    df = torch.sum(f1 * torch.exp(f3 * sigma.unsqueeze(-1) ** 2 / 2) * num, dim=1)  # reference sums dim=1: omega must be 1-D
    df = df / (den + tol)
    df = torch.where(omega <= tol, torch.zeros_like(df), df)
    return _clean(df)


def dlog_igso3_expansion(omega, sigma, l_grid, tol=1e-7):
    """df / (f + tol) (so3_sde.py:1916-1940)."""
    return digso3_expansion(omega, sigma, l_grid, tol) / (igso3_expansion(omega, sigma, l_grid, tol) + tol)


def igso3_marginal_pdf(omega, omega_0, sigma, l_grid, tol=1e-7):
    """Mixture-component marginal angle pdf (so3_sde.py:1795-1854)."""
    d0 = torch.sin(0.5 * omega_0)
    d = torch.sin(0.5 * omega)
    f1 = 2.0 * l_grid + 1.0
    f2 = -l_grid * (l_grid + 1.0)
    n0 = torch.sin((l_grid + 1 / 2) * omega_0.unsqueeze(-1))
    n = torch.sin((l_grid + 1 / 2) * omega.unsqueeze(-1))
    e = torch.exp(f2 * sigma.unsqueeze(-1) ** 2 / 2)
    f = torch.sum(e * n * n0, dim=-1) * d / (d0 + tol)
    flim = torch.sum(e * f1 * n, dim=-1) * d
    f = torch.where(omega_0 <= tol, flim, f)
    f = _clean(f) * 2.0 / PI
    return torch.clamp(f, min=0.0)


def score_so3(sigma, rotvec, l_max, tol=1e-7):
    """q/(|q|+tol) * dlog f(|q|, sigma) with l = 0..l_max inclusive (so3_sde.py:1698-1715, 1554)."""
    l_grid = torch.arange(l_max + 1)
    ang = torch.norm(rotvec, dim=-1)
    d = dlog_igso3_expansion(ang, sigma, l_grid, tol=tol)
    return rotvec / (ang[..., None] + tol) * d[..., None]


# --------------------------------------------------------------------------------------------
# lookup tables (fp64 build, cast at the end)              so3_sde.py:1131-1187,1455-1492,1637-1696
# --------------------------------------------------------------------------------------------
def _table(fn, omega_grid, sigma_grid, l_max, tol):
    """Row loop of generate_lookup_table (so3_sde.py:1943-1983)."""
    l_grid = torch.arange(l_max + 1).to(omega_grid.dtype)
    out = torch.zeros(len(sigma_grid), len(omega_grid), dtype=omega_grid.dtype)
    for r in range(len(sigma_grid)):
        out[r] = fn(omega_grid, torch.ones_like(omega_grid) * sigma_grid[r], l_grid, tol=tol)
    return out


def trapz_cumulative(f, x):
    """so3_sde.py:1475-1492."""
    return torch.cumsum(((f[..., :-1] + f[..., 1:]) * torch.diff(x, dim=-1)[None, :]) / 2.0, dim=-1)


def build_cdf_table(sigma_grid, num_omega, omega_exponent=3, l_max=None, tol=1e-7):
    """CDF lookup for inverse-transform sampling.  ``l_max=None`` -> uniform SO(3) (one row).
    Returns (omega_grid[1:], cdf) cast to sigma_grid.dtype (so3_sde.py:1131-1187, 1455-1472)."""
    sg = sigma_grid.to(torch.float64)
    om = torch.linspace(0.0, 1, num_omega + 1).to(sg)
    om = om**omega_exponent
    om = om * PI
    if l_max is None:
        pdf = torch.ones(1, om.shape[0])
    else:
        pdf = _table(igso3_expansion, om, sg, l_max, tol)
    pdf = pdf * (1.0 - torch.cos(om)) / PI
    cdf = trapz_cumulative(pdf, om)
    cdf = cdf / cdf[:, -1][:, None]
    return om[1:].to(sigma_grid.dtype), cdf.to(sigma_grid.dtype)


def build_score_scaling(sigma_grid, num_omega, omega_exponent=3, l_max=2000, tol=1e-7):
    """sqrt(sum dlog^2 pdf / (3 sum pdf + tol)) on the grid WITHOUT the +1 (so3_sde.py:1637-1696)."""
    sg = sigma_grid.to(torch.float64)
    om = torch.linspace(0.0, 1, num_omega).to(sg)
    om = om**omega_exponent
    om = om * PI
    pdf = _table(igso3_expansion, om, sg, l_max, tol)
    pdf = torch.abs(pdf * ((1.0 - torch.cos(om)) / PI)[None, :])
    dlog = _table(dlog_igso3_expansion, om, sg, l_max, tol)
    sc = torch.sqrt(torch.sum(dlog**2 * pdf, dim=1) / (3.0 * torch.sum(pdf, dim=1) + tol))
    return sc.to(sigma_grid.dtype)


# --------------------------------------------------------------------------------------------
# inverse-CDF sampler                                                so3_sde.py:1189-1286,1374-1391
# --------------------------------------------------------------------------------------------
def sample_angle(cdf, omega_grid, sigma_idx, u, tol=1e-7):
    """u [n, m] uniforms, sigma_idx [n] rows of ``cdf`` (so3_sde.py:1244-1286)."""
    rows = cdf[sigma_idx, :]
    stop = torch.sum(rows[..., None] < u[:, None, :], dim=1).long()
    start = torch.clamp(stop - 1, min=0)
    c0 = torch.gather(rows, 1, start)
    c1 = torch.gather(rows, 1, stop)
    w = torch.clamp((u - c0) / torch.clamp(c1 - c0, min=tol), min=0.0, max=1.0)
    return torch.lerp(omega_grid[start], omega_grid[stop], w)


def sample_rotations(cdf, omega_grid, sigma_idx, normals, u, sigma=None, tol=1e-7):
    """normals [n, m, 3], u [n, m] -> [n, m, 3, 3].  ``sigma`` given -> IGSO3 zeroing of angles when
    sigma < tol (so3_sde.py:1189-1213, 1229-1242, 1374-1391)."""
    axis = normals / torch.norm(normals, dim=2, keepdim=True)
    ang = sample_angle(cdf, omega_grid, sigma_idx, u, tol)
    if sigma is not None:
        ang = torch.where(sigma[..., None] < tol, torch.zeros_like(ang), ang)
    return rotvec_to_rotmat(axis * ang[..., None], tol=tol)


class SO3Tables:
    """The three lookup buffers of DiGSO3SDE (so3_sde.py:77-116, 292-379)."""

    def __init__(self, eps_t=1e-4, num_sigma=1000, num_omega=2000, omega_exponent=3, l_max=2000,
                 sigma_min=0.02, sigma_max=1.65, tol=1e-7, tables: dict | None = None):
        self.sigma_min, self.sigma_max, self.tol, self.l_max = sigma_min, sigma_max, tol, l_max
        self.sigma_grid = self.marginal_std(torch.linspace(eps_t, 1.0, num_sigma))
        if tables is None:
            self.omega_grid, self.cdf_igso3 = build_cdf_table(self.sigma_grid, num_omega, omega_exponent, l_max, tol)
            _, self.cdf_uso3 = build_cdf_table(self.sigma_grid, num_omega, omega_exponent, None, tol)
            self.score_scaling = build_score_scaling(self.sigma_grid, num_omega, omega_exponent, l_max, tol)
        else:
            self.omega_grid = torch.as_tensor(tables["omega_grid"])
            self.cdf_igso3 = torch.as_tensor(tables["cdf_igso3"])
            self.cdf_uso3 = torch.as_tensor(tables["cdf_uso3"])
            self.score_scaling = torch.as_tensor(tables["score_scaling"])

    def marginal_std(self, t):  # so3_sde.py:365-379
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def beta(self, t):  # g(t), so3_sde.py:346-363
        return self.marginal_std(t) * np.sqrt(2.0 * np.log(self.sigma_max / self.sigma_min))

    def score_scaling_at(self, t):  # so3_sde.py:142-161, 1610-1635
        return self.score_scaling[torch.bucketize(self.marginal_std(t), self.sigma_grid)]

    def prior(self, n):
        """Uniform SO(3) prior; RNG order randn(n,1,3) then rand(n,1) (so3_sde.py:206-247, 1448-1453)."""
        normals = torch.randn(n, 1, 3)
        u = torch.rand(n, 1)
        idx = torch.zeros(n, dtype=torch.long)
        return sample_rotations(self.cdf_uso3, self.omega_grid, idx, normals, u, None, self.tol).squeeze(-3)

    def sample_igso3(self, sigma, normals=None, u=None):
        """IGSO3 draw for per-element sigma [n]; one sample each (so3_sde.py:278)."""
        n = sigma.shape[0]
        normals = torch.randn(n, 1, 3) if normals is None else normals
        u = torch.rand(n, 1) if u is None else u
        idx = torch.bucketize(sigma, self.sigma_grid)
        return sample_rotations(self.cdf_igso3, self.omega_grid, idx, normals, u, sigma, self.tol)

    def sample_marginal(self, x, t, normals=None, u=None):
        """x(t) | x(0) = x . r (so3_sde.py:249-288); t per element."""
        r = self.sample_igso3(self.marginal_std(t), normals, u)
        return torch.einsum("b...j,b...sjk->b...sk", x, r).squeeze(-2)

    def compute_score(self, rotvec, t):  # so3_sde.py:118-140
        return score_so3(self.marginal_std(t), rotvec, self.l_max, self.tol)
