"""Edge cases of the elementwise C-ABI entries (K1c / K2 / K3), as size-independent properties: a residue's result does not
depend on how many residues travel with it, where the arrays start, or whether the batch is empty.

For every entry the outputs of a call on a PREFIX of the operands (lengths 0, 1, 31, 33, 255, 257: ragged last warps and tiles)
and on a window that starts one element in (every [n,3] array then starts 12 bytes off a 16-byte boundary and every [n,3,3]
array 4 bytes off: the kernels' unaligned editions) are BIT-IDENTICAL to the same rows of one call on all 1000 residues.  The
values themselves are pinned to the oracle and to the reference's goldens in test_gpu_parity.py; this file pins the raggedness /
alignment handling around them (the reference's own tests exercise empty and single-element batches through torch broadcasting,
tests/test_so3_utils.py)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
N = 1000
WINDOWS = [(0, 0), (0, 1), (0, 31), (0, 33), (0, 255), (0, 257), (1, 1), (1, 2), (1, 34), (1, 258), (3, 1000), (5, 5 + 64)]


def _as_tuple(o):
    if isinstance(o, dict):
        return tuple(o[k] for k in sorted(o) if o[k] is not None)
    if isinstance(o, (tuple, list)):
        return tuple(t for t in o if t is not None)
    return (o,)


def _check(name, fn, operands):
    """fn(*operands) on windows of the operands against the rows of the full call."""
    with torch.no_grad():
        full = _as_tuple(fn(*operands))
        torch.cuda.synchronize()
        for lo, hi in WINDOWS:
            part = _as_tuple(fn(*[t[lo:hi] for t in operands]))
            torch.cuda.synchronize()
            assert len(part) == len(full), name
            for k, (p, f) in enumerate(zip(part, full)):
                assert p.shape == f[lo:hi].shape, (name, lo, hi, k, tuple(p.shape))
                # bit-identical, NaNs (none expected) would fail the comparison
                assert torch.equal(p, f[lo:hi]), (name, lo, hi, k, float((p - f[lo:hi]).abs().max()) if p.numel() else 0.0)


@pytest.fixture(scope="module")
def data():
    from se3diff_b200 import ops

    g = torch.Generator(device=DEV).manual_seed(11)
    r = lambda *s: torch.randn(*s, generator=g, device=DEV)
    v = r(N, 3)
    v[:8] *= torch.tensor([0.0, 1e-9, 1e-7, 1e-3, 1.0, 3.0, 3.13, 3.14159], device=DEV)[:, None] / v[:8].norm(dim=-1, keepdim=True).clamp_min(1e-30)
    d = dict(v=v, w=0.1 * r(N, 3), m_rot=r(N, 3), m_rot2=r(N, 3), m_pos=r(N, 3), m_pos2=r(N, 3), z1=r(N, 3), z2=r(N, 3), u1=r(N, 3), u2=r(N, 3),
             pos=10.0 * r(N, 3), pos2=10.0 * r(N, 3), quat=r(N, 4))
    d["rot"] = ops.so3_exp(v)
    d["rot2"] = ops.so3_exp(r(N, 3))
    d["sigma"] = 0.02 * (2.33 / 0.02) ** torch.rand(N, generator=g, device=DEV)
    d["omega"] = torch.rand(N, generator=g, device=DEV) * math.pi
    d["omega"][:3] = torch.tensor([0.0, 1e-8, math.pi], device=DEV)
    d["u"] = torch.rand(N, generator=g, device=DEV)
    return d


def test_so3_entries_on_ragged_unaligned_and_empty_batches(data):
    from se3diff_b200 import ops

    d = data
    _check("so3_exp", ops.so3_exp, [d["v"]])
    _check("so3_log", ops.so3_log, [d["rot"]])
    _check("so3_angle", ops.so3_angle, [d["rot"]])
    _check("so3_compose_rotvec", ops.so3_compose_rotvec, [d["rot"], d["w"]])
    _check("so3_matmul", ops.so3_matmul, [d["rot"], d["rot2"]])
    _check("so3_matmul(transpose_a)", lambda a, b: ops.so3_matmul(a, b, transpose_a=True), [d["rot"], d["rot2"]])
    _check("so3_rel_log", ops.so3_rel_log, [d["rot"], d["rot2"]])
    _check("so3_geodesic", lambda a, b: ops.so3_geodesic(a, b, 0.37), [d["rot"], d["rot2"]])
    _check("so3_from_quat", ops.so3_from_quat, [d["quat"]])


def test_frame_update_entries_on_ragged_unaligned_and_empty_batches(data):
    from se3diff_b200 import _lib as L
    from se3diff_b200 import ops

    d = data
    em = L.EmScalars(-0.02, 0.1414, 1.0, 1.0, 0.67, 4.0, 3.1, 1.76, 0.7, 1e-7)
    dp = L.DpmScalars(0.7, 1.01, 0.02, 0.69, 1.02, 0.04, 4.0, 4.1, 0.67, 0.65, -0.01, -0.02, 1e-7)
    hs = L.HeunScalars(0.004, 0.0632, 0.66, 3.0, 1.73, -0.01, 0.67, 4.0, 3.1, 1.76, 0.7, 0.65, 4.1, 3.0, 1.73, 0.69, 1e-7)
    _check("frame_update_em", lambda rot, pos, mr, mp, z1, z2: ops.frame_update_em(rot, pos, mr, mp, z1, z2, em),
           [d["rot"], d["pos"], d["m_rot"], d["m_pos"], d["z1"], d["z2"]])
    _check("frame_update_em(u, dW)", lambda rot, pos, mr, mp, z1, z2, u1, u2: ops.frame_update_em(rot, pos, mr, mp, z1, z2, em, u_rot=u1, u_pos=u2, want_dw=True),
           [d["rot"], d["pos"], d["m_rot"], d["m_pos"], d["z1"], d["z2"], d["u1"], d["u2"]])
    _check("so3_update_em", lambda rot, mr, z1, u1: ops.so3_update_em(rot, mr, z1, em, u_rot=u1, want_dw=True), [d["rot"], d["m_rot"], d["z1"], d["u1"]])
    _check("r3_update_em", lambda pos, mp, z2, u2: ops.r3_update_em(pos, mp, z2, em, u_pos=u2, want_dw=True), [d["pos"], d["m_pos"], d["z2"], d["u2"]])
    _check("r3_update_dpm(mid)", lambda pos, mp: ops.r3_update_dpm(pos, mp, dp, final_half=False), [d["pos"], d["m_pos"]])
    _check("r3_update_dpm(final)", lambda pos, mp: ops.r3_update_dpm(pos, mp, dp, final_half=True), [d["pos"], d["m_pos"]])
    _check("r3_heun_churn", lambda pos, z2: ops.r3_heun_churn(pos, z2, hs), [d["pos"], d["z2"]])
    _check("r3_heun_step(first order)", lambda pos, mp: ops.r3_heun_step(pos, mp, hs), [d["pos"], d["m_pos"]])
    _check("r3_heun_step(corrected)", lambda pos, mp, pp, mn: ops.r3_heun_step(pos, mp, hs, pos_pred=pp, m_pos_next=mn), [d["pos"], d["m_pos"], d["pos2"], d["m_pos2"]])
    _check("frame_update_dpm_mid", lambda rot, pos, mr, mp: ops.frame_update_dpm_mid(rot, pos, mr, mp, dp), [d["rot"], d["pos"], d["m_rot"], d["m_pos"]])
    _check("frame_update_dpm_final", lambda rot, pos, mr, ml, mp: ops.frame_update_dpm_final(rot, pos, mr, ml, mp, dp),
           [d["rot"], d["pos"], d["m_rot"], d["m_rot2"], d["m_pos"]])
    _check("frame_heun_churn", lambda rot, pos, z1, z2: ops.frame_heun_churn(rot, pos, z1, z2, hs), [d["rot"], d["pos"], d["z1"], d["z2"]])
    _check("frame_heun_predict", lambda rot, pos, mr, mp: ops.frame_heun_predict(rot, pos, mr, mp, hs), [d["rot"], d["pos"], d["m_rot"], d["m_pos"]])
    _check("frame_heun_correct", lambda rot, pos, mr, mp, pp, mrn, mpn: ops.frame_heun_correct(rot, pos, mr, mp, pp, mrn, mpn, hs),
           [d["rot"], d["pos"], d["m_rot"], d["m_pos"], d["pos2"], d["m_rot2"], d["m_pos2"]])
    _check("frame_traceback", lambda rot, pos, rn, pn, mr, mp, u1, u2: ops.frame_traceback(rot, pos, rn, pn, mr, mp, em, u_rot=u1, u_pos=u2),
           [d["rot"], d["pos"], d["rot2"], d["pos2"], d["m_rot"], d["m_pos"], d["u1"], d["u2"]])


def test_igso3_entries_on_ragged_unaligned_and_empty_batches(data):
    from se3diff_b200 import ops

    d = data
    _check("igso3_series_f32", lambda om, sg: ops.igso3_series(om, sg, 500), [d["omega"], d["sigma"]])
    _check("igso3_series_f64", lambda om, sg: ops.igso3_series(om.double(), sg.double(), 500), [d["omega"], d["sigma"]])
    _check("igso3_score", lambda v, sg: ops.igso3_score(v, sg, 500), [d["v"], d["sigma"]])
    _check("igso3_marginal_pdf", lambda om, o0, sg: ops.igso3_marginal_pdf(om, o0, sg, 200), [d["omega"], d["omega"].flip(0).contiguous(), d["sigma"]])
    # the sampler: a 200-row table with guide records; noise passed in (every operand), and the in-kernel Philox draws, which
    # depend on (seed, element index) only -- so only PREFIXES of a Philox call can be compared
    sig_grid = 0.02 * (2.33 / 0.02) ** torch.linspace(0.001, 1.0, 200, device=DEV)
    om_pts = torch.linspace(0.0, 1, 501, device=DEV, dtype=torch.float64) ** 3 * math.pi
    cdf = ops.igso3_build_cdf(sig_grid, om_pts, 300)
    omg = om_pts[1:].float()
    idx = ops.igso3_build_cdf_index(cdf)
    for index in (idx, None):
        _check("igso3_sample(noise passed in)",
               lambda sg, z, u, x: ops.igso3_sample(cdf, omg, sg.shape[0], sigma=sg, sigma_grid=sig_grid, normals=z, u=u, x=x, cdf_index=index, want_angle=True),
               [d["sigma"], d["z1"], d["u"], d["rot"]])
        _check("igso3_sample(noise passed in, no x)",
               lambda sg, z, u: ops.igso3_sample(cdf, omg, sg.shape[0], sigma=sg, sigma_grid=sig_grid, normals=z, u=u, cdf_index=index), [d["sigma"], d["z1"], d["u"]])
    with torch.no_grad():
        full = ops.igso3_sample(cdf, omg, N, sigma=d["sigma"], sigma_grid=sig_grid, seed=7, x=d["rot"], cdf_index=idx)
        for k in (0, 1, 31, 33, 255, 257, 999):
            part = ops.igso3_sample(cdf, omg, k, sigma=d["sigma"][:k], sigma_grid=sig_grid, seed=7, x=d["rot"][:k], cdf_index=idx)
            assert part.shape == (k, 3, 3) and torch.equal(part, full[:k]), k
        # the draws are rotations
        eye = torch.eye(3, device=DEV)
        assert (full.transpose(-1, -2) @ full - eye).abs().max() < 5e-6
