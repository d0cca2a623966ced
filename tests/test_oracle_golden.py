"""Pins the CPU oracle (oracle/) against fixtures produced by the UNMODIFIED reference
(tests/golden/*.npz, minted by oracle/gen_golden.py) and against independent known answers
(scipy Rotation / matrix_exp, mirroring bioemu/tests/test_so3_utils.py).  CPU only."""
import numpy as np
import pytest
import torch
import yaml
from scipy.spatial.transform import Rotation

from oracle import samplers, so3
from oracle.score_model import ScoreModelOracle, relative_position_bucket

from conftest import load_golden

T = torch.from_numpy


def _sd(g, prefix):
    return {k[len(prefix):]: T(v) for k, v in g.items() if k.startswith(prefix)}


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_so3_maps_bit_exact(dt):
    g = load_golden("so3_maps.npz")
    v, w = T(g[f"v_{dt}"]), T(g[f"w_{dt}"])
    rm = so3.rotvec_to_rotmat(v)
    assert torch.equal(rm, T(g[f"exp_{dt}"]))
    assert torch.equal(so3.rotmat_to_rotvec(rm), T(g[f"log_{dt}"]))
    assert torch.equal(so3.apply_rotvec_to_rotmat(rm, w), T(g[f"compose_{dt}"]))
    assert torch.equal(so3.rot_vf(rm, so3.apply_rotvec_to_rotmat(rm, w)), T(g[f"rel_log_{dt}"]))
    assert torch.equal(so3.angle_from_rotmat(rm)[0], T(g[f"angle_{dt}"]))
    assert torch.equal(so3.geodesic_t(0.3, rm.flip(0), rm), T(g[f"geodesic_t_{dt}"]))
    assert torch.equal(so3.scale_rotmat(rm, torch.full((len(v), 1), 0.5, dtype=v.dtype)), T(g[f"scale_{dt}"]))
    q = T(g[f"quat_{dt}"])
    assert torch.equal(so3.rotquat_to_rotvec(q), T(g[f"quat_rotvec_{dt}"]))
    assert torch.equal(so3.rotquat_to_rotmat(q), T(g[f"quat_rotmat_{dt}"]))


def test_so3_maps_known_answers():
    """scipy is the independent authority (bioemu/tests/test_so3_utils.py:195-330 does the same)."""
    rng = np.random.default_rng(0)
    v = rng.normal(size=(200, 3))
    v = v / np.linalg.norm(v, axis=1, keepdims=True) * rng.uniform(0, np.pi - 0.02, size=(200, 1))
    rm = so3.rotvec_to_rotmat(T(v))
    assert np.allclose(rm.numpy(), Rotation.from_rotvec(v).as_matrix(), atol=1e-12)
    assert np.allclose(so3.rotmat_to_rotvec(rm).numpy(), v, atol=1e-9)
    assert np.allclose(rm.numpy(), torch.linalg.matrix_exp(so3.hat(T(v))).numpy(), atol=1e-12)
    q = Rotation.from_rotvec(v).as_quat()[:, [3, 0, 1, 2]]
    assert np.allclose(so3.rotquat_to_rotmat(T(q)).numpy(), Rotation.from_rotvec(v).as_matrix(), atol=1e-6)


def test_igso3_series_bit_exact():
    g = load_golden("igso3_series.npz")
    om, sg = T(g["omega"]), T(g["sigma"])
    for l_max in (2000, 500):
        lg = torch.arange(l_max + 1)
        for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
            o, s = om.to(dt), sg.to(dt)
            assert torch.equal(so3.igso3_expansion(o, s, lg), T(g[f"f_{name}_l{l_max}"]))
            assert torch.equal(so3.digso3_expansion(o, s, lg), T(g[f"df_{name}_l{l_max}"]))
            assert torch.equal(so3.dlog_igso3_expansion(o, s, lg), T(g[f"dlog_{name}_l{l_max}"]))
    lg = torch.arange(1000)
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        got = so3.igso3_marginal_pdf(om.to(dt), T(g["omega0"]).to(dt), sg.to(dt), lg)
        assert torch.equal(got, T(g[f"marginal_{name}"]))
    tt = T(g["score_t"])
    sigma = 0.02 * (2.33 / 0.02) ** tt
    assert torch.equal(so3.score_so3(sigma, T(g["score_q"]), 2000), T(g["score"]))


def test_igso3_derivative_matches_autograd():
    """bioemu/tests/test_so3_utils.py:333-391: digso3/dlog vs autograd of the series, atol=rtol=1e-3."""
    om = torch.linspace(0.05, np.pi - 0.05, 50, dtype=torch.float64, requires_grad=True)
    sg = torch.full_like(om, 0.5)
    lg = torch.arange(501)
    f = so3.igso3_expansion(om, sg, lg)
    (df,) = torch.autograd.grad(f.sum(), om)
    assert torch.allclose(so3.digso3_expansion(om.detach(), sg, lg), df, atol=1e-3, rtol=1e-3)
    (dl,) = torch.autograd.grad(torch.log(torch.abs(so3.igso3_expansion(om, sg, lg)) + 1e-7).sum(), om)
    assert torch.allclose(so3.dlog_igso3_expansion(om.detach(), sg, lg), dl, atol=1e-3, rtol=1e-3)


def test_so3_tables_and_sampling_bit_exact():
    from oracle.gen_golden import FULL_ROWS, FULL_SDE, SMALL_SDE

    g = load_golden("so3_tables.npz")
    tab = so3.SO3Tables(**SMALL_SDE)
    assert torch.equal(tab.sigma_grid, T(g["small_sigma_grid"]))
    assert torch.equal(tab.omega_grid, T(g["small_omega_grid"]))
    assert torch.equal(tab.cdf_igso3, T(g["small_cdf_igso3"]))
    assert torch.equal(tab.cdf_uso3, T(g["small_cdf_uso3"]))
    assert torch.equal(tab.score_scaling, T(g["small_score_scaling"]))
    # full-size rows (l_max 2000, num_omega 2000) of config.yaml:23-35
    full_grid = tab.marginal_std(torch.linspace(FULL_SDE["eps_t"], 1.0, FULL_SDE["num_sigma"]))
    assert torch.equal(full_grid, T(g["full_sigma_grid"]))
    sel = full_grid[FULL_ROWS]
    om, cdf = so3.build_cdf_table(sel, 2000, 3, 2000, 1e-7)
    assert torch.equal(om, T(g["full_omega_grid"])) and torch.equal(cdf, T(g["full_cdf_igso3_rows"]))
    assert torch.equal(so3.build_cdf_table(sel, 2000, 3, None, 1e-7)[1], T(g["full_cdf_uso3"]))
    assert torch.equal(so3.build_score_scaling(sel, 2000, 3, 2000, 1e-7), T(g["full_score_scaling_rows"]))
    # SURVEY Appendix A known answers measured on the reference's full tables
    ka = {0: 43.875389, 1: 43.667198, 100: 27.265198, 499: 4.0615740, 900: 0.27267385, 999: 0.011169741}
    sc = so3.build_score_scaling(sel, 2000, 3, 2000, 1e-7)
    for r, val in ka.items():
        assert abs(sc[FULL_ROWS.index(r)].item() - val) <= 2e-7 * max(1, val) * 4
    # sampling: explicit noise == seeded global RNG in the reference's order
    n = len(g["prior_u"])
    idx0 = torch.zeros(n, dtype=torch.long)
    pr = so3.sample_rotations(tab.cdf_uso3, tab.omega_grid, idx0, T(g["prior_normals"]), T(g["prior_u"]), None)
    assert torch.equal(pr.squeeze(-3), T(g["prior"]))
    torch.manual_seed(21)
    assert torch.equal(tab.prior(n), T(g["prior"]))
    t = T(g["marg_t"])
    assert torch.equal(torch.bucketize(tab.marginal_std(t), tab.sigma_grid), T(g["marg_sigma_idx"]))
    got = tab.sample_marginal(T(g["prior"]), t, T(g["marg_normals"]), T(g["marg_u"]))
    assert torch.equal(got, T(g["marg"]))
    assert torch.equal(tab.score_scaling_at(t), T(g["score_scaling_at_t"]))
    assert torch.equal(tab.beta(t), T(g["beta_at_t"]))


def test_schedule_scalars_bit_exact():
    """alpha/std/lambda/h/t_lambda/beta for the shipped dpm(50)/heun(100)/em(200) schedules, and the
    SURVEY Appendix A known-answer rows."""
    from oracle.gen_golden import SMALL_SDE

    g = load_golden("schedules.npz")
    cols = list(g["columns"])
    r3 = samplers.CosineVP(0.008)
    tab = so3.SO3Tables(**SMALL_SDE)
    for name, steps in (("dpm", 50), ("heun", 100), ("em", 200)):
        ts = torch.linspace(0.99, 0.001, steps + 1)
        dts = torch.diff(ts)
        ref = g[name]
        for i in range(steps):
            t = torch.full((1,), ts[i].item())
            tn = t + dts[i]
            lam = torch.log(r3.alpha(t) / r3.std(t))
            lam_n = torch.log(r3.alpha(tn) / r3.std(tn))
            tl = torch.full((1,), r3.t_from_lambda((lam + lam_n) / 2).item())
            got = dict(t=t, t_next=tn, dt=dts[i], alpha_t=r3.alpha(t), std_t=r3.std(t), alpha_next=r3.alpha(tn),
                       std_next=r3.std(tn), lambda_t=lam, h=lam_n - lam, t_lambda=tl, alpha_lambda=r3.alpha(tl),
                       std_lambda=r3.std(tl), beta_t=r3.beta(t), beta_lambda=r3.beta(tl),
                       so3_sigma_t=tab.marginal_std(t), so3_g_t=tab.beta(t), so3_g_lambda=tab.beta(tl),
                       score_scaling_t=tab.score_scaling_at(t), score_scaling_lambda=tab.score_scaling_at(tl))
            for c, col in enumerate(cols):
                assert float(got[col].item() if torch.is_tensor(got[col]) else got[col]) == ref[i, c], (name, i, col)
    d = g["dpm"]
    ka = {0: (0.990000010, 0.015583814, 0.999878585, -4.161400795, 1.091890574, 0.982740402, 199.984634399),
          25: (0.495500028, 0.707712471, 0.706500590, 0.001713833, 0.061700039, 0.485605031, 3.111807108),
          49: (0.020780001, 0.999072134, 0.043068510, 3.144034863, 1.903917789, 0.005340052, 0.139872015)}
    for i, (t, a, s, lam, h, tl, beta) in ka.items():
        row = d[i]
        got = (row[0], row[3], row[4], row[7], row[8], row[9], row[12])
        assert np.allclose(got, (t, a, s, lam, h, tl, beta), rtol=2e-6, atol=2e-9)


def test_relative_position_bucket_known_answer():
    """SURVEY Appendix C probe of models.py:94-125."""
    rel = torch.arange(40)
    exp = list(range(16)) + [16, 16, 16, 17, 17, 18, 18, 18, 19, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21, 22, 22, 22, 22, 22]
    assert relative_position_bucket(rel, 64, 128).tolist() == exp
    assert relative_position_bucket(-torch.arange(1, 17), 64, 128).tolist() == list(range(33, 49))
    b = relative_position_bucket(torch.arange(-600, 600), 64, 128)
    assert b.min() == 0 and b.max() == 63


def _pairs(pair_flat, lengths):
    out, o = [], 0
    for n in lengths:
        out.append(pair_flat[o:o + n * n].reshape(n, n, -1))
        o += n * n
    return out


def test_score_model_reference_golden():
    """The reference's own golden vector (bioemu/tests/test_models.py, expected.npz), atol 1e-5."""
    g = load_golden("score_model_tiny.npz")
    cfg = yaml.safe_load(str(g["cfg_json"]))
    m = ScoreModelOracle(_sd(g, "sd::"), num_heads=cfg["num_heads"], num_buckets=cfg["num_buckets"],
                         max_distance=cfg["max_distance_relative"])
    lengths = [10, 10]
    m.set_context(T(g["single"]), _pairs(T(g["pair"]), lengths), lengths)
    with torch.no_grad():
        p, r = m(T(g["in_pos"]), T(g["in_rot"]), T(g["t"])[:2])
    assert np.allclose(p.numpy(), g["expected_pos"], atol=1e-5)
    assert np.allclose(r.numpy(), g["expected_rot"], atol=1e-5)
    assert torch.equal(p, T(g["out_pos"])) and torch.equal(r, T(g["out_rot"]))


def test_score_model_small_ragged_masked():
    g = load_golden("score_model_small.npz")
    cfg = yaml.safe_load(str(g["cfg_json"]))
    lengths = g["lengths"].tolist()
    m = ScoreModelOracle(_sd(g, "sd::"), num_heads=cfg["num_heads"])
    pairs = _pairs(T(g["pair"]), lengths)
    with torch.no_grad():
        m.set_context(T(g["single"]), pairs, lengths)
        p, r = m(T(g["in_pos"]), T(g["in_rot"]), T(g["t"]))
        m.set_context(T(g["single"]), pairs, lengths, pos_is_known=T(g["known"]))
        pk, rk = m(T(g["in_pos"]), T(g["in_rot"]), T(g["t"]))
    assert torch.equal(p, T(g["out_pos"])) and torch.equal(r, T(g["out_rot"]))
    assert torch.equal(pk, T(g["out_pos_known"])) and torch.equal(rk, T(g["out_rot_known"]))


def test_sampler_trajectories_bit_exact():
    from oracle.gen_golden import SMALL_SDE

    g = load_golden("trajectories.npz")
    cfg = yaml.safe_load(str(g["cfg_json"]))
    L, B = int(g["L"]), int(g["B"])
    lengths = [L] * B
    single = T(g["single"]).repeat(B, 1)
    pairs = [T(g["pair"])] * B
    m = ScoreModelOracle(_sd(g, "sd::"), num_heads=cfg["num_heads"]).set_context(single, pairs, lengths)
    fm = ScoreModelOracle(_sd(g, "ft::"), num_heads=cfg["num_heads"]).set_context(single, pairs, lengths)
    tab, r3 = so3.SO3Tables(**SMALL_SDE), samplers.CosineVP(0.008)
    with torch.no_grad():
        torch.manual_seed(int(g["dpm_seed"]))
        p, r = samplers.dpm_solver(m, lengths, r3, tab, int(g["dpm_steps"]), 0.99, 0.001)
        assert torch.equal(p, T(g["dpm_pos"])) and torch.equal(r, T(g["dpm_rot"]))
        torch.manual_seed(int(g["em_seed"]))
        p, r = samplers.euler_maruyama(m, lengths, r3, tab, int(g["em_steps"]), 0.99, 0.001)
        assert torch.equal(p, T(g["em_pos"])) and torch.equal(r, T(g["em_rot"]))
        torch.manual_seed(int(g["heun_seed"]))
        p, r = samplers.heun(m, lengths, r3, tab, int(g["heun_steps"]), 0.99, 0.001, 0.5)
        assert torch.equal(p, T(g["heun_pos"])) and torch.equal(r, T(g["heun_rot"]))
        torch.manual_seed(int(g["emft_seed"]))
        path = samplers.euler_maruyama(m, lengths, r3, tab, int(g["emft_steps"]), 0.99, 0.001, finetune_fn=fm)
        assert torch.equal(torch.stack(path.pos), T(g["emft_pos"]))
        assert torch.equal(torch.stack(path.rot), T(g["emft_rot"]))
        assert torch.equal(path.us["pos"], T(g["emft_us_pos"]))
        assert torch.equal(path.us["node_orientations"], T(g["emft_us_rot"]))
        assert torch.equal(path.dWs["pos"], T(g["emft_dWs_pos"]))
        assert torch.equal(path.dWs["node_orientations"], T(g["emft_dWs_rot"]))
        assert torch.equal(path.timesteps, T(g["emft_timesteps"]))
        # Heun fine-tune variant: the reference's `batches` all alias its final state (recorded in the fixture)
        assert int(g["heunft_aliased"]) == 1
        torch.manual_seed(int(g["heunft_seed"]))
        path = samplers.heun_finetune(m, fm, lengths, r3, tab, int(g["heunft_steps"]), 0.99, 0.001, 0.5)
        assert torch.equal(path.pos[-1], T(g["heunft_pos"])) and torch.equal(path.rot[-1], T(g["heunft_rot"]))
        assert torch.equal(path.us["pos"], T(g["heunft_us_pos"])) and torch.equal(path.us["node_orientations"], T(g["heunft_us_rot"]))
        assert torch.equal(path.dWs["pos"], T(g["heunft_dWs_pos"])) and torch.equal(path.dWs["node_orientations"], T(g["heunft_dWs_rot"]))


def test_analytic_score_moments():
    """bioemu/tests/test_denoiser.py logic (fork kwarg names), oracle dpm sampler, tol 1e-1; the
    reference's own result on the same seed is stored in analytic_denoise.npz for comparison."""
    g = load_golden("analytic_denoise.npz")
    torch.manual_seed(1)
    bs = 1000
    x0_mean, x0_std = torch.tensor(-3.0), torch.tensor(4.3)
    r3 = samplers.CosineVP()
    tab = so3.SO3Tables(num_sigma=10)

    def score_fn(pos, rot, t):
        a, s = r3.alpha(t)[:, None], r3.std(t)[:, None]
        x0 = (x0_mean * s**2 + pos * a * x0_std**2) / (s**2 + a**2 * x0_std**2)
        return (x0 * a - pos) / s, tab.compute_score(so3.rotmat_to_rotvec(rot), t)

    # one graph of 1000 "residues" with 1-D positions, as the reference test builds it
    init = (torch.randn(bs, 1), tab.prior(bs))
    # the samplers draw the prior themselves; emulate the reference order: data batch first, then prior
    p, r = samplers.dpm_solver(score_fn, [bs], r3, tab, 200, 0.99, 0.001,
                               init=(torch.randn(bs, 1), tab.prior(bs)))
    assert torch.isclose(p.mean(), x0_mean, rtol=1e-1, atol=1e-1)
    assert torch.isclose(p.std(), x0_std, rtol=1e-1, atol=1e-1)
    assert torch.allclose(r.mean(dim=0), torch.eye(3), atol=1e-1)
    assert torch.allclose(r.std(dim=0), torch.zeros(3, 3), atol=1e-1)
    assert abs(float(g["dpm_pos_mean"]) - x0_mean.item()) < 0.5 and abs(float(g["heun_pos_std"]) - 4.3) < 0.5


def test_toy_layer_bit_exact():
    """oracle/toy.py vs the unmodified se3diff/{models,train,finetune}.py (+ bioemu/ppft.py) on recorded seeds."""
    from oracle import toy
    from oracle.gen_golden import SMALL_SDE

    g = load_golden("toy.npz")
    tab = so3.SO3Tables(**SMALL_SDE)
    net = toy.ScoreNetOracle(_sd(g, "net::"))
    ctrl = toy.ScoreNetOracle(_sd(g, "ctrl::"))
    mus, sigmas, weights, h_stars = T(g["mus"]), T(g["sigmas"]), T(g["weights"]), T(g["h_stars"])
    with torch.no_grad():
        assert torch.equal(net(T(g["fw_x"]), T(g["fw_t"])), T(g["fw_out"]))
        om, pdf = toy.igso3_mixture_marginal_pdf(mus, sigmas, weights, l_max=200, num_points=64)
        assert torch.equal(om, T(g["mix_omega"])) and torch.equal(pdf, T(g["mix_pdf"]))
        assert torch.equal(toy.assign_igso3(T(g["assign_x0"]), mus, sigmas, weights, l_max=200), T(g["assign_hs"]))
        torch.manual_seed(int(g["mixsample_seed"]))
        assert torch.equal(toy.sample_multiple_igso3(tab, mus, sigmas, weights, 32), T(g["mixsample"]))
    torch.manual_seed(int(g["train_seed"]))
    loss = toy.compute_train_loss(tab, net, mus, sigmas, weights, batch_size=64)
    assert torch.equal(loss.detach(), T(g["train_loss"]))
    grads = torch.autograd.grad(loss, net.parameters())
    assert torch.allclose(torch.stack([x.norm() for x in grads]), T(g["train_grad_norms"]), rtol=1e-5, atol=1e-8)
    torch.manual_seed(int(g["rev_seed"]))
    xs, ts = toy.reverse_diffusion(tab, net, 16, 8)
    assert torch.equal(xs, T(g["rev_xs"])) and torch.equal(ts, T(g["rev_ts"]))
    torch.manual_seed(int(g["revft_seed"]))
    xs, ts, us, dWs = toy.reverse_diffusion(tab, net, 16, 6, finetune_model=ctrl)
    assert torch.equal(xs, T(g["revft_xs"])) and torch.equal(us, T(g["revft_us"])) and torch.equal(dWs, T(g["revft_dWs"]))
    torch.manual_seed(int(g["ft_seed"]))
    loss = toy.compute_finetune_loss(tab, net, ctrl, mus, sigmas, h_stars, batch_size=16, num_steps=6, l_max=200)
    assert torch.equal(loss.detach(), T(g["ft_loss"]))
    grads = torch.autograd.grad(loss, ctrl.parameters())
    assert torch.allclose(torch.stack([x.norm() for x in grads]), T(g["ft_grad_norms"]), rtol=1e-4, atol=1e-9)
    assert torch.allclose(grads[-1], T(g["ft_grad_last"]), rtol=1e-4, atol=1e-9)
