"""TEST INFRASTRUCTURE -- mint tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (where /root/reference is mounted):

    python -m oracle.gen_golden

Every array written here is an *output of the reference's own code* (imported in place through
oracle/ref_harness.py); the inputs are stored alongside so that the fixtures are self-contained on
the GPU box, where the reference tree does not exist.  Only data is committed, never reference code.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch
import yaml

from . import ref_harness
from .score_model import make_edge_index  # noqa: F401

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ADV_ANGLES = [0.0, 1e-9, 1e-8, 1e-7, 2e-7, 1e-4, 1.0, 3.0, np.pi - 0.0101, np.pi - 0.0100, np.pi - 1e-3,
              np.pi - 1e-6, np.pi]


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def so3_maps(ns):
    """K2: exp / log / compose / relative log / quaternion maps on random + adversarial angles."""
    R = ns.so3_sde
    g = torch.Generator().manual_seed(11)
    n = 512
    ax = torch.randn(n, 3, generator=g, dtype=torch.float64)
    ax /= ax.norm(dim=-1, keepdim=True)
    ang = torch.rand(n, generator=g, dtype=torch.float64) * np.pi
    adv = torch.tensor(ADV_ANGLES, dtype=torch.float64)
    ang[: len(adv) * 8] = adv.repeat(8)
    # axis-aligned axes for the first block of each adversarial angle (sign logic of the pi branch)
    for k in range(len(adv)):
        ax[k] = torch.tensor([[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [-1.0, 0, 0]][k % 4], dtype=torch.float64)
    out = {}
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        v = (ax * ang[:, None]).to(dt)
        rm = R.rotvec_to_rotmat(v)
        w = v.flip(0) * 0.37
        out[f"v_{name}"] = v
        out[f"exp_{name}"] = rm
        out[f"log_{name}"] = R.rotmat_to_rotvec(rm)
        out[f"w_{name}"] = w
        out[f"compose_{name}"] = R.apply_rotvec_to_rotmat(rm, w)
        out[f"rel_log_{name}"] = R.rot_vf(rm, R.apply_rotvec_to_rotmat(rm, w))
        a, s, c = R.angle_from_rotmat(rm)
        out[f"angle_{name}"] = a
        out[f"geodesic_t_{name}"] = R.geodesic_t(0.3, rm.flip(0), rm)
        out[f"scale_{name}"] = R.scale_rotmat(rm, torch.full((n, 1), 0.5, dtype=dt))
        q = torch.randn(n, 4, generator=g, dtype=torch.float64).to(dt)
        q = q / q.norm(dim=-1, keepdim=True)
        out[f"quat_{name}"] = q
        out[f"quat_rotvec_{name}"] = R.rotquat_to_rotvec(q)
        out[f"quat_rotmat_{name}"] = R.rotquat_to_rotmat(q)
    np.savez_compressed(os.path.join(OUT, "so3_maps.npz"), **_np(out))


def igso3_series(ns):
    """K1a: series f, df, dlog, marginal pdf, score on an (omega, sigma) grid."""
    R = ns.so3_sde
    g = torch.Generator().manual_seed(12)
    n = 400
    om = torch.rand(n, generator=g, dtype=torch.float64) ** 2 * np.pi
    om[:6] = torch.tensor([0.0, 1e-8, 1e-7, 2e-7, np.pi, 3.0], dtype=torch.float64)
    t = torch.rand(n, generator=g, dtype=torch.float64) * 0.989 + 0.001
    sg = 0.02 * (2.33 / 0.02) ** t
    out = {"omega": om, "sigma": sg}
    for l_max in (2000, 500):
        lg = torch.arange(l_max + 1)
        for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
            o, s = om.to(dt), sg.to(dt)
            out[f"f_{name}_l{l_max}"] = R.igso3_expansion(o, s, lg)
            out[f"df_{name}_l{l_max}"] = R.digso3_expansion(o, s, lg)
            out[f"dlog_{name}_l{l_max}"] = R.dlog_igso3_expansion(o, s, lg)
    lg = torch.arange(1000)  # se3diff/train.py:90 uses arange(l_max) without +1
    om0 = om.flip(0).clone()
    out["omega0"] = om0
    for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        out[f"marginal_{name}"] = R.igso3_marginal_pdf(om.to(dt), om0.to(dt), sg.to(dt), lg)
    sde = R.DiGSO3SDE(eps_t=0.001, num_sigma=8, num_omega=16, l_max=2000, sigma_min=0.02, sigma_max=2.33)
    q = torch.randn(n, 3, generator=g) * torch.rand(n, 1, generator=g) * 2.0
    tt = t.float()
    out["score_q"] = q
    out["score_t"] = tt
    out["score"] = sde.compute_score(q, tt)
    np.savez_compressed(os.path.join(OUT, "igso3_series.npz"), **_np(out))


SMALL_SDE = dict(eps_t=0.001, num_sigma=64, num_omega=256, omega_exponent=3, l_max=256, sigma_min=0.02,
                 sigma_max=2.33, tol=1e-7)
FULL_SDE = dict(eps_t=0.001, num_sigma=1000, num_omega=2000, omega_exponent=3, l_max=2000, sigma_min=0.02,
                sigma_max=2.33, tol=1e-7)
FULL_ROWS = [0, 1, 100, 250, 499, 750, 900, 998, 999]


def so3_tables(ns):
    """K1b/K1c: a complete small table set, selected rows of the full-size tables, inverse-CDF draws."""
    R = ns.so3_sde
    sde = R.DiGSO3SDE(**SMALL_SDE)
    out = {
        "small_sigma_grid": sde.igso3.sigma_grid, "small_omega_grid": sde.igso3.omega_grid,
        "small_cdf_igso3": sde.igso3.cdf_igso3, "small_cdf_uso3": sde.uso3.cdf_igso3,
        "small_score_scaling": sde.score_function.score_scaling,
    }
    # full-size rows: the reference classes evaluated on a sub-grid of sigma (rows are independent)
    full_grid = sde._marginal_std(torch.linspace(FULL_SDE["eps_t"], 1.0, FULL_SDE["num_sigma"]))
    sel = full_grid[FULL_ROWS]
    ig = R.SampleIGSO3(num_omega=2000, sigma_grid=sel, omega_exponent=3, l_max=2000, tol=1e-7)
    us = R.SampleUSO3(num_omega=2000, sigma_grid=sel, omega_exponent=3, tol=1e-7)
    sc = R.ScoreSO3(num_omega=2000, sigma_grid=sel, omega_exponent=3, l_max=2000, tol=1e-7)
    out.update(full_rows=np.asarray(FULL_ROWS), full_sigma_grid=full_grid, full_omega_grid=ig.omega_grid,
               full_cdf_igso3_rows=ig.cdf_igso3, full_cdf_uso3=us.cdf_igso3,
               full_score_scaling_rows=sc.score_scaling)
    # sampling with explicit RNG (seeded global generator, reference order)
    n = 300
    torch.manual_seed(21)
    prior = sde.prior_sampling((n, 3, 3))
    torch.manual_seed(21)
    normals, u = torch.randn(n, 1, 3), torch.rand(n, 1)
    t = torch.linspace(0.001, 1.0, n)
    torch.manual_seed(22)
    marg = sde.sample_marginal(prior, t)
    torch.manual_seed(22)
    normals2, u2 = torch.randn(n, 1, 3), torch.rand(n, 1)
    out.update(prior=prior, prior_normals=normals, prior_u=u, marg_t=t, marg=marg, marg_normals=normals2,
               marg_u=u2, marg_sigma_idx=sde.igso3.get_sigma_idx(sde._marginal_std(t)),
               score_scaling_at_t=sde.get_score_scaling(t), beta_at_t=sde.beta(t), sigma_at_t=sde._marginal_std(t))
    np.savez_compressed(os.path.join(OUT, "so3_tables.npz"), **_np(out))


def schedules(ns):
    """K3 scalars for the three shipped schedules (config/denoiser/*.yaml), straight from the
    reference SDE objects in fp32: alpha, std, beta(t), SO(3) sigma/g, and dpm's lambda/h/t_lambda."""
    r3 = ns.sde_lib.CosineVPSDE(s=0.008)
    so3 = ns.so3_sde.DiGSO3SDE(**SMALL_SDE)
    out = {}
    for name, steps in (("dpm", 50), ("heun", 100), ("em", 200)):
        ts = torch.linspace(0.99, 0.001, steps + 1)
        dts = torch.diff(ts)
        rows = []
        for i in range(steps):
            t = torch.full((1,), ts[i].item())
            t_next = t + dts[i]
            x = torch.ones(1, 3)
            bi = torch.zeros(1, dtype=torch.long)
            a_t, s_t = r3.mean_coeff_and_std(x=x, t=t, batch_idx=bi)
            a_n, s_n = r3.mean_coeff_and_std(x=x, t=t_next, batch_idx=bi)
            lam, lam_n = torch.log(a_t / s_t), torch.log(a_n / s_n)
            t_lam = ns.denoiser._t_from_lambda(r3, (lam + lam_n) / 2)[0][0].item()
            tl = torch.full((1,), t_lam)
            a_l, s_l = r3.mean_coeff_and_std(x=x, t=tl, batch_idx=bi)
            rows.append([t.item(), t_next.item(), dts[i].item(), a_t[0, 0].item(), s_t[0, 0].item(),
                         a_n[0, 0].item(), s_n[0, 0].item(), lam[0, 0].item(), (lam_n - lam)[0, 0].item(),
                         t_lam, a_l[0, 0].item(), s_l[0, 0].item(), r3.beta(t).item(), r3.beta(tl).item(),
                         so3._marginal_std(t).item(), so3.beta(t).item(), so3.beta(tl).item(),
                         so3.get_score_scaling(t).item(), so3.get_score_scaling(tl).item()])
        out[name] = np.asarray(rows, dtype=np.float64)
    out["columns"] = np.asarray(["t", "t_next", "dt", "alpha_t", "std_t", "alpha_next", "std_next", "lambda_t", "h",
                                 "t_lambda", "alpha_lambda", "std_lambda", "beta_t", "beta_lambda", "so3_sigma_t",
                                 "so3_g_t", "so3_g_lambda", "score_scaling_t", "score_scaling_lambda"])
    np.savez_compressed(os.path.join(OUT, "schedules.npz"), **out)


def score_model_tiny(ns):
    """The reference's own golden (bioemu/tests/test_models.py + expected.npz + state_dict.ptkeep),
    re-run through the unmodified reference model and checked against expected.npz here."""
    import sys

    sys.path.insert(0, os.path.join(ns.root, "bioemu"))
    from tests.conftest import get_dicts

    tests = os.path.join(ns.root, "bioemu", "tests")
    cfg = yaml.safe_load(open(os.path.join(tests, "tiny_config.yaml")))["score_model"]
    cfg.pop("_target_")
    m = ns.models.DiGConditionalScoreModel(**cfg)
    sd = torch.load(os.path.join(tests, "state_dict.ptkeep"))
    m.load_state_dict(sd)
    m.eval()
    batch = ns.Batch.from_data_list([ns.chemgraph.ChemGraph(**d) for d in get_dicts()])
    with torch.no_grad():
        out = m(batch, t=torch.tensor([0.0] * 10))
    exp = np.load(os.path.join(tests, "expected.npz"))
    for k in ("pos", "node_orientations"):
        assert np.allclose(out[k].numpy(), exp[k], atol=1e-5), k
    save = {"sd::" + k: v for k, v in sd.items()}
    save.update(cfg_json=np.asarray(yaml.safe_dump(cfg)), in_pos=batch.pos, in_rot=batch.node_orientations,
                single=batch.single_embeds, pair=batch.pair_embeds, edge_index=batch.edge_index,
                batch_idx=batch.batch, t=torch.tensor([0.0] * 10), out_pos=out["pos"],
                out_rot=out["node_orientations"], expected_pos=exp["pos"], expected_rot=exp["node_orientations"])
    np.savez_compressed(os.path.join(OUT, "score_model_tiny.npz"), **_np(save))


SMALL_MODEL = dict(dim_model=64, dim_pair=32, num_layers=2, num_heads=4, dim_single_rep=16, dim_hidden=128,
                   num_buckets=64, max_distance_relative=128, dropout=0.1)


def _small_model(ns, seed):
    torch.manual_seed(seed)
    m = ns.models.DiGConditionalScoreModel(**SMALL_MODEL).eval()
    with torch.no_grad():  # make LayerNorm affine / biases non-trivial
        for name, p in m.named_parameters():
            if p.ndim == 1 and p.numel() > 0:
                p.add_(0.1 * torch.randn_like(p))
    return m


def score_model_small(ns):
    """Multi-head, 2-layer, ragged lengths, pos_is_known mask, physical-scale frames: everything the
    tiny golden leaves unpinned."""
    m = _small_model(ns, 31)
    g = torch.Generator().manual_seed(32)
    lengths = [7, 12, 5]
    single = [torch.randn(n, 384, generator=g) for n in lengths]
    pair = [torch.randn(n, n, 128, generator=g) for n in lengths]
    N = sum(lengths)
    pos = torch.randn(N, 3, generator=g) * 1.5
    rot = ns.so3_sde.rotvec_to_rotmat(torch.randn(N, 3, generator=g))
    known = (torch.rand(N, generator=g) > 0.2).float()
    t = torch.tensor([0.3, 0.7, 0.05])
    save = {"sd::" + k: v for k, v in m.state_dict().items()}
    with torch.no_grad():
        o1 = m(ref_harness.make_batch(ns, single, pair, lengths, pos, rot, extra={"pos_is_known": known}), t)
        o2 = m(ref_harness.make_batch(ns, single, pair, lengths, pos, rot), t)
    save.update(lengths=np.asarray(lengths), single=torch.cat(single), pair=torch.cat([p.reshape(-1, 128) for p in pair]),
                in_pos=pos, in_rot=rot, known=known, t=t, out_pos_known=o1["pos"], out_rot_known=o1["node_orientations"],
                out_pos=o2["pos"], out_rot=o2["node_orientations"], cfg_json=np.asarray(yaml.safe_dump(SMALL_MODEL)))
    # gradients of a fixed linear functional of the outputs (the differentiable forward the fine-tune step needs, finetune.py:338-393)
    wp, wr = torch.randn(N, 3, generator=g), torch.randn(N, 3, generator=g)
    o3 = m(ref_harness.make_batch(ns, single, pair, lengths, pos, rot, extra={"pos_is_known": known}), t)
    loss = (o3["pos"] * wp).sum() + (o3["node_orientations"] * wr).sum()
    named = [(k, p) for k, p in m.named_parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, [p for _, p in named])
    save.update(grad_wp=wp, grad_wr=wr, grad_loss=loss.detach(), grad_names=np.asarray([k for k, _ in named]),
                grad_norms=torch.stack([gr.norm() for gr in grads]),
                **{"grad::" + k: gr for (k, _), gr in zip(named, grads) if gr.numel() <= 4096})
    np.savez_compressed(os.path.join(OUT, "score_model_small.npz"), **_np(save))


def trajectories(ns):
    """Full reference sampler runs (dpm 10 steps, EM 12, Heun 8, EM-finetune 6) with the small model,
    small SO(3) tables, B=3 copies of one 11-residue sequence, seeds recorded."""
    m = _small_model(ns, 41)
    fm = _small_model(ns, 42)
    with torch.no_grad():
        for p in fm.parameters():
            p.mul_(0.3)
    g = torch.Generator().manual_seed(43)
    L, B = 11, 3
    single = [torch.randn(L, 384, generator=g)] * B
    pair = [torch.randn(L, L, 128, generator=g)] * B
    lengths = [L] * B
    so3 = ns.so3_sde.DiGSO3SDE(**SMALL_SDE)
    r3 = ns.sde_lib.CosineVPSDE(s=0.008)
    sdes = {"node_orientations": so3, "pos": r3}

    def mk():
        return ref_harness.make_batch(ns, single, pair, lengths)

    save = {"sd::" + k: v for k, v in m.state_dict().items()}
    save.update({"ft::" + k: v for k, v in fm.state_dict().items()})
    save.update(single=single[0], pair=pair[0], L=L, B=B, cfg_json=np.asarray(yaml.safe_dump(SMALL_MODEL)))
    D = ns.denoiser
    kw = dict(max_t=0.99, min_t=0.001, device="cpu")
    with torch.no_grad():
        torch.manual_seed(51)
        o = D.dpm_solver(batch=mk(), sdes=sdes, score_model=m, num_steps=10, **kw)
        save.update(dpm_pos=o.pos, dpm_rot=o.node_orientations, dpm_seed=51, dpm_steps=10)
        torch.manual_seed(52)
        o = D.euler_maruyama_predictor(batch=mk(), sdes=sdes, score_model=m, num_steps=12, **kw)
        save.update(em_pos=o.pos, em_rot=o.node_orientations, em_seed=52, em_steps=12)
        torch.manual_seed(53)
        o = D.heun_denoiser(batch=mk(), sdes=sdes, score_model=m, num_steps=8, noise=0.5, **kw)
        save.update(heun_pos=o.pos, heun_rot=o.node_orientations, heun_seed=53, heun_steps=8)
        torch.manual_seed(54)
        o = D.euler_maruyama_predictor_finetune(batch=mk(), sdes=sdes, score_model=m, finetune_model=fm,
                                                num_steps=6, **kw)
        save.update(emft_pos=torch.stack([b.pos for b in o.batches]),
                    emft_rot=torch.stack([b.node_orientations for b in o.batches]),
                    emft_us_pos=o.us_batch["pos"], emft_us_rot=o.us_batch["node_orientations"],
                    emft_dWs_pos=o.dWs_batch["pos"], emft_dWs_rot=o.dWs_batch["node_orientations"],
                    emft_timesteps=o.timesteps, emft_seed=54, emft_steps=6)
        torch.manual_seed(55)
        o = D.heun_denoiser_finetune(batch=mk(), sdes=sdes, score_model=m, finetune_model=fm, num_steps=5, noise=0.5, **kw)
        # every entry of o.batches is the same in-place-mutated object (denoiser.py:518,596): only the final state is meaningful
        save.update(heunft_pos=o.batches[-1].pos, heunft_rot=o.batches[-1].node_orientations,
                    heunft_us_pos=o.us_batch["pos"], heunft_us_rot=o.us_batch["node_orientations"],
                    heunft_dWs_pos=o.dWs_batch["pos"], heunft_dWs_rot=o.dWs_batch["node_orientations"],
                    heunft_aliased=int(all(bb is o.batches[0] for bb in o.batches)), heunft_seed=55, heunft_steps=5)
    np.savez_compressed(os.path.join(OUT, "trajectories.npz"), **_np(save))


def analytic_denoise(ns):
    """bioemu/tests/test_denoiser.py with the fork's kwarg names (num_steps/min_t): analytic Gaussian
    + IGSO3 scores; stores the reference's final moments for dpm and heun (tolerance 1e-1 there)."""
    torch.manual_seed(1)
    steps, bs = 200, 1000
    x0_mean, x0_std = torch.tensor(-3.0), torch.tensor(4.3)
    r3 = ns.sde_lib.CosineVPSDE()
    so3 = ns.so3_sde.DiGSO3SDE(num_sigma=10)
    sdes = {"pos": r3, "node_orientations": so3}

    def score_fn(x, t):
        a_t, s_t = r3.marginal_prob(x=torch.ones_like(x.pos), t=t)
        x0 = (x0_mean * s_t**2 + x.pos * a_t * x0_std**2) / (s_t**2 + a_t**2 * x0_std**2)
        return x.replace(pos=(x0 * a_t - x.pos) / s_t,
                         node_orientations=so3.compute_score(ns.so3_sde.rotmat_to_rotvec(x.node_orientations), t))

    out = {}
    for name, solver, kw in (("dpm", ns.denoiser.dpm_solver, {}), ("heun", ns.denoiser.heun_denoiser, {"noise": 0.5})):
        data = ns.Batch.from_data_list([ns.chemgraph.ChemGraph(pos=torch.randn(bs, 1),
                                                               node_orientations=so3.prior_sampling((bs,)))])
        s = solver(sdes=sdes, batch=data, num_steps=steps, score_model=score_fn, max_t=0.99, min_t=0.001,
                   device=torch.device("cpu"), **kw)
        out[f"{name}_pos_mean"] = s.pos.mean()
        out[f"{name}_pos_std"] = s.pos.std()
        out[f"{name}_rot_mean"] = s.node_orientations.mean(dim=0)
        out[f"{name}_rot_std"] = s.node_orientations.std(dim=0)
    np.savez_compressed(os.path.join(OUT, "analytic_denoise.npz"), **_np(out))


def toy(ns):
    """The SO(3)-only toy layer (se3diff/*): ScoreNet forward, mixture sampling / pdf / responsibilities, the DSM loss,
    reverse diffusion with and without control, and the fine-tune loss with its gradient -- all from the unmodified
    reference with recorded seeds, small SO(3) tables."""
    import contextlib
    import io

    M, T, FT = ns.toy_models, ns.toy_train, ns.toy_finetune
    torch.manual_seed(61)
    net = M.ScoreNet()
    torch.manual_seed(62)
    ctrl = M.ScoreNet()
    with torch.no_grad():
        for p in ctrl.parameters():
            p.mul_(0.2)
    sde = M.DiGMixSO3SDE(**SMALL_SDE)
    g = torch.Generator().manual_seed(63)
    K = 3
    mus = ns.so3_sde.rotvec_to_rotmat(torch.randn(K, 3, generator=g))
    sigmas = torch.tensor([0.1, 0.3, 0.8])
    weights = torch.tensor([0.5, 0.3, 0.2])
    h_stars = torch.tensor([0.2, 0.3, 0.5])
    save = {"net::" + k: v for k, v in net.state_dict().items()}
    save.update({"ctrl::" + k: v for k, v in ctrl.state_dict().items()})
    save.update(mus=mus, sigmas=sigmas, weights=weights, h_stars=h_stars)
    # ScoreNet forward on generic + adversarial rotations
    v = torch.randn(40, 3, generator=g)
    v[:4] *= 1e-4
    v[4:8] = v[4:8] / v[4:8].norm(dim=-1, keepdim=True) * (math.pi - 1e-3)
    x = ns.so3_sde.rotvec_to_rotmat(v)
    t = torch.rand(40, generator=g)
    with torch.no_grad():
        save.update(fw_x=x, fw_t=t, fw_out=net(x, t))
        om, pdf = T.igso3_mixture_marginal_pdf(mus, sigmas, weights, l_max=200, num_points=64)
        save.update(mix_omega=om, mix_pdf=pdf)
        x0 = ns.so3_sde.rotvec_to_rotmat(torch.randn(24, 3, generator=g))
        save.update(assign_x0=x0, assign_hs=FT.assign_igso3(x0, mus, sigmas, weights, l_max=200))
        torch.manual_seed(64)
        save.update(mixsample_seed=64, mixsample=sde.sample_multiple_igso3(mus, sigmas, weights, 32))
    torch.manual_seed(65)
    loss = T.compute_train_loss(sde, net, mus, sigmas, weights, batch_size=64)
    grads = torch.autograd.grad(loss, [p for p in net.parameters() if p.requires_grad])
    save.update(train_seed=65, train_loss=loss.detach(), train_grad_norms=torch.stack([gr.norm() for gr in grads]))
    torch.manual_seed(66)
    xs, ts = T.reverse_diffusion(sde, net, device="cpu", batch_size=16, num_steps=8)
    save.update(rev_seed=66, rev_xs=xs, rev_ts=ts)
    torch.manual_seed(67)
    xs, ts, us, dWs = FT.reverse_finetune_diffusion(sde, net, ctrl, device="cpu", batch_size=16, num_steps=6)
    save.update(revft_seed=67, revft_xs=xs, revft_us=us, revft_dWs=dWs)
    torch.manual_seed(68)
    with contextlib.redirect_stdout(io.StringIO()):      # ppft.compute_ev_loss prints its arguments
        loss = FT.compute_finetune_loss(sde, net, ctrl, mus, sigmas, h_stars, device="cpu", batch_size=16, num_steps=6, l_max=200)
    grads = torch.autograd.grad(loss, [p for p in ctrl.parameters() if p.requires_grad])
    save.update(ft_seed=68, ft_loss=loss.detach(), ft_grad_norms=torch.stack([gr.norm() for gr in grads]),
                ft_grad_last=grads[-1])
    np.savez_compressed(os.path.join(OUT, "toy.npz"), **_np(save))


def finetune_step(ns):
    """The fine-tune step (finetune.py:338-393 `_chunk_update`, observables/folding_stability.py:52-101) from the reference's
    own function bodies (executed in memory: their modules need hydra / mdtraj / Bio to import) on a reference rollout."""
    import contextlib
    import io
    from collections import defaultdict

    import torch.nn.functional as F
    from torch_geometric.utils import to_dense_batch

    import bioemu.ppft as ppft

    fs = ref_harness.extract_functions("bioemu/src/bioemu/observables/folding_stability.py",
                                       ["compute_folded_proportion", "compute_dG", "compute_folded_proportion_from_dG"],
                                       {"torch": torch, "F": F, "K_BOLTZMANN": 0.001987203599772605})
    ft = ref_harness.extract_functions("bioemu/src/bioemu/finetune.py", ["_chunk_update"], {
        "torch": torch, "defaultdict": defaultdict, "to_dense_batch": to_dense_batch, "ChemGraph": ns.chemgraph.ChemGraph,
        "DiGConditionalScoreModel": ns.models.DiGConditionalScoreModel, "DeviceLikeType": object,
        "compute_int_dws": ppft.compute_int_dws, "compute_int_u_u_dt": ppft.compute_int_u_u_dt,
        "compute_ev_loss": ppft.compute_ev_loss, "compute_kl_loss": ppft.compute_kl_loss})
    m, fm = _small_model(ns, 41), _small_model(ns, 42)
    with torch.no_grad():
        for p in fm.parameters():
            p.mul_(0.3)
    g = torch.Generator().manual_seed(43)                   # the sequence of the `trajectories` fixture
    L, B, T = 11, 4, 6
    single = [torch.randn(L, 384, generator=g)] * B
    pair = [torch.randn(L, L, 128, generator=g)] * B
    so3 = ns.so3_sde.DiGSO3SDE(**SMALL_SDE)
    sdes = {"node_orientations": so3, "pos": ns.sde_lib.CosineVPSDE(s=0.008)}
    with torch.no_grad():
        torch.manual_seed(71)
        path = ns.denoiser.euler_maruyama_predictor_finetune(batch=ref_harness.make_batch(ns, single, pair, [L] * B), sdes=sdes, score_model=m,
                                                             finetune_model=fm, num_steps=T, max_t=0.99, min_t=0.001, device="cpu")
    batches, timesteps, us_sg, dWs = path
    fields = list(sdes.keys())
    # observable: random-weight frames live at the 100-nm scale, so the sigmoid is given a matching width / threshold
    gg = torch.Generator().manual_seed(72)
    ref_coords = torch.randn(L, 3, generator=gg) * 40.0
    coords = to_dense_batch(batches[-1].pos, batches[-1].batch)[0]
    k_, d0_ = -0.02, 120.0
    p_folded = fs["compute_folded_proportion"](coords, ref_coords, k_, d0_, 1e-7)
    hs = p_folded.unsqueeze(-1)
    h_stars = torch.tensor([0.6])
    dts = torch.diff(timesteps)
    int_sg = sum(ppft.compute_int_u_u_dt(us=us_sg[f].flatten(-2, -1), dts=dts) for f in fields)
    fm.zero_grad()
    micro = 4
    with contextlib.redirect_stdout(io.StringIO()):
        for lo in range(0, T, micro):
            hi = min(lo + micro, T)
            ft["_chunk_update"](batches=batches[lo:hi], timesteps=timesteps[lo:hi], dts=dts[lo:hi], dWs_batch={f: dWs[f][lo:hi] for f in fields},
                                int_u_u_dt_sg=int_sg, hs=hs, h_stars=h_stars, finetune_model=fm, fields=fields, batch_size=B, device="cpu",
                                lambda_=0.1, tol=1e-7)
        ws = torch.ones_like(int_sg)
        val = ppft.compute_ev_loss(ws=ws, hs=hs, h_stars=h_stars, from_int_dws=False, use_stab=False, tol=1e-7) + 0.1 * ppft.compute_kl_loss(
            ws=ws, int_u_u_dt=int_sg, int_u_u_dt_sg=int_sg, from_int_dws=False, use_rloo=False)
    named = [(k, p) for k, p in fm.named_parameters() if p.grad is not None]
    save = dict(L=L, B=B, T=T, seed=71, micro=micro, k=k_, d_0=d0_, ref_coords=ref_coords, coords=coords, p_folded=p_folded,
                dG=fs["compute_dG"](p_folded), p_from_dG=fs["compute_folded_proportion_from_dG"](torch.tensor([-1.0, 0.0, 2.5])),
                h_stars=h_stars, val_loss=val.detach(), final_pos=batches[-1].pos, grad_names=np.asarray([k for k, _ in named]),
                grad_norms=torch.stack([p.grad.norm() for _, p in named]),
                **{"grad::" + k: p.grad for k, p in named if p.numel() <= 4096})
    np.savez_compressed(os.path.join(OUT, "finetune_step.npz"), **_np(save))


def backbone(ns):
    """Frames -> atom37 (convert_chemgraph.py:17-293) from the reference's own function bodies (the module itself needs mdtraj
    / modelcif to import), on a random-walk chain whose sequence holds all twenty residue types."""
    from bioemu.openfold.np import residue_constants
    from bioemu.openfold.utils.rigid_utils import Rigid, Rotation

    cc = ref_harness.extract_functions("bioemu/src/bioemu/convert_chemgraph.py",
                                       ["_torsion_angles_to_frames", "frames_to_atom14_pos", "compute_backbone", "_adjust_oxygen_pos",
                                        "get_atom37_from_frames"],
                                       {"torch": torch, "residue_constants": residue_constants, "Rigid": Rigid, "Rotation": Rotation,
                                        "C_O_BOND_LENGTH": 1.23})
    seq = "ARNDCQEGHILKMFPSTWYVGGAPX"
    g = torch.Generator().manual_seed(81)
    L = len(seq)
    steps = torch.randn(L, 3, generator=g)
    pos = torch.cumsum(3.8 * steps / steps.norm(dim=-1, keepdim=True), dim=0)      # Angstrom, CA spacing 3.8
    rot = ns.so3_sde.rotvec_to_rotmat(torch.randn(L, 3, generator=g))
    atom37, mask, aatype = cc["get_atom37_from_frames"](pos.clone(), rot.clone(), seq)
    np.savez_compressed(os.path.join(OUT, "backbone.npz"), **_np(dict(sequence=np.asarray(seq), pos=pos, rot=rot, atom37=atom37,
                                                                       mask=mask, aatype=aatype)))


CACHE_SDE = dict(eps_t=0.001, num_sigma=8, num_omega=64, omega_exponent=3, l_max=64, sigma_min=0.02, sigma_max=2.33, tol=1e-7)


def so3_cache(ns):
    """The reference's own npz lookup-table cache (so3_sde.py:914-990 `SO3LookupCache.save_cache`, file names at :1098, 1354,
    1607) for a tiny table set: three files written by the unmodified `DiGSO3SDE.__init__` into tests/golden/so3_cache/."""
    import shutil

    d = os.path.join(OUT, "so3_cache")
    shutil.rmtree(d, ignore_errors=True)
    ns.so3_sde.DiGSO3SDE(**CACHE_SDE, cache_dir=d, overwrite_cache=False)
    print("  cache files:", sorted(os.listdir(d)))


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_harness.load()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    every = (so3_maps, igso3_series, so3_tables, schedules, score_model_tiny, score_model_small, trajectories,
             analytic_denoise, toy, finetune_step, backbone, so3_cache)
    only = set(sys.argv[1:])                      # e.g. `python -m oracle.gen_golden toy` regenerates one file
    for fn in every:
        if only and fn.__name__ not in only:
            continue
        print("golden:", fn.__name__, flush=True)
        fn(ns)
    print({f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT)) if os.path.isfile(os.path.join(OUT, f))})


if __name__ == "__main__":
    main()
