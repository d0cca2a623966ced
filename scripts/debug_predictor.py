"""EulerMaruyamaPredictor (GPU class) vs oracle EM on random inputs (developer diagnostics)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from oracle import samplers as osamp, so3 as oso3
from oracle.gen_golden import SMALL_SDE
from se3diff_b200 import sdes as S
from se3diff_b200.denoiser import EulerMaruyamaPredictor
dev = "cuda"
tab, r3o = oso3.SO3Tables(**SMALL_SDE), osamp.CosineVP(0.008)
so3 = S.DiGSO3SDE(**SMALL_SDE); so3.score_function.score_scaling.copy_(tab.score_scaling); so3 = so3.to(dev)
r3 = S.CosineVPSDE(0.008)
B, L = 3, 5
bi = torch.repeat_interleave(torch.arange(B), L)
g = torch.Generator().manual_seed(0)
pos = torch.randn(B * L, 3, generator=g); rot = oso3.rotvec_to_rotmat(torch.randn(B * L, 3, generator=g))
score = torch.randn(B * L, 3, generator=g); u = torch.randn(B * L, 3, generator=g) * 0.1
t = torch.full((B,), 0.6); dt = torch.tensor(0.013)
for kind, x, sde_o, sde_g in (("pos", pos, None, r3), ("rot", rot, None, so3)):
    for nw in (0.0, 1.0):
        eo = osamp.EM(kind, r3o, tab, nw)
        eg = EulerMaruyamaPredictor(corruption=sde_g, noise_weight=nw)
        torch.manual_seed(1); a = eo.forward_step(x, t, dt, bi)
        with S.host_noise():
            torch.manual_seed(1); b = eg.forward_sde_step(x=x.to(dev), t=t.to(dev), dt=dt.to(dev), batch_idx=bi.to(dev))
        print(kind, nw, "forward", [(p.cpu() - q).abs().max().item() for p, q in zip(b, a)])
        d_o = eo.drift_diffusion(x, t, score, bi, u); d_g = eg.reverse_drift_and_diffusion(x=x.to(dev), t=t.to(dev), score=score.to(dev), finetune_score=u.to(dev), batch_idx=bi.to(dev))
        print(kind, nw, "drift", (d_g[0].cpu() - d_o[0]).abs().max().item(), (d_g[1].cpu() - d_o[1]).abs().max().item())
        torch.manual_seed(2); a = eo.update(x, -dt, d_o[0], 0.0)
        with S.host_noise():
            torch.manual_seed(2); b = eg.update_given_drift_and_diffusion(x=x.to(dev), dt=-dt.to(dev), drift=d_g[0], diffusion=0.0)
        print(kind, nw, "update", [(p.cpu() - q).abs().max().item() for p, q in zip(b, a)])
        tb_o = eo.traceback(a[0], x, t, -dt, score, bi, u); tb_g = eg.traceback_brownian_motion(x_next=b[0], x=x.to(dev), t=t.to(dev), dt=-dt.to(dev), score=score.to(dev), finetune_score=u.to(dev), batch_idx=bi.to(dev))
        print(kind, nw, "traceback", (tb_g.cpu() - tb_o).abs().max().item())
