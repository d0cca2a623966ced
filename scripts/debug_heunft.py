"""Per-step comparison of heun_denoiser_finetune (GPU) with the CPU oracle on the golden setup (developer diagnostics)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, yaml
import test_gpu_parity as tg
from oracle import samplers as osamp, so3 as oso3
from oracle.gen_golden import SMALL_SDE
from oracle.score_model import ScoreModelOracle
from se3diff_b200 import shortcuts

g, m, fm, sdes, batch, S = tg._traj_setup()
cfg = yaml.safe_load(str(g["cfg_json"])); L, B = int(g["L"]), int(g["B"]); lengths = [L] * B
T = torch.from_numpy
single = T(g["single"]).repeat(B, 1); pairs = [T(g["pair"])] * B
om = ScoreModelOracle(tg._sd(g, "sd::"), num_heads=cfg["num_heads"]).set_context(single, pairs, lengths)
ofm = ScoreModelOracle(tg._sd(g, "ft::"), num_heads=cfg["num_heads"]).set_context(single, pairs, lengths)
tab, r3 = oso3.SO3Tables(**SMALL_SDE), osamp.CosineVP(0.008)
n = int(g["heunft_steps"])
with torch.no_grad():
    torch.manual_seed(55); ref = osamp.heun_finetune(om, ofm, lengths, r3, tab, n, 0.99, 0.001, 0.5)
    with S.host_noise():
        torch.manual_seed(55)
        path = shortcuts.heun_denoiser_finetune(batch=batch, sdes=sdes, score_model=m, finetune_model=fm, noise=0.5, num_steps=n, max_t=0.99, min_t=0.001, device="cuda")
for i in range(n + 1):
    print(i, "pos diff", (path.batches[i]["pos"].cpu() - ref.pos[i]).abs().max().item(), "rot diff", (path.batches[i]["node_orientations"].cpu() - ref.rot[i]).abs().max().item())
for f in ("pos", "node_orientations"):
    print(f, "us diff per step", (path.us_batch[f].cpu() - ref.us[f]).abs().amax(dim=(1, 2, 3)).tolist())
    print(f, "dWs diff per step", (path.dWs_batch[f].cpu() - ref.dWs[f]).abs().amax(dim=(1, 2, 3)).tolist())

# ---- call-by-call comparison of the network evaluations
log_g, log_o = [], []
def wrap_g(model, tag):
    orig = model.forward
    def f(b, t):
        out = orig(b, t)
        log_g.append((tag, float(t[0]), b["pos"].detach().cpu().clone(), b["node_orientations"].detach().cpu().clone(), out["pos"].detach().cpu().clone()))
        return out
    model.forward = f
def wrap_o(model, tag):
    class W:
        def __call__(self, pos, rot, t):
            p, r = model(pos, rot, t)
            log_o.append((tag, float(t[0]), pos.clone(), rot.clone(), p.clone()))
            return p, r
    return W()
wrap_g(m, "score"); wrap_g(fm, "ctrl")
with torch.no_grad():
    torch.manual_seed(55); osamp.heun_finetune(wrap_o(om, "score"), wrap_o(ofm, "ctrl"), lengths, r3, tab, 2, 0.99, 0.001, 0.5)
    with S.host_noise():
        torch.manual_seed(55)
        shortcuts.heun_denoiser_finetune(batch=batch, sdes=sdes, score_model=m, finetune_model=fm, noise=0.5, num_steps=2, max_t=0.99, min_t=0.001, device="cuda")
print(len(log_g), len(log_o))
for a, b in zip(log_g, log_o):
    print(a[0], b[0], a[1], b[1], "in pos", (a[2] - b[2]).abs().max().item(), "in rot", (a[3] - b[3]).abs().max().item(), "out", (a[4] - b[4]).abs().max().item())
