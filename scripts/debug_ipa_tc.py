"""Tensor-core IPA vs fp64 reference and vs the SIMT kernel; timing at the bench shape (GPU box)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from se3diff_b200 import ops, _lib as L

dev = "cuda"
H, dk, D = 32, 16, 512


def make(B, Lm, seed=0, pos_scale=1.5):
    g = torch.Generator(device=dev).manual_seed(seed)
    proj = torch.randn(B * Lm, 3 * D + 48 * H, generator=g, device=dev)
    rot = ops.so3_exp(torch.randn(B * Lm, 3, generator=g, device=dev)).reshape(B * Lm, 9)
    trans = torch.randn(B * Lm, 3, generator=g, device=dev) * pos_scale
    pair_bias = torch.randn(1, H, Lm, Lm, generator=g, device=dev)
    pair_value = torch.randn(1, Lm, Lm, H * dk, generator=g, device=dev)
    hw = -0.5 * (1 / math.sqrt(54)) * F.softplus(torch.rand(H, generator=g, device=dev))
    shape = L.IpaShape(B, Lm, H, dk, 4, 8, 3 * D + 48 * H, 0, D, 2 * D, 3 * D, 3 * D + 12 * H, 3 * D + 24 * H, 1)
    return proj, rot, trans, pair_bias, pair_value, hw, shape


def ref(proj, rot, trans, pair_bias, pair_value, hw, B, Lm, dt=torch.float64):
    sw = 1 / math.sqrt(3 * dk)
    P = proj.to(dt).view(B, Lm, -1)
    blk = lambda o, w: P[..., o:o + w]
    q = blk(0, H * dk).reshape(B, Lm, H, dk); k = blk(D, H * dk).reshape(B, Lm, H, dk); v = blk(2 * D, H * dk).reshape(B, Lm, H, dk)
    qp = blk(3 * D, H * 12).reshape(B, Lm, H, 4, 3); kp = blk(3 * D + 12 * H, H * 12).reshape(B, Lm, H, 4, 3)
    vp = blk(3 * D + 24 * H, H * 24).reshape(B, Lm, H, 8, 3)
    R = rot.to(dt).view(B, Lm, 3, 3); T = trans.to(dt).view(B, Lm, 3)
    glob = lambda x: torch.matmul(R[:, :, None, None], x.unsqueeze(-1)).squeeze(-1) + T[:, :, None, None]
    qp, kp, vp = glob(qp), glob(kp), glob(vp)
    s = torch.einsum("bihc,bjhc->bhij", q * sw, k)
    d = torch.norm(qp.unsqueeze(2) - kp.unsqueeze(1), dim=-1).sum(-1).permute(0, 3, 1, 2)
    a = torch.softmax(s + hw.to(dt)[None, :, None, None] * d + pair_bias.to(dt), -1)
    o_s = torch.einsum("bhij,bjhc->bihc", a, v).reshape(B, Lm, -1)
    o_pg = torch.einsum("bhij,bjhcp->bihcp", a, vp)
    o_pl = torch.matmul(R.transpose(-1, -2)[:, :, None, None], (o_pg - T[:, :, None, None]).unsqueeze(-1)).squeeze(-1)
    o_n = torch.norm(o_pl, dim=-1).reshape(B, Lm, -1)
    pvv = pair_value.to(dt).view(1, Lm, Lm, H, dk).expand(B, -1, -1, -1, -1)
    o_pair = torch.einsum("bhij,bijhc->bihc", a, pvv).reshape(B, Lm, -1)
    return torch.cat([o_s, o_pl.reshape(B, Lm, -1), o_pair, o_n], -1).reshape(B * Lm, -1)


names = [("scalar", 0, 512), ("point", 512, 1280), ("pair", 1280, 1792), ("norm", 1792, 2048)]
for B, Lm, scale in ((3, 84, 1.5), (2, 56, 1.5), (130, 20, 1.5), (2, 200, 1.5), (2, 84, 100.0)):
    proj, rot, trans, pb, pv, hw, shape = make(B, Lm, seed=Lm, pos_scale=scale)
    r64 = ref(proj, rot, trans, pb, pv, hw, B, Lm)
    ws = ops.ipa_tc_workspace(shape, dev)
    pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
    for odt in (torch.float32, torch.bfloat16):
        o = ops.ipa_attention_tc_fwd(proj, rot, trans, pbt, pvp, None, hw, 1 / math.sqrt(48), shape, ws, out_dtype=odt)
        torch.cuda.synchronize()
        e = (o.double() - r64).abs()
        print(f"B={B} L={Lm} scale={scale} out={odt}:", {n: f"{e[:, a:b].max().item():.2e}/{r64[:, a:b].abs().max().item():.1f}" for n, a, b in names},
              "nan" if torch.isnan(o).any() else "")

# timing at the bench shape
B, Lm = 256, 84
proj, rot, trans, pb, pv, hw, shape = make(B, Lm)
ws = ops.ipa_tc_workspace(shape, dev); pvp = ops.ipa_tc_pack_pair_value(pv, H); pbt = ops.ipa_tc_pack_pair_bias(pb.permute(0, 2, 3, 1))
out = torch.empty(B * Lm, 2048, dtype=torch.bfloat16, device=dev)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("tc   ms:", t(lambda: ops.ipa_attention_tc_fwd(proj, rot, trans, pbt, pvp, None, hw, 1 / math.sqrt(48), shape, ws, out=out)))
print("simt ms:", t(lambda: ops.ipa_attention_fwd(proj, rot, trans, pb, pv, None, hw, 1 / math.sqrt(48), shape, 1)))
