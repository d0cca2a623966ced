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
    shape = ops.ipa_shape(B, Lm, H, dk, 1, head_major=False)
    return proj, rot, trans, pair_bias, pair_value, hw, shape


def head_major(proj, shape):
    return proj[:, ops.ipa_head_major_perm(H, dk, proj.device)].contiguous(), ops.ipa_shape(shape.batch, shape.len, H, dk, 1, head_major=True)


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




def split(proj):
    """(bf16 scalar records with pre-scaled q, fp32 point records) of se3_ipa_attention_tc_fwd."""
    rows_s, rows_p, qpos = (i.to(proj.device) for i in ops.ipa_split_perms(H, dk))
    sc = proj[:, rows_s].clone()
    sc[:, qpos] *= (1 / math.sqrt(3 * dk)) * 1.4426950408889634
    return sc.to(torch.bfloat16).contiguous(), proj[:, rows_p].contiguous()
