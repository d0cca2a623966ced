"""Truth function of the tensor-core IPA operator tests: SAAttention.forward between the input projections and fc_out
(structure_module.py:131-216) as fp64 einsums on the operator's own inputs (fused projection rows, frames, pair bias / pair
values already projected).  It is pinned to `ScoreModelOracle._ipa` -- the bit-exact restatement of the reference -- by
tests/test_host_logic.py::test_tc_operator_truth_function_equals_the_oracle_on_cpu; the GPU tests and the developer scripts
(scripts/debug_ipa_tc_common.py re-exports this module) compare the kernels against it."""
import math

import torch
import torch.nn.functional as F

H, dk, D = 32, 16, 512


def make(B, Lm, seed=0, pos_scale=1.5, dev="cuda"):
    from se3diff_b200 import ops

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
    from se3diff_b200 import ops

    return proj[:, ops.ipa_head_major_perm(H, dk, proj.device)].contiguous(), ops.ipa_shape(shape.batch, shape.len, H, dk, 1, head_major=True)


def ref(proj, rot, trans, pair_bias, pair_value, hw, B, Lm, dt=torch.float64, heads=H, d_k=dk, key_bias=None):
    """proj [B*L, 3*H*dk + 48*H] = q | k | v | q_pt | k_pt | v_pt (block-major, the reference's parameter order), rot [B*L, 9],
    trans [B*L, 3], pair_bias [1|B, H, L, L] (already times pair_weight), pair_value [1|B, L, L, H*dk], hw [H] =
    -0.5 * point_weight * softplus(gamma); key_bias [B, L] additive mask or None
    -> [B*L, H*(2*dk + 32)] = scalar | point_local | pair | point_norm (structure_module.py:216)."""
    Hh, dd = heads, d_k
    Dm = Hh * dd
    sw = 1 / math.sqrt(3 * dd)
    P = proj.to(dt).view(B, Lm, -1)
    blk = lambda o, w: P[..., o:o + w]
    q = blk(0, Dm).reshape(B, Lm, Hh, dd); k = blk(Dm, Dm).reshape(B, Lm, Hh, dd); v = blk(2 * Dm, Dm).reshape(B, Lm, Hh, dd)
    qp = blk(3 * Dm, Hh * 12).reshape(B, Lm, Hh, 4, 3); kp = blk(3 * Dm + 12 * Hh, Hh * 12).reshape(B, Lm, Hh, 4, 3)
    vp = blk(3 * Dm + 24 * Hh, Hh * 24).reshape(B, Lm, Hh, 8, 3)
    R = rot.to(dt).view(B, Lm, 3, 3); T = trans.to(dt).view(B, Lm, 3)
    glob = lambda x: torch.matmul(R[:, :, None, None], x.unsqueeze(-1)).squeeze(-1) + T[:, :, None, None]
    qp, kp, vp = glob(qp), glob(kp), glob(vp)
    s = torch.einsum("bihc,bjhc->bhij", q * sw, k)
    d = torch.norm(qp.unsqueeze(2) - kp.unsqueeze(1), dim=-1).sum(-1).permute(0, 3, 1, 2)
    logits = s + hw.to(dt)[None, :, None, None] * d + pair_bias.to(dt)
    if key_bias is not None:
        logits = logits + key_bias.to(dt)[:, None, None, :]
    a = torch.softmax(logits, -1)
    o_s = torch.einsum("bhij,bjhc->bihc", a, v).reshape(B, Lm, -1)
    o_pg = torch.einsum("bhij,bjhcp->bihcp", a, vp)
    o_pl = torch.matmul(R.transpose(-1, -2)[:, :, None, None], (o_pg - T[:, :, None, None]).unsqueeze(-1)).squeeze(-1)
    o_n = torch.norm(o_pl, dim=-1).reshape(B, Lm, -1)
    pvv = pair_value.to(dt).view(pair_value.shape[0], Lm, Lm, Hh, dd).expand(B, -1, -1, -1, -1)
    o_pair = torch.einsum("bhij,bijhc->bihc", a, pvv).reshape(B, Lm, -1)
    return torch.cat([o_s, o_pl.reshape(B, Lm, -1), o_pair, o_n], -1).reshape(B * Lm, -1)


def split(proj):
    """(bf16 scalar records with pre-scaled q, fp32 point records) of se3_ipa_attention_tc_fwd."""
    from se3diff_b200 import ops

    rows_s, rows_p, qpos = (i.to(proj.device) for i in ops.ipa_split_perms(H, dk))
    sc = proj[:, rows_s].clone()
    sc[:, qpos] *= (1 / math.sqrt(3 * dk)) * 1.4426950408889634
    return sc.to(torch.bfloat16).contiguous(), proj[:, rows_p].contiguous()
