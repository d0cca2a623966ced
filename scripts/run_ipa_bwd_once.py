"""A few launches of se3_ipa_attention_bwd (resident edition at the fine-tune shape, tiled edition at L = 256) -- ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se3diff_b200 import ops
dev = "cuda"
for B, L, H, dk in ((1280, 84, 4, 16), (64, 256, 4, 16)):
    g = torch.Generator(device=dev).manual_seed(0)
    sh = ops.ipa_shape(B, L, H, dk, 1, head_major=False)
    proj = torch.randn(B * L, sh.proj_stride, device=dev, generator=g)
    rot = ops.so3_exp(torch.randn(B * L, 3, device=dev, generator=g)).reshape(B * L, 9)
    trans = torch.randn(B * L, 3, device=dev, generator=g) * 2
    pb = torch.randn(1, H, L, L, device=dev, generator=g)
    pv = torch.randn(1, L, L, H * dk, device=dev, generator=g)
    hw = -torch.rand(H, device=dev, generator=g) * 0.1
    out = ops.ipa_attention_fwd(proj, rot, trans, pb, pv, None, hw, 0.25, sh)
    d_out = torch.randn(out.shape, device=dev, generator=g)
    for _ in range(2):
        r = ops.ipa_attention_bwd(proj, rot, trans, pb, pv, None, hw, 0.25, out, d_out, sh)
    torch.cuda.synchronize()
    print("ok", B, L, float(r[0].abs().mean()))
