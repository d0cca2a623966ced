"""Times se3_ipa_attention_bwd at the fine-tune loss-side shape (K*B = 1280, L = 84, H = 4, dk = 16) and a tiled-edition shape
(developer microbenchmark; SE3DIFF_B200_LIB selects another build for A/B)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se3diff_b200 import ops
dev = "cuda"
def run(B, L, H, dk, iters=10):
    g = torch.Generator(device=dev).manual_seed(0)
    sh = ops.ipa_shape(B, L, H, dk, 1, head_major=False)
    proj = torch.randn(B * L, sh.proj_stride, device=dev, generator=g)
    rot = ops.so3_exp(torch.randn(B * L, 3, device=dev, generator=g)).reshape(B * L, 9)
    trans = torch.randn(B * L, 3, device=dev, generator=g) * 2
    pb = torch.randn(1, H, L, L, device=dev, generator=g)
    pv = torch.randn(1, L, L, H * dk, device=dev, generator=g)
    hw = -torch.rand(H, device=dev, generator=g) * 0.1
    out = ops.ipa_attention_fwd(proj, rot, trans, pb, pv, None, hw, 0.25, sh)
    d_out = torch.randn_like(out)
    f = lambda: ops.ipa_attention_bwd(proj, rot, trans, pb, pv, None, hw, 0.25, out, d_out, sh)
    for _ in range(3): r = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): r = f()
    e1.record(); torch.cuda.synchronize()
    print(f"{os.environ.get('SE3DIFF_B200_LIB', 'default')}: B={B} L={L} H={H} dk={dk}: {e0.elapsed_time(e1) / iters:.3f} ms per call, checksum {float(r[0].abs().mean()):.6f}")
run(1280, 84, 4, 16)
run(256, 84, 32, 16)
if os.environ.get("SE3DIFF_B200_LIB") is None:
    run(64, 256, 4, 16, iters=5)
    run(16, 512, 4, 16, iters=5)
