import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se3diff_b200 import ops
torch.manual_seed(0)
for N, K in ((16, 16), (96, 16), (64, 96), (256, 64), (16, 128), (128, 256)):
    a = torch.randn(128, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    d = ops.debug_umma_gemm(a, b); torch.cuda.synchronize()
    ref = a.float() @ b.float().t()
    print(N, K, "max abs err", (d - ref).abs().max().item(), "ref scale", ref.abs().max().item())
