"""Which GELU does cuBLASLt's GELU_BIAS epilogue (torch._addmm_activation(use_gelu=True)) evaluate -- erf or the tanh approximation? (developer probe)"""
import torch, torch.nn.functional as F
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
N, K, M = 21504, 512, 1024
a = torch.randn(N, K, device=dev, generator=g)
w = torch.randn(M, K, device=dev, generator=g) * 0.05
b = torch.randn(M, device=dev, generator=g) * 0.1
torch.backends.cuda.matmul.allow_tf32 = False
for dt in (torch.float32, torch.bfloat16):
    A, W, Bb = a.to(dt), w.to(dt), b.to(dt)
    pre = F.linear(A.double(), W.double(), Bb.double())
    y = torch._addmm_activation(Bb, A, W.t(), use_gelu=True).double()
    for name, ref in (("erf", F.gelu(pre)), ("tanh", F.gelu(pre, approximate="tanh"))):
        print(dt, name, "max abs diff", (y - ref).abs().max().item(), "mean", (y - ref).abs().mean().item())
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
A, W, Bb = a.bfloat16(), w.bfloat16(), b.bfloat16()
print("linear + bias        us:", t(lambda: F.linear(A, W, Bb)))
print("addmm_activation     us:", t(lambda: torch._addmm_activation(Bb, A, W.t(), use_gelu=True)))
