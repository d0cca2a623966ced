"""Projection GEMM layouts: one [N,1536] GEMM vs head-batched [H][N][48] (developer microbenchmark)."""
import torch
dev = "cuda"
N, D, H = 21504, 512, 32
x = torch.randn(N, D, device=dev, dtype=torch.bfloat16)
w = torch.randn(H * 48, D, device=dev, dtype=torch.bfloat16)
w3 = w.view(H, 48, D)
wt3 = w3.transpose(1, 2).contiguous()        # [H, D, 48]
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
print("mm bf16 out      us:", t(lambda: torch.mm(x, w.t())))
print("mm fp32 out      us:", t(lambda: torch.mm(x, w.t(), out_dtype=torch.float32)))
xe = x.unsqueeze(0).expand(H, N, D)
print("bmm bf16 out (x expand, w [H,48,D]^T view) us:", t(lambda: torch.bmm(xe, w3.transpose(1, 2))))
print("bmm bf16 out (x expand, wt3 contiguous)   us:", t(lambda: torch.bmm(xe, wt3)))
try:
    print("bmm fp32 out us:", t(lambda: torch.bmm(xe, w3.transpose(1, 2), out_dtype=torch.float32)))
except Exception as e:
    print("bmm out_dtype unsupported:", e)
a = torch.bmm(xe, w3.transpose(1, 2)); b = torch.mm(x, w.t()).view(N, H, 48).permute(1, 0, 2)
print("max diff", (a.float() - b.float()).abs().max().item(), a.shape, a.is_contiguous())
print("mem after", torch.cuda.max_memory_allocated() / 1e6, "MB")
# transposed product: [H*48, D] @ [D, N] -> [H*48, N]
print("mm transposed (w @ x^T) fp32 us:", t(lambda: torch.mm(w, x.t(), out_dtype=torch.float32)))
