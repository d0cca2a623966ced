"""The four GEMMs of one structure-module layer at the bench shape (21504 rows, bf16): torch default (cuBLASLt heuristic) against
torch.cuda.tunable's pick (developer microbenchmark)."""
import os, sys, torch, torch.nn.functional as F
dev, N = "cuda", 21504
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, device=dev, generator=g).to(torch.bfloat16)
cases = {   # name: (fn builder) -- the calls models._forward_fused makes
    "proj     [N,512]x[3072,512]^T bf16": (r(N, 512), r(3072, 512), None, "mm"),
    "fc_out   [N,2048]x[512,2048]^T bf16": (r(N, 2048), r(512, 2048), None, "mm"),
    "ffn up   [N,512]x[1024,512]^T + bias": (r(N, 512), r(1024, 512), r(1024), "linear"),
    "ffn down [N,1024]x[512,1024]^T bf16": (r(N, 1024), r(512, 1024), None, "mm"),
}
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
def fns():
    out = {}
    for k, (a, w, b, kind) in cases.items():
        out[k] = (lambda a=a, w=w: torch.mm(a, w.t())) if kind == "mm" else (lambda a=a, w=w, b=b: F.linear(a, w, b))
    return out
base = {k: t(f) for k, f in fns().items()}
import torch.cuda.tunable as tun
tun.enable(True); tun.tuning_enable(True); tun.set_max_tuning_duration(200); tun.set_max_tuning_iterations(20)
tun.set_filename(os.environ.get("TUNE_FILE", "/tmp/tunableop.csv"))
tuned = {k: t(f) for k, f in fns().items()}
for k in cases:
    fl = 2.0 * cases[k][0].shape[0] * cases[k][0].shape[1] * cases[k][1].shape[0]
    print(f"{k:40s} default {base[k]:6.1f} us ({fl / base[k] / 1e6:6.0f} TF/s)   tuned {tuned[k]:6.1f} us ({fl / tuned[k] / 1e6:6.0f} TF/s)")
print(tun.get_results())
