"""Eager vs CUDA-graph timing of one score-model forward at the bench shape (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from se3diff_b200 import shortcuts, ops
from se3diff_b200.chemgraph import Batch, ChemGraph, complete_graph_edge_index
dev = torch.device("cuda")
L, B = int(os.environ.get("L", 84)), int(os.environ.get("B", 256))
torch.manual_seed(0)
model = shortcuts.DiGConditionalScoreModel(precision="bf16").eval().to(dev)
g = torch.Generator().manual_seed(0)
single, pair = torch.randn(L, 384, generator=g), torch.randn(L * L, 128, generator=g)
graph = ChemGraph(pos=torch.randn(L, 3), node_orientations=torch.eye(3).repeat(L, 1, 1), edge_index=complete_graph_edge_index(L),
                  single_embeds=single, pair_embeds=pair)
batch = Batch.from_data_list([graph] * B).to(dev)
batch = batch.replace(pos=torch.randn(B * L, 3, device=dev), node_orientations=ops.so3_exp(torch.randn(B * L, 3, device=dev)))
t = torch.full((B,), 0.5, device=dev)
torch.set_grad_enabled(False)   # the kernel path is inference only (a forward that needs gradients takes the autograd path)
for _ in range(3): out = model(batch, t)
torch.cuda.synchronize()
def timeit(fn, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); w0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - w0) * 1e3 / n
print("eager  ms/forward (gpu, wall):", timeit(lambda: model(batch, t)))
# cpu-side cost only
w0 = time.perf_counter()
for _ in range(20): model(batch, t)
cpu = (time.perf_counter() - w0) * 1e3 / 20
torch.cuda.synchronize()
print("eager  cpu enqueue ms/forward:", cpu)
g_ = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2): model(batch, t)
    torch.cuda.synchronize()
    with torch.cuda.graph(g_, stream=s):
        out = model(batch, t)
torch.cuda.synchronize()
print("graph  ms/forward (gpu, wall):", timeit(lambda: g_.replay()))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): model(batch, t)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:18]
tot = sum(r.device_time_total for r in prof.key_averages())
print("total kernel us / forward:", tot / 3)
for r in rows: print(f"{r.device_time_total/3:9.1f} us  n={r.count//3:4d}  {r.key[:100]}")
