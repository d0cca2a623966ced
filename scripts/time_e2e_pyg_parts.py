"""Where the PyG-shaped end-to-end step (every field materialised B times on the pinned host, bench.py `e2e_pyg`) spends its time
beyond the device-resident step (GPU box; developer diagnostics)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from se3diff_b200 import shortcuts
from se3diff_b200.chemgraph import Batch, ChemGraph, complete_graph_edge_index
dev = torch.device("cuda", 0)
L, B, S = 84, 256, 50
torch.manual_seed(0)
model = shortcuts.DiGConditionalScoreModel(precision="bf16").eval().to(dev)
so3 = shortcuts.DiGSO3SDE(**Bn.FULL_SDE).to(dev)
sdes = {"node_orientations": so3, "pos": shortcuts.CosineVPSDE(0.008)}
single, pair = Bn.synthetic_inputs(L)
nan = float("nan")
graph = ChemGraph(pos=torch.full((L, 3), nan), node_orientations=torch.full((L, 3, 3), nan), edge_index=complete_graph_edge_index(L), single_embeds=single, pair_embeds=pair)
rep = Batch.from_data_list([ChemGraph(**dict(graph.items())) for _ in range(B)])
for k, v in rep.items():
    if torch.is_tensor(v): rep[k] = v.pin_memory()
kw = dict(sdes=sdes, score_model=model, num_steps=S, max_t=0.99, min_t=0.001, device=dev)
def sync_time(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
dev_batch = rep.to(dev)
for _ in range(4): shortcuts.dpm_solver(batch=dev_batch, **kw)
print("device-resident step (same device batch again): %.1f ms" % sync_time(lambda: shortcuts.dpm_solver(batch=dev_batch, **kw))[0])
big = rep["pair_embeds"]
for _ in range(3):
    t, _ = sync_time(lambda: big.to(dev, non_blocking=True))
    print("pair_embeds alone, pinned -> device           : %.1f ms (%.0f MB, %.1f GB/s)" % (t, big.numel() * 4 / 1e6, big.numel() * 4 / t / 1e6))
for _ in range(3):
    t, nb = sync_time(lambda: rep.to(dev))
    print("rep.to(device)                                : %.1f ms (%.0f MB)" % (t, sum(v.numel() * v.element_size() for _, v in rep.items() if torch.is_tensor(v)) / 1e6))
    t, out = sync_time(lambda: shortcuts.dpm_solver(batch=nb, **kw))
    print("step on the fresh device copy                 : %.1f ms" % t)
for _ in range(3):
    print("end-to-end (host batch in)                    : %.1f ms" % sync_time(lambda: shortcuts.dpm_solver(batch=rep, **kw))[0])
# allocator behaviour per end-to-end call: cudaMalloc / cudaFree counts and reserved bytes (outlier steps are allocator traffic?)
for i in range(8):
    s0 = torch.cuda.memory_stats()
    t = sync_time(lambda: shortcuts.dpm_solver(batch=rep, **kw))[0]
    s1 = torch.cuda.memory_stats()
    print("call %d: %.1f ms  cudaMalloc +%d  cudaFree +%d  reserved %.2f GB  allocated %.2f GB  retries +%d" % (
        i, t, s1["num_device_alloc"] - s0["num_device_alloc"], s1["num_device_free"] - s0["num_device_free"],
        s1["reserved_bytes.all.current"] / 2**30, s1["allocated_bytes.all.current"] / 2**30, s1["num_alloc_retries"] - s0["num_alloc_retries"]))
