"""Where the end-to-end step (host batch in, frames out) spends its time beyond the device-resident step (GPU box)."""
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
host_batch = Batch.from_data_list([graph] * B)
for k, v in host_batch.items():
    if torch.is_tensor(v): host_batch[k] = v.pin_memory()
kw = dict(sdes=sdes, score_model=model, num_steps=S, max_t=0.99, min_t=0.001, device=dev)
def sync_time(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r
dev_batch = host_batch.to(dev)
for _ in range(4): shortcuts.dpm_solver(batch=dev_batch, **kw)
print("device-resident step      : %.1f ms" % sync_time(lambda: shortcuts.dpm_solver(batch=dev_batch, **kw))[0])
for _ in range(3):
    t, nb = sync_time(lambda: host_batch.to(dev))
    print("host_batch.to(device)     : %.1f ms (%.0f MB)" % (t, sum(v.numel() * v.element_size() for _, v in host_batch.items() if torch.is_tensor(v)) / 1e6))
    t, out = sync_time(lambda: shortcuts.dpm_solver(batch=nb, **kw))
    print("step on the fresh copy    : %.1f ms" % t)
    t, _ = sync_time(lambda: torch.cat([out["pos"].view(B, L, 3), out["node_orientations"].view(B, L, 9)], dim=-1).to("cpu"))
    print("frames to host            : %.1f ms" % t)
for _ in range(3):
    print("end-to-end (host batch in): %.1f ms" % sync_time(lambda: shortcuts.dpm_solver(batch=host_batch, **kw))[0])
