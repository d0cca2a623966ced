"""Residue-steps/s of the three samplers (dpm_solver 50 steps, euler_maruyama_predictor 200 steps, heun_denoiser 100 steps; the
shipped YAML settings) at BASELINE config 1 (SH3 L = 56, B = 10) and config 2 (PDZ3 L = 84, B = 256) shapes on one GPU; device-
resident batch, CUDA events, third call timed (graphs captured)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from se3diff_b200 import shortcuts
from se3diff_b200.chemgraph import Batch, ChemGraph, complete_graph_edge_index
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = shortcuts.DiGConditionalScoreModel(precision="bf16").eval().to(dev)
so3 = shortcuts.DiGSO3SDE(**Bn.FULL_SDE).to(dev)
sdes = {"node_orientations": so3, "pos": shortcuts.CosineVPSDE(0.008)}
rows = []
for L, B in ((56, 10), (84, 256)):
    single, pair = Bn.synthetic_inputs(L)
    nan = float("nan")
    g = ChemGraph(pos=torch.full((L, 3), nan), node_orientations=torch.full((L, 3, 3), nan), edge_index=complete_graph_edge_index(L), single_embeds=single, pair_embeds=pair)
    batch = Batch.from_data_list([g] * B).to(dev)
    for name, fn, kw, evals in (("dpm_solver", shortcuts.dpm_solver, dict(num_steps=50), 2), ("euler_maruyama_predictor", shortcuts.euler_maruyama_predictor, dict(num_steps=200), 1),
                                ("heun_denoiser", shortcuts.heun_denoiser, dict(num_steps=100, noise=0.5), 2)):
        call = lambda: fn(batch=batch, sdes=sdes, score_model=model, max_t=0.99, min_t=0.001, device=dev, **kw)
        for _ in range(3): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = call(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        rows.append(dict(sampler=name, L=L, B=B, steps=kw["num_steps"], ms=round(ms, 1), residue_steps_per_s=round(B * L * kw["num_steps"] / ms * 1e3),
                         score_evals_per_s=round(kw["num_steps"] * evals / ms * 1e3), finite=bool(torch.isfinite(out["pos"]).all())))
        print(json.dumps(rows[-1]), flush=True)
