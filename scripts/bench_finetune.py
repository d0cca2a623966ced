"""BASELINE config 4: one fine-tune step (finetune.py:599-626) on PDZ3 L = 84, B = 64 samples per GPU -- rollout through the
CUDA path (200 Euler-Maruyama steps with control: 200 evaluations of the 31 M-parameter score model in bf16 mode and of the
0.19 M-parameter control model), observable, chunked loss + backward through the control model, gradient all-reduce
(N > 1), AdamW step.  Synthetic embeddings, seeded random-init weights (control model scaled towards zero like
`initialize_weights_to_near_zero`, finetune.py:102-122), synthetic reference C-alpha trace.

    python scripts/bench_finetune.py [--steps K --warmup W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_finetune.py --gpus N

Prints ONE JSON line (rank 0): residue-steps/s = N * B * L * 200 / step time, and the split rollout / loss+backward / exchange+optimizer."""
import argparse, functools, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench as Bn
from se3diff_b200 import finetune_step as FS
from se3diff_b200 import shortcuts
from se3diff_b200.chemgraph import ChemGraph, complete_graph_edge_index
from se3diff_b200.distributed import allreduce_gradients, init_from_env


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--length", type=int, default=84)
    ap.add_argument("--em-steps", type=int, default=200)
    ap.add_argument("--micro", type=int, default=20, help="stored states per backward chunk (finetune.py micro_batch_size)")
    ap.add_argument("--profile", default=None, help="write a torch-profiler kernel table of one extra step to this file (rank 0)")
    a = ap.parse_args()
    result_out = Bn.claim_stdout()                          # stdout = the one JSON line (NCCL's banner goes to stderr)
    rank, world, local = init_from_env(a.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    L, B, T = a.length, a.batch, a.em_steps
    torch.manual_seed(0)
    score = shortcuts.DiGConditionalScoreModel(precision="bf16").eval().to(dev)
    ctrl = shortcuts.DiGConditionalScoreModel(dim_hidden=256, dim_model=64, dim_pair=32, dim_single_rep=16, num_heads=4, num_layers=2).to(dev)
    with torch.no_grad():
        for p in ctrl.parameters():
            p.mul_(1e-2)
    for p in score.parameters():
        p.requires_grad_(False)
    ctrl.eval()                                               # dropout off for the timing (finetune.py trains with dropout 0.1)
    sdes = {"node_orientations": shortcuts.DiGSO3SDE(**Bn.FULL_SDE).to(dev), "pos": shortcuts.CosineVPSDE(0.008)}
    single, pair = Bn.synthetic_inputs(L)
    nan = float("nan")
    graph = ChemGraph(pos=torch.full((L, 3), nan), node_orientations=torch.full((L, 3, 3), nan), edge_index=complete_graph_edge_index(L),
                      single_embeds=single, pair_embeds=pair)
    g = torch.Generator().manual_seed(1)
    ref = torch.cumsum(torch.nn.functional.normalize(torch.randn(L, 3, generator=g), dim=-1) * 0.38, dim=0).to(dev)   # 0.38 nm C-alpha steps
    denoiser = functools.partial(shortcuts.euler_maruyama_predictor_finetune, num_steps=T, max_t=0.99, min_t=0.001)
    bundle = FS.FinetuneBundle(sdes, score, ctrl, denoiser, FS.FoldingStability(ref_coords=ref))
    opt = torch.optim.AdamW(ctrl.parameters(), lr=1e-5)
    h_stars = torch.tensor([0.5])
    n_train = sum(p.numel() for p in ctrl.parameters() if p.requires_grad)

    def step(seed, ev=None):
        mark = (lambda k: ev[k].record()) if ev is not None else (lambda k: None)
        opt.zero_grad()
        mark(0)
        path = FS.generate_finetune_batch(chemgraph=graph, finetune_bundle=bundle, batch_size=B, device=dev, seed=seed)
        mark(1)
        loss = FS.compute_finetune_loss(sequence="A" * L, h_stars=h_stars, finetune_bundle=bundle, denoised_sde_path=path, batch_size=B,
                                        device=dev, for_grad=True, micro_batch_size=a.micro)
        mark(2)
        allreduce_gradients(ctrl.parameters())
        opt.step()
        mark(3)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    from se3diff_b200 import denoiser as _den
    from se3diff_b200 import ops as _ops

    h2d = sum(v.numel() * v.element_size() for _, v in graph.items() if torch.is_tensor(v))
    for w in range(a.warmup):
        step(rank * 1000 + w)
    barrier()
    _ops.launch_count_reset()
    parts = torch.zeros(3, dtype=torch.float64)
    for k in range(a.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        loss = step(10_000 + rank * 1000 + k, ev)
        torch.cuda.synchronize(dev)
        parts += torch.tensor([ev[i].elapsed_time(ev[i + 1]) for i in range(3)], dtype=torch.float64)
    barrier()
    Bn_launches = _ops.launch_count()
    t = parts.to(dev) / 1e3
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    per_rank = [[round(float(v) / a.steps * 1e3, 1) for v in r.cpu()] for r in per_rank]
    t = t.cpu()
    gn = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in ctrl.parameters() if p.grad is not None)).item()
    if rank == 0:
        total = float(t.sum())
        print(json.dumps({
            "metric": "fine-tune residue-steps/sec (rollout + loss + backward + exchange + optimizer)", "unit": "residue-steps/s",
            "value": world * B * L * T * a.steps / total, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": total / a.steps * 1e3, "ms_rollout": float(t[0]) / a.steps * 1e3, "ms_loss_backward": float(t[1]) / a.steps * 1e3,
            "ms_exchange_optimizer": float(t[2]) / a.steps * 1e3, "scaling": "weak", "dtype": "bf16 score model, fp32 control model and SDE algebra",
            "higher_is_better": True, "vs_baseline": None, "gpu_launches": int(Bn_launches), "loop_graphs": dict(_den.GRAPH_STATS), "ms_per_rank_rollout_loss_exchange": per_rank,
            "e2e": {"value": world * B * L * T * a.steps / total, "unit": "residue-steps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4,
                    "note": "the step itself starts from the host-resident graph (generate_finetune_batch moves it) and ends with the loss scalar on the host"},
            "data": "synthetic", "trainable_parameters": n_train, "loss": float(loss), "grad_norm": gn, "finite": bool(torch.isfinite(loss)),
            "config": {"workload": f"PDZ3 fine-tune step L={L} B={B}/GPU, {T} EM steps with control, micro_batch_size={a.micro}, "
                                   f"allreduce of {n_train} gradient floats" + (" over NCCL" if world > 1 else " (single rank: none)")}}), file=result_out, flush=True)
    if a.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step(77_000)
            torch.cuda.synchronize(dev)
        with open(a.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
