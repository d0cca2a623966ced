#!/usr/bin/env python
"""Benchmark of the batched SE(3) reverse-diffusion sampling path (BASELINE.json metric:
residue-steps/s = B * L * num_steps / time of the denoiser call).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c4|c5]

--config selects the BASELINE.json configuration: c2 (default, the one the metric is quoted on: PDZ3 L = 84, B = 256 per GPU),
c5 (synthetic 512-residue ensemble, 1024 samples over 8 GPUs = 128 per GPU) -- both this file's dpm_solver step -- and c4 (PDZ3
fine-tune step, B = 64 per GPU with the gradient all-reduce: scripts/bench_finetune.py under the same launch contract).

One bench "step" = ONE complete `dpm_solver` call (config/denoiser/dpm.yaml: 50 diffusion steps,
0.99 -> 0.001, i.e. 100 score-model evaluations + prior sampling) over one batch of synthetic input:
PSD95-PDZ3 length (L = 84), batch 256 per GPU, bioemu-v1.0 architecture with seeded random-init
weights, synthetic N(0,1) embeddings (BASELINE.json configs[1]).  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="PSD95-PDZ3 dpm_solver", L=84, B=256, num_steps=50, max_t=0.99, min_t=0.001)
CONFIGS = {"c2": dict(L=84, B=256, name="PSD95-PDZ3 dpm_solver"),                       # BASELINE.json configs[1]
           "c5": dict(L=512, B=128, name="synthetic 512-residue ensemble dpm_solver")}  # configs[4]: 1024 samples / 8 GPUs
HBM_NOMINAL_GBS = 8000.0   # north_star's "~8 TB/s" (DGX figure; HGX B200: 7.7 TB/s); reported next to the measured copy bandwidth
FULL_SDE = dict(eps_t=0.001, num_sigma=1000, num_omega=2000, omega_exponent=3, l_max=2000, sigma_min=0.02, sigma_max=2.33,
                tol=1e-7)
METRIC, UNIT = "SE(3) reverse-diffusion residue-steps/sec", "residue-steps/s"


def synthetic_inputs(L: int):
    g = torch.Generator().manual_seed(0)
    return torch.randn(L, 384, generator=g), torch.randn(L * L, 128, generator=g)


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference sampler on the host cores (bounded sample)
# ---------------------------------------------------------------------------------------------------
def cpu_reference_run(L: int, sample_B: int, sample_steps: int, state_dict, tables: dict | None, repeats: int = 1):
    """Times oracle.samplers.dpm_solver (the CPU restatement of bioemu.denoiser.dpm_solver driving the DiG
    score model exactly as the reference does, per-sample pair tensors recomputed every call)."""
    from oracle import samplers as osamp
    from oracle import so3 as oso3
    from oracle.score_model import ScoreModelOracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    single, pair = synthetic_inputs(L)
    if tables is None:  # reduced table resolution: table construction is not part of the timed path
        tab = oso3.SO3Tables(eps_t=0.001, num_sigma=200, num_omega=500, l_max=500, sigma_min=0.02, sigma_max=2.33)
    else:
        tab = oso3.SO3Tables(**FULL_SDE, tables=tables)
    r3 = osamp.CosineVP(0.008)
    model = ScoreModelOracle(state_dict, num_heads=32)
    model.set_context(single.repeat(sample_B, 1), [pair.view(L, L, 128)] * sample_B, [L] * sample_B)
    times = []
    with torch.no_grad():
        for rep in range(repeats):
            torch.manual_seed(rep)
            t0 = time.perf_counter()
            osamp.dpm_solver(model, [L] * sample_B, r3, tab, sample_steps, WORKLOAD["max_t"], WORKLOAD["min_t"])
            times.append(time.perf_counter() - t0)
    return times, cores


def claim_stdout():
    """stdout carries exactly ONE JSON line.  NCCL prints its version banner on file descriptor 1 from native code (at any
    NCCL_DEBUG level from VERSION up), so the descriptor itself is pointed at stderr for the whole run and the result line
    is written to a private duplicate of the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def reference_arm(args):
    """`--impl reference`: rank 0 only, CPU, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.score_model import init_state_dict

    L = WORKLOAD["L"]
    sB, sS = 8, 6
    sd = init_state_dict(seed=0)
    total = args.steps + args.warmup
    times, cores = cpu_reference_run(L, sB, sS, sd, None, repeats=total)
    timed = times[args.warmup:]
    per = sum(timed) / len(timed)
    value = sB * L * sS / per
    sample = f"dpm_solver on B={sB} of {WORKLOAD['B']} samples, {sS} of {WORKLOAD['num_steps']} diffusion steps, L={L}, fp32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD['name']} L={L} B={WORKLOAD['B']} steps={WORKLOAD['num_steps']} (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"], help="BASELINE.json configuration (default: the headline one)")
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (default: the configuration's)")
    ap.add_argument("--length", type=int, default=None)
    ap.add_argument("--diffusion-steps", type=int, default=WORKLOAD["num_steps"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true", help="skip the instrumented per-kernel pass (used for ncu launch lists)")
    ap.add_argument("--no-extras", action="store_true", help="skip the fp32-mode and replicated-host-batch passes after the timed region")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        return reference_arm(args)
    if args.config == "c4":                                   # the fine-tune step has its own driver under the same launch contract
        import runpy

        sys.argv = [os.path.join(ROOT, "scripts", "bench_finetune.py"), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup)]
        if args.batch is not None:
            sys.argv += ["--batch", str(args.batch)]
        if args.length is not None:
            sys.argv += ["--length", str(args.length)]
        return runpy.run_path(sys.argv[0], run_name="__main__")
    cfg = CONFIGS[args.config]
    args.batch = cfg["B"] if args.batch is None else args.batch
    args.length = cfg["L"] if args.length is None else args.length

    import torch.distributed as dist

    from se3diff_b200 import ops, shortcuts
    from se3diff_b200.chemgraph import Batch, ChemGraph, complete_graph_edge_index
    from se3diff_b200.distributed import gather_ensemble, init_from_env

    result_out = claim_stdout()
    rank, world, local_rank = init_from_env(args.gpus)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    L, B, S = args.length, args.batch, args.diffusion_steps

    # model, SDEs (table construction excluded from timing, as BASELINE.md section 2 prescribes)
    torch.manual_seed(0)
    model = shortcuts.DiGConditionalScoreModel(precision=args.precision).eval().to(dev)
    so3 = shortcuts.DiGSO3SDE(**FULL_SDE).to(dev)
    sdes = {"node_orientations": so3, "pos": shortcuts.CosineVPSDE(0.008)}
    single, pair = synthetic_inputs(L)
    nan = float("nan")
    graph = ChemGraph(pos=torch.full((L, 3), nan), node_orientations=torch.full((L, 3, 3), nan),
                      edge_index=complete_graph_edge_index(L), single_embeds=single, pair_embeds=pair)
    for k, v in graph.items():                                          # pinned host buffers (the batch below holds B references to
        if torch.is_tensor(v):                                          # this one graph and ships it to the device once per step)
            graph[k] = v.pin_memory()
    host_batch = Batch.from_data_list([graph] * B)                      # sample.py:223
    dev_batch = host_batch.to(dev)
    kw = dict(sdes=sdes, score_model=model, num_steps=S, max_t=WORKLOAD["max_t"], min_t=WORKLOAD["min_t"], device=dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def one_step(seed, batch, gather):
        torch.manual_seed(seed)                                             # seed = global sample offset (sample.py:288-306)
        out = shortcuts.dpm_solver(batch=batch, **kw)
        frames = torch.cat([out["pos"].view(B, L, 3), out["node_orientations"].view(B, L, 9)], dim=-1)
        return gather_ensemble(frames) if gather else frames

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing (value) -----------------------------------------------------------------
    for w in range(args.warmup):
        one_step(rank * B + w, dev_batch, True)
    clocks = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        clocks.start()
    ops.launch_count_reset()
    evs = []
    for k in range(args.steps):
        flush.zero_()                                                        # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step(1000 + rank * B + k, dev_batch, True)
        e1.record()
        evs.append((e0, e1))
    barrier()
    launches = ops.launch_count()
    clk = clocks.stop() if rank == 0 else None
    t_dev = sum(a.elapsed_time(b) for a, b in evs) / 1e3
    # ---- end-to-end timing through the public API with host buffers ----------------------------------------
    h2d = host_batch.h2d_nbytes()                                           # counted from the tensors the step copies
    d2h = B * L * 12 * 4
    for w in range(args.warmup):                                             # the host-buffer path has its own cold costs (a second
        one_step(1500 + rank * B + w, host_batch, False).to("cpu")           # 1 GB device block from cudaMalloc, pinned staging): untimed
    barrier()
    evs = []
    for k in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frames = one_step(2000 + rank * B + k, host_batch, False)             # batch.to(device) inside dpm_solver
        host_frames = frames.to("cpu")                                       # sample.py:235-236
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_e2e = sum(a.elapsed_time(b) for a, b in evs) / 1e3
    assert host_frames.shape == (B, L, 12) and bool(torch.isfinite(host_frames).all())
    times = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = times.tolist()

    # ---- after the timed region: the parity-precision figure and the PyG-shaped host batch -------------------------------
    extras = {}
    if not args.no_extras:
        def timed(fn, warm, reps):
            for w in range(warm):
                fn(w)
            barrier()
            ev = []
            for k in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(100 + k)
                e1.record()
                ev.append((e0, e1))
            barrier()
            tt = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / 1e3], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return tt.item()

        # (1) the same step in the OTHER precision mode (fp32 = the parity mode: SIMT fp32 attention, fp32 GEMMs)
        other = "fp32" if args.precision == "bf16" else "bf16"
        model.set_precision(other)
        t_o = timed(lambda k: one_step(4000 + rank * B + k, dev_batch, True), args.warmup, args.steps)
        model.set_precision(args.precision)
        extras[other] = {"value": world * B * L * S * args.steps / t_o, "unit": UNIT, "ms_per_step": t_o / args.steps * 1e3,
                         "note": "same workload, W warm-up + K timed steps, device-resident batch"}
        # (2) end to end with the batch an unchanged sample.py builds through PyG: every field materialised B times on the host
        # (pinned), i.e. the whole replicated pair embedding crosses PCIe every step
        if args.config == "c2" or B * L * L * 128 * 4 <= (2 << 30):
            rep = Batch.from_data_list([ChemGraph(**dict(graph.items())) for _ in range(B)])     # distinct graph objects: no replica shortcut
            for k, v in rep.items():
                if torch.is_tensor(v):
                    rep[k] = v.pin_memory()
            h2d_rep = rep.h2d_nbytes()
            # warm-up: the caching allocator settles after the third call (each of the first calls adds one 1 GB block by cudaMalloc:
            # the incoming copy, the copy the context cache still holds, the one being released -- scripts/time_e2e_pyg_parts.py:
            # 370 ms per call and no allocator traffic from then on, +40..110 ms on the calls before)
            t_p = timed(lambda k: one_step(5000 + rank * B + k, rep, False).to("cpu"), max(args.warmup, 4), args.steps)
            extras["e2e_pyg"] = {"value": world * B * L * S * args.steps / t_p, "unit": UNIT, "h2d_bytes_per_step": h2d_rep, "d2h_bytes_per_step": d2h,
                                 "ms_per_step": t_p / args.steps * 1e3,
                                 "note": "host batch with the pair embedding replicated B times (what PyG's Batch.from_data_list holds at sample.py:223), pinned"}
            del rep

    # ---- per-kernel roofline pass (same workload, instrumented, after the timed region) -------------------
    roof, roof_extra = None, []
    if rank == 0 and not args.no_roofline:
        try:
            from se3diff_b200.profiling import kernel_rooflines

            roof, roof_extra = kernel_rooflines(lambda: one_step(3000, dev_batch, False), L=L, B=B, sm_mhz=(clk or {}).get("sm_mhz"),
                                                with_elementwise=args.config == "c2")
        except Exception as e:  # noqa: BLE001
            roof = {"error": repr(e)}

    if rank == 0:
        value = world * B * L * S * args.steps / t_dev
        e2e = world * B * L * S * args.steps / t_e2e
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
                tables = dict(omega_grid=so3.igso3.omega_grid.cpu(), cdf_igso3=so3.igso3.cdf_igso3.cpu(),
                              cdf_uso3=so3.uso3.cdf_igso3.cpu(), score_scaling=so3.score_function.score_scaling.cpu())
                sB, sS = (8, 20) if L <= 128 else (1, 2)                   # ~10-20 s of CPU work on the box's host cores
                (t_cpu,), cores = cpu_reference_run(L, sB, sS, sd, tables)
                cpu = {"value": sB * L * sS / t_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                       "sample": f"oracle dpm_solver, B={sB} of {B} samples, {sS} of {S} diffusion steps, L={L}, fp32, "
                                 f"{cores} threads, {t_cpu:.1f} s"}
            except Exception as e:  # noqa: BLE001
                cpu = {"error": repr(e)}
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision if args.precision != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['name'] if L == cfg['L'] else 'synthetic-sequence dpm_solver'} L={L} B={B}/GPU {S} diffusion steps (2 score evals each), "
                                   f"bioemu-v1.0 architecture random-init, synthetic embeddings",
                       "step": "one dpm_solver call incl. prior sampling" + (" + NCCL ensemble all_gather" if world > 1 else ""),
                       "l2": "512 MiB buffer written between timed iterations", "parallelism": f"dp{world} (independent samples)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e / args.steps * 1e3},
            "gpu_launches": launches, "clocks": clk, "roofline": roof, "roofline_other_kernels": roof_extra, "cpu_baseline": cpu,
            **extras,
        }), file=result_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
