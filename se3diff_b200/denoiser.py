"""Samplers at the drop-in boundary -- same keyword-only signatures and return types as
bioemu/src/bioemu/denoiser.py:206-215 (euler_maruyama_predictor), :267-277 (..._finetune),
:351-361 (heun_denoiser), :464-475 (heun_denoiser_finetune), :634-643 (dpm_solver).

What differs from the reference is only *how* a step is executed:
  * the schedule (alpha, sigma, lambda, h, t_lambda, beta, g, score scaling) is evaluated once per call
    on the host (schedule.py) -- no `.item()` syncs inside the loop (denoiser.py:669, 692);
  * the score conversion of `_get_score` (denoiser.py:169-203) and the whole per-field update
    (`EulerMaruyamaPredictor`, denoiser.py:30-166) run as ONE fused CUDA kernel per half-step
    (se3_frame_update_*), reading the raw score-model output;
  * noise that the reference draws but multiplies by zero (denoiser.py:80 under diffusion=0.0) is not
    generated, unless `sdes.host_noise()` parity mode is active, where every draw of the reference is
    reproduced in order on the CPU generator.
`score_model` may be any callable `(batch, t) -> mapping with "pos" and "node_orientations"` as in the
reference (denoiser.py:219-221).
"""
from __future__ import annotations

import os
from collections import defaultdict
from typing import NamedTuple

import torch
from torch import nn

from . import ops, schedule
from . import sdes as S
from .chemgraph import batch_lengths


class DenoisedSDEPath(NamedTuple):
    """denoiser.py:23-27."""

    batches: list
    timesteps: torch.Tensor
    us_batch: dict
    dWs_batch: dict


def _prepare(batch, sdes, score_model, device, extra_models=()):
    if device is None:
        device = batch["pos"].device
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"se3diff_b200 samplers run on CUDA devices only (got device={device}); there is no CPU fallback")
    batch = batch.to(device)
    moved = []
    for m in (score_model, *extra_models):
        moved.append(m.to(device) if isinstance(m, nn.Module) else m)
    so3 = sdes["node_orientations"]
    if isinstance(so3, nn.Module):
        so3 = so3.to(device)
    return batch, device, so3, moved


def _prior(batch, sdes, so3, device):
    """denoiser.py:224-229: positions first, then orientations (kwarg evaluation order)."""
    pos_shape, rot_shape = tuple(batch["pos"].shape), tuple(batch["node_orientations"].shape)
    if isinstance(sdes["pos"], S.CosineVPSDE):
        pos = S.noise_randn(pos_shape, device)
    else:
        pos = sdes["pos"].prior_sampling(pos_shape, device=device)
    rot = so3.prior_sampling(rot_shape, device=device)
    return batch.replace(pos=pos.float().contiguous(), node_orientations=rot.float().contiguous())


def _dense(x, batch, lengths):
    """to_dense_batch(x, batch_idx)[0] for the stored controls (denoiser.py:334-335)."""
    lmax = max(lengths)
    if all(n == lmax for n in lengths):
        return x.reshape(len(lengths), lmax, *x.shape[1:])
    out = x.new_zeros((len(lengths), lmax) + tuple(x.shape[1:]))
    o = 0
    for g, n in enumerate(lengths):
        out[g, :n] = x[o:o + n]
        o += n
    return out


def _t(value: float, num_graphs: int, device):
    return torch.full((num_graphs,), value, device=device)


def _fields(sdes):
    return list(sdes.keys())


# ------------------------------------------------------------------------------------------------
# CUDA-graph replay of a whole sampler loop (dpm_solver; since r1l also the plain Euler-Maruyama and Heun loops).  All
# per-step quantities are kernel arguments fixed at capture time and the loops have no host synchronisation, so the
# network evaluations + frame updates (+ noise draws) of one call are a single graph launch.  Used when the same device-resident batch (same embedding tensors), model and SDEs
# come back: first call eager, second call captures, later calls replay.  SE3DIFF_B200_CUDA_GRAPH=0 disables it.
_GRAPHS: "collections.OrderedDict" = None  # type: ignore[assignment]
_GRAPH_SEEN: set = set()
_MAX_GRAPHS = 4
GRAPH_STATS = {"eager_first_sighting": 0, "captures": 0, "replays": 0}     # whole-loop graphs: what the calls of this process did


def _graph_key(batch, sdes, so3, score_model, num_steps, max_t, min_t, device, tag=("dpm",), control=None):
    import os

    from .models import DiGConditionalScoreModel

    if os.environ.get("SE3DIFF_B200_CUDA_GRAPH", "1") == "0" or S._HOST_NOISE or torch.cuda.is_current_stream_capturing():
        return None
    if not isinstance(score_model, DiGConditionalScoreModel) or not isinstance(so3, S.DiGSO3SDE):
        return None
    if control is not None and not isinstance(control, DiGConditionalScoreModel):
        return None
    if "single_embeds" not in batch or "pair_embeds" not in batch:
        return None
    sc = so3.score_function.score_scaling
    parts, keep = [], []
    for m in (score_model,) if control is None else (score_model, control):
        nn_ = m.model_nn
        if m.training and nn_.dropout_p > 0:
            return None             # the dropout path is torch autograd code with fresh masks per call: never replayed
        ctx = nn_._context(batch)   # by identity, else by exact value: a fresh Batch of the same sequence maps to the same context
        weights = nn_._layer_weights(torch.bfloat16 if nn_.precision == "bf16" else torch.float32)
        # `_struct_gen` moves when the model's cached tensors were REPLACED; in-place weight updates (optimizer steps on the
        # control model) refresh them inside their storage, so a captured loop stays valid across them
        parts.append((id(m), id(ctx), nn_.precision, nn_._struct_gen, m.training, nn_.x1d_proj[1].weight.data_ptr()))
        keep += [m, ctx, weights]
    key = (tuple(parts), tuple(batch["pos"].shape), id(so3), sc.data_ptr(), sc._version, so3.sigma_min, so3.sigma_max, so3.tol,
           type(sdes["pos"]).__name__, getattr(sdes["pos"], "s", None), num_steps, max_t, min_t, str(device), tag)
    return key, (*keep, so3, sc)


def _dpm_loop(batch, score_model, steps, device):
    B = batch.num_graphs
    pos, rot = batch["pos"], batch["node_orientations"]
    for st in steps:
        out = score_model(batch, _t(st.t, B, device))
        m_rot_t = out["node_orientations"]
        rot_u, pos_u = ops.frame_update_dpm_mid(rot, pos, m_rot_t, out["pos"], st.scalars)
        out_u = score_model(batch.replace(pos=pos_u, node_orientations=rot_u), _t(st.t_lambda, B, device))
        rot, pos = ops.frame_update_dpm_final(rot, pos, m_rot_t, out_u["node_orientations"], out_u["pos"], st.scalars)
        batch = batch.replace(pos=pos, node_orientations=rot)
    return batch


def _dpm_graphed(key, keep_alive, batch, score_model, steps, device):
    return _loop_graphed(key, keep_alive, batch, lambda static: _dpm_loop(static, score_model, steps, device), device)


_SIDE_STREAMS: dict = {}


def _score_and_control(score_fn, control_fn, device, distinct: bool = True):
    """`(score_fn(), control_fn())` for one state of a recording rollout.  The two evaluations are independent (denoiser.py:299-303
    calls them back to back on the same batch), and the 0.19 M-parameter control model is a chain of small, latency-bound
    launches: it runs on a side stream -- inside a whole-loop capture that makes it a parallel branch of the graph -- so that it
    fills the tails of the score model's kernels instead of queueing behind them.  Same kernels on the same inputs: results are
    bit-identical to the sequential order.  Ordering: the side stream waits for everything the main stream has enqueued (the
    state), the main stream waits for the side stream before it consumes the control; tensors either stream allocated are only
    released after such a join.  `SE3DIFF_B200_FORK_CONTROL=0` keeps one stream (measurement switch); so does `distinct=False`
    (one module passed as both models: its attention workspace must not be used by two streams at once)."""
    device = torch.device(device)
    if not distinct or device.type != "cuda" or os.environ.get("SE3DIFF_B200_FORK_CONTROL", "1") == "0":
        return score_fn(), control_fn()
    main = torch.cuda.current_stream(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    side = _SIDE_STREAMS.get(idx)
    if side is None:
        if torch.cuda.is_current_stream_capturing():
            return score_fn(), control_fn()
        side = _SIDE_STREAMS[idx] = torch.cuda.Stream(device)   # (a high-priority side stream measured the same: 291 vs 289 ms per rollout)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        u = control_fn()
    out = score_fn()
    main.wait_stream(side)
    return out, u


def _tree_map(fn, tree):
    if torch.is_tensor(tree):
        return fn(tree)
    if isinstance(tree, dict):
        return {k: _tree_map(fn, v) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(_tree_map(fn, v) for v in tree)
    return tree


def _loop_graphed(key, keep_alive, batch, loop_fn, device, returns_batch: bool = True):
    """Returns the loop's result through capture/replay, or None when this key has only been seen once.
    `loop_fn(batch)` is the whole sampler loop after the prior draw; it returns the denoised batch (`returns_batch`) or any tree
    (dict / list / tuple) of tensors -- the recording fine-tune rollouts return their stacked states, controls and Brownian
    increments.  Loops that draw noise (Euler-Maruyama, Heun) are captured too: torch's CUDA generator hands a captured graph
    its seed and offset at every replay, so a replay consumes the generator exactly like the eager loop (same seed => same
    trajectory; `test_em_heun_loop_graphs_match_eager`).  `keep_alive` are the objects whose device pointers the graph bakes in
    (context, cached weights, SDE tables); the entry owns them so that a replay can never read recycled memory.  Results are
    handed out as clones: the graph's own output buffers are overwritten by the next replay."""
    import collections

    global _GRAPHS
    if _GRAPHS is None:
        _GRAPHS = collections.OrderedDict()
    entry = _GRAPHS.get(key)
    if entry is None:
        if key not in _GRAPH_SEEN:      # first sighting: stay eager (one-shot callers never pay for a capture)
            if len(_GRAPH_SEEN) > 256:
                _GRAPH_SEEN.clear()
            _GRAPH_SEEN.add(key)
            GRAPH_STATS["eager_first_sighting"] += 1
            return None
        GRAPH_STATS["captures"] += 1
        entry = dict(pos_in=torch.empty_like(batch["pos"]), rot_in=torch.empty_like(batch["node_orientations"]), keep_alive=keep_alive)
        static = batch.replace(pos=entry["pos_in"], node_orientations=entry["rot_in"])
        entry["pos_in"].copy_(batch["pos"])
        entry["rot_in"].copy_(batch["node_orientations"])
        graph = torch.cuda.CUDAGraph()
        before = ops.launch_count()
        torch.cuda.synchronize(device)
        with torch.cuda.graph(graph):
            out = loop_fn(static)
        if returns_batch:
            out = (out["pos"], out["node_orientations"])
        # the captured kernels may read ANY field of the batch they were recorded on (graph index, pointers) and whatever device
        # tensors the loop closed over: the entry keeps both alive for as long as the graph can be replayed
        entry.update(graph=graph, out=out, launches=ops.launch_count() - before, static=static, loop_fn=loop_fn)
        ops.count_replayed_launches(-entry["launches"])     # recorded, not executed: the replay below is what runs
        _GRAPHS[key] = entry
        while len(_GRAPHS) > _MAX_GRAPHS:
            _GRAPHS.popitem(last=False)
    else:
        _GRAPHS.move_to_end(key)
    entry["pos_in"].copy_(batch["pos"])
    entry["rot_in"].copy_(batch["node_orientations"])
    entry["graph"].replay()
    GRAPH_STATS["replays"] += 1
    ops.count_replayed_launches(entry["launches"])
    out = _tree_map(torch.clone, entry["out"])
    return batch.replace(pos=out[0], node_orientations=out[1]) if returns_batch else out


@torch.no_grad()
def dpm_solver(*, batch, sdes, score_model, num_steps: int, max_t: float, min_t: float, device=None):
    """DPM-Solver-2 on positions + midpoint / extrapolated-score exp-map step on orientations
    (denoiser.py:634-764)."""
    assert max_t < 1.0
    batch, device, so3, (score_model,) = _prepare(batch, sdes, score_model, device)
    steps = schedule.dpm_schedule(sdes["pos"], so3, num_steps, max_t, min_t)
    batch = _prior(batch, sdes, so3, device)
    keyed = _graph_key(batch, sdes, so3, score_model, num_steps, max_t, min_t, device)
    if keyed is not None:
        done = _dpm_graphed(keyed[0], keyed[1], batch, score_model, steps, device)
        if done is not None:
            return done
    if not S._HOST_NOISE:
        return _dpm_loop(batch, score_model, steps, device)
    B = batch.num_graphs
    pos, rot = batch["pos"], batch["node_orientations"]
    for st in steps:
        out = score_model(batch, _t(st.t, B, device))
        m_rot_t = out["node_orientations"]
        if S._HOST_NOISE:  # the two randn_like of denoiser.py:80 that are multiplied by zero
            torch.randn(pos.shape[0], 3)
        rot_u, pos_u = ops.frame_update_dpm_mid(rot, pos, m_rot_t, out["pos"], st.scalars)
        out_u = score_model(batch.replace(pos=pos_u, node_orientations=rot_u), _t(st.t_lambda, B, device))
        if S._HOST_NOISE:
            torch.randn(pos.shape[0], 3)
        rot, pos = ops.frame_update_dpm_final(rot, pos, m_rot_t, out_u["node_orientations"], out_u["pos"], st.scalars)
        batch = batch.replace(pos=pos, node_orientations=rot)
    return batch


def _em_loop(batch, sdes, score_model, finetune_model, num_steps, max_t, min_t, device):
    batch, device, so3, (score_model, finetune_model) = _prepare(batch, sdes, score_model, device, (finetune_model,))
    steps = schedule.em_schedule(sdes["pos"], so3, num_steps, max_t, min_t)
    batch = _prior(batch, sdes, so3, device)
    B = batch.num_graphs
    fields = _fields(sdes)
    record = finetune_model is not None
    lengths = batch_lengths(batch) if record else None
    if not record and not S._HOST_NOISE:
        def plain(b):
            for st in steps:
                out = score_model(b, _t(st.t, B, device))
                z = {f: S.noise_randn((b["pos"].shape[0], 3), device) for f in fields}  # per-field draw order = sdes key order
                rot, pos, _, _ = ops.frame_update_em(b["node_orientations"], b["pos"], out["node_orientations"], out["pos"],
                                                     z["node_orientations"], z["pos"], st.scalars)
                b = b.replace(pos=pos, node_orientations=rot)
            return b

        keyed = _graph_key(batch, sdes, so3, score_model, num_steps, max_t, min_t, device, tag=("em", tuple(fields)))
        if keyed is not None:
            done = _loop_graphed(keyed[0], keyed[1], batch, plain, device)
            if done is not None:
                return done
        return plain(batch)
    def recorded(b):
        """The recording loop on stacked outputs: states [T+1, N, .], controls and Brownian increments [T, B, L, 3] per field."""
        pos_all, rot_all, us, dWs = [b["pos"]], [b["node_orientations"]], defaultdict(list), defaultdict(list)
        for st in steps:
            t = _t(st.t, B, device)
            pos, rot = b["pos"], b["node_orientations"]
            if record and not S._HOST_NOISE:
                out, u = _score_and_control(lambda: score_model(b, t), lambda: finetune_model(b, t), device, finetune_model is not score_model)
            else:
                out = score_model(b, t)
                u = finetune_model(b, t) if record else None
            z = {f: S.noise_randn((pos.shape[0], 3), device) for f in fields}  # per-field draw order = sdes key order
            rot, pos, dw_rot, dw_pos = ops.frame_update_em(
                rot, pos, out["node_orientations"], out["pos"], z["node_orientations"], z["pos"], st.scalars,
                u_rot=None if u is None else u["node_orientations"], u_pos=None if u is None else u["pos"], want_dw=record)
            b = b.replace(pos=pos, node_orientations=rot)
            if record:
                dw = {"pos": dw_pos, "node_orientations": dw_rot}
                for f in fields:
                    us[f].append(_dense(u[f], b, lengths))
                    dWs[f].append(_dense(dw[f], b, lengths))
                pos_all.append(pos)
                rot_all.append(rot)
        if not record:
            return dict(pos=b["pos"][None], rot=b["node_orientations"][None])
        return dict(pos=torch.stack(pos_all), rot=torch.stack(rot_all), us={f: torch.stack(us[f], dim=0) for f in fields},
                    dWs={f: torch.stack(dWs[f], dim=0) for f in fields})

    res = None
    if record and not S._HOST_NOISE:
        # the whole recording rollout (200 score + 200 control evaluations, noise draws, frame updates) as ONE graph launch: the
        # per-step host work (two model calls, bookkeeping) is what a rank waits for when several ranks share a host
        keyed = _graph_key(batch, sdes, so3, score_model, num_steps, max_t, min_t, device, tag=("em-record", tuple(fields)), control=finetune_model)
        if keyed is not None:
            res = _loop_graphed(keyed[0], keyed[1], batch, recorded, device, returns_batch=False)
    if res is None:
        res = recorded(batch)
    if not record:
        return batch.replace(pos=res["pos"][-1], node_orientations=res["rot"][-1])
    ts, _ = schedule.timesteps(max_t, min_t, num_steps)
    batches = [batch.replace(pos=res["pos"][i], node_orientations=res["rot"][i]) for i in range(res["pos"].shape[0])]
    return DenoisedSDEPath(batches=batches, timesteps=ts.to(device), us_batch=res["us"], dWs_batch=res["dWs"])


@torch.no_grad()
def euler_maruyama_predictor(*, batch, sdes, score_model, num_steps: int, max_t: float, min_t: float, device=None):
    """denoiser.py:206-264."""
    return _em_loop(batch, sdes, score_model, None, num_steps, max_t, min_t, device)


@torch.no_grad()
def euler_maruyama_predictor_finetune(*, batch, sdes, score_model, finetune_model, num_steps: int, max_t: float,
                                      min_t: float, device=None) -> DenoisedSDEPath:
    """denoiser.py:267-348: EM with the fine-tune control u in the drift; records every batch, u_t and
    dW_t = sqrt|dt| z as dense [T, B, L, 3]."""
    return _em_loop(batch, sdes, score_model, finetune_model, num_steps, max_t, min_t, device)


def _get_score(batch, sdes, score_model, t):
    """Score-model output -> scores (denoiser.py:169-203)."""
    bi = batch["batch"]
    out = score_model(batch, t)
    rot = out["node_orientations"] * sdes["node_orientations"].get_score_scaling(t, batch_idx=bi).unsqueeze(-1)
    _, std = sdes["pos"].marginal_prob(x=torch.ones_like(out["pos"]), t=t, batch_idx=bi)
    return {"node_orientations": rot, "pos": out["pos"] / std}


def _heun_finetune_loop(batch, sdes, score_model, finetune_model, num_steps, max_t, min_t, noise, device):
    """Rollout for the fine-tune objective: 3 score + 3 control evaluations per churned step, so the SDE algebra (composed
    here from the per-field predictor, i.e. the so3 / elementwise kernels) is a vanishing share; every random draw of the
    reference happens in the same order (the deterministic updates consume `randn_like` too, denoiser.py:80)."""
    ts, dts = schedule.timesteps(max_t, min_t, num_steps)
    ts_dev, dts_dev = ts.to(device), dts.to(device)
    fields = _fields(sdes)
    pred = {f: EulerMaruyamaPredictor(corruption=sdes[f], noise_weight=0.0) for f in fields}
    nois = {f: EulerMaruyamaPredictor(corruption=sdes[f], noise_weight=1.0) for f in fields}
    bi, B = batch["batch"], batch.num_graphs
    lengths = batch_lengths(batch)

    def both(state, time):
        if S._HOST_NOISE:
            return _get_score(state, sdes, score_model, time), finetune_model(state, time)
        return _score_and_control(lambda: _get_score(state, sdes, score_model, time), lambda: finetune_model(state, time), device,
                                  finetune_model is not score_model)

    def recorded(batch):
        """The whole rollout on stacked outputs.  Control flow uses the HOST copies of the time grid only (no device reads), so the
        loop can be captured as one CUDA graph."""
        pos_all, rot_all, us, dWs = [batch["pos"]], [batch["node_orientations"]], defaultdict(list), defaultdict(list)
        for i in range(num_steps):
            t = _t(float(ts[i]), B, device)
            t_next = t + dts_dev[i]
            churn = i > 0 and 0.0 < float(ts[i]) < 1.0
            t_hat = t - noise * dts_dev[i] if churn else t
            hat = batch.replace(**{f: nois[f].forward_sde_step(x=batch[f], t=t, dt=(t_hat - t)[0], batch_idx=bi)[0] for f in fields})
            sc_h, u_h = both(hat, t_hat)
            if churn:
                sc, u = both(batch, t)
            else:
                sc, u = sc_h, u_h
            dh = {f: pred[f].reverse_drift_and_diffusion(x=hat[f], t=t_hat, score=sc_h[f], finetune_score=u_h[f], batch_idx=bi)[0] for f in fields}
            step = (t_next - t_hat)[0]
            new = batch.replace(**{f: pred[f].update_given_drift_and_diffusion(x=hat[f], dt=step, drift=dh[f], diffusion=0.0)[1] for f in fields})
            if float(ts[i] + dts[i]) > 0.0:                    # = t_next, evaluated on the host in the same fp32 arithmetic
                sc_n, u_n = both(new, t_next)
                avg = {f: (pred[f].reverse_drift_and_diffusion(x=new[f], t=t_next, score=sc_n[f], finetune_score=u_n[f], batch_idx=bi)[0]
                           + dh[f]) / 2 for f in fields}
                new = batch.replace(**{f: pred[f].update_given_drift_and_diffusion(x=hat[f], dt=step, drift=avg[f], diffusion=0.0)[1] for f in fields})
            for f in fields:
                dW = nois[f].traceback_brownian_motion(x_next=new[f], x=batch[f], t=t, dt=dts_dev[i], score=sc[f], finetune_score=u[f],
                                                       batch_idx=bi)
                us[f].append(_dense(u[f], batch, lengths))
                dWs[f].append(_dense(dW, batch, lengths))
            batch = new
            pos_all.append(batch["pos"])
            rot_all.append(batch["node_orientations"])
        return dict(pos=torch.stack(pos_all), rot=torch.stack(rot_all), us={f: torch.stack(us[f], dim=0) for f in fields},
                    dWs={f: torch.stack(dWs[f], dim=0) for f in fields})

    res = None
    if not S._HOST_NOISE:
        keyed = _graph_key(batch, sdes, sdes["node_orientations"], score_model, num_steps, max_t, min_t, device,
                           tag=("heun-record", float(noise), tuple(fields)), control=finetune_model)
        if keyed is not None:
            res = _loop_graphed(keyed[0], keyed[1], batch, recorded, device, returns_batch=False)
    if res is None:
        res = recorded(batch)
    batches = [batch.replace(pos=res["pos"][i], node_orientations=res["rot"][i]) for i in range(res["pos"].shape[0])]
    return DenoisedSDEPath(batches=batches, timesteps=ts_dev, us_batch=res["us"], dWs_batch=res["dWs"])


def _heun_loop(batch, sdes, score_model, finetune_model, num_steps, max_t, min_t, noise, device):
    batch, device, so3, (score_model, finetune_model) = _prepare(batch, sdes, score_model, device, (finetune_model,))
    steps = schedule.heun_schedule(sdes["pos"], so3, num_steps, max_t, min_t, noise)
    batch = _prior(batch, sdes, so3, device)
    B = batch.num_graphs
    fields = _fields(sdes)
    if finetune_model is not None:
        return _heun_finetune_loop(batch, sdes, score_model, finetune_model, num_steps, max_t, min_t, noise, device)
    n = batch["pos"].shape[0]

    def draws():
        if S._HOST_NOISE:  # deterministic updates still consume randn_like in the reference
            for _ in fields:
                torch.randn(n, 3)

    def loop(batch):
        for st in steps:
            pos, rot = batch["pos"], batch["node_orientations"]
            z = {f: S.noise_randn((n, 3), device) for f in fields}
            rot_h, pos_h = ops.frame_heun_churn(rot, pos, z["node_orientations"], z["pos"], st.scalars)
            batch_hat = batch.replace(pos=pos_h, node_orientations=rot_h)
            out_h = score_model(batch_hat, _t(st.t_hat, B, device))
            draws()
            rot1, pos1 = ops.frame_heun_predict(rot_h, pos_h, out_h["node_orientations"], out_h["pos"], st.scalars)
            batch = batch.replace(pos=pos1, node_orientations=rot1)
            if st.correct:
                out_n = score_model(batch, _t(st.t_next, B, device))
                draws()
                rot2, pos2 = ops.frame_heun_correct(rot_h, pos_h, out_h["node_orientations"], out_h["pos"], pos1,
                                                    out_n["node_orientations"], out_n["pos"], st.scalars)
                batch = batch.replace(pos=pos2, node_orientations=rot2)
        return batch

    keyed = _graph_key(batch, sdes, so3, score_model, num_steps, max_t, min_t, device, tag=("heun", float(noise), tuple(fields)))
    if keyed is not None:                                   # (None under host noise, with a foreign score model, ...)
        done = _loop_graphed(keyed[0], keyed[1], batch, loop, device)
        if done is not None:
            return done
    return loop(batch)


@torch.no_grad()
def heun_denoiser(*, batch, sdes, score_model, num_steps: int, max_t: float, min_t: float, noise: float, device=None):
    """Karras-style churn + Heun second-order correction on both fields (denoiser.py:351-461)."""
    return _heun_loop(batch, sdes, score_model, None, num_steps, max_t, min_t, noise, device)


@torch.no_grad()
def heun_denoiser_finetune(*, batch, sdes, score_model, finetune_model, num_steps: int, max_t: float, min_t: float,
                           noise: float, device=None):
    """denoiser.py:464-620: Heun with the control in every drift; returns the path with dense `us`, `dWs` [T, B, L, 3].

    One deliberate difference: the reference appends the SAME in-place-mutated batch object every step
    (denoiser.py:518, 564, 588, 596), so all entries of its `batches` alias the final state; here `batches[i]` is the
    state after step i (what `finetune.py:338-393` needs to re-evaluate the control).  `batches[-1]`, `us_batch`,
    `dWs_batch` and `timesteps` equal the reference's."""
    return _heun_loop(batch, sdes, score_model, finetune_model, num_steps, max_t, min_t, noise, device)


def sde_dpm_solver_finetune(*, batch, sdes, score_model, finetune_model, num_steps, max_t, min_t, device=None):
    """Unimplemented stub in the reference as well (denoiser.py:767-777 has body `...`)."""
    raise NotImplementedError("sde_dpm_solver_finetune is an empty stub in the reference (denoiser.py:767-777)")


class EulerMaruyamaPredictor:
    """Per-field predictor object used directly by se3diff/train.py:54-70 and se3diff/finetune.py:33-56
    (denoiser.py:30-166).  Works on `[n,3,3]` rotations (SO3SDE corruption) or `[n,d]` positions
    (CosineVPSDE corruption) with per-element t."""

    def __init__(self, *, corruption, noise_weight: float = 1.0, marginal_concentration_factor: float = 1.0):
        self.corruption, self.noise_weight, self.marginal_concentration_factor = corruption, noise_weight, marginal_concentration_factor

    def reverse_drift_and_diffusion(self, *, x, t, score, finetune_score=None, batch_idx=None):
        w = 0.5 * self.marginal_concentration_factor * (1 + self.noise_weight**2)
        drift, diffusion = self.corruption.sde(x=x, t=t, batch_idx=batch_idx)
        drift = drift - diffusion**2 * score * w
        if finetune_score is not None:
            drift = drift + diffusion * finetune_score * w
        return drift, diffusion

    def update_given_drift_and_diffusion(self, *, x, dt, drift, diffusion):
        z = S.noise_randn(tuple(drift.shape), drift.device)
        dW = self.noise_weight * torch.sqrt(dt.abs()) * z
        if isinstance(self.corruption, S.SO3SDE):
            mean = ops.so3_compose_rotvec(x, drift * dt, self.corruption.tol)
            sample = ops.so3_compose_rotvec(mean, diffusion * dW, self.corruption.tol)
        elif isinstance(self.corruption, S.CosineVPSDE):
            mean = x + drift * dt
            sample = mean + diffusion * dW
        else:
            raise NotImplementedError(f"Update for {type(self.corruption)} not implemented.")
        return sample, mean, dW

    def update_given_score(self, *, x, t, dt, score, finetune_score=None, batch_idx=None):
        drift, diffusion = self.reverse_drift_and_diffusion(x=x, t=t, score=score, finetune_score=finetune_score, batch_idx=batch_idx)
        return self.update_given_drift_and_diffusion(x=x, dt=dt, drift=drift, diffusion=diffusion)

    def forward_sde_step(self, *, x, t, dt, batch_idx=None):
        drift, diffusion = self.corruption.sde(x=x, t=t, batch_idx=batch_idx)
        return self.update_given_drift_and_diffusion(x=x, dt=dt, drift=drift, diffusion=diffusion)

    def traceback_brownian_motion(self, *, x_next, x, t, dt, score, finetune_score=None, batch_idx=None):
        drift, diffusion = self.reverse_drift_and_diffusion(x=x, t=t, score=score, finetune_score=finetune_score, batch_idx=batch_idx)
        if S._HOST_NOISE:   # the reference obtains the mean through update_given_drift_and_diffusion (denoiser.py:152-157),
            torch.randn(*drift.shape)   # which draws a randn_like(drift) it then discards: keep the stream aligned
        if isinstance(self.corruption, S.SO3SDE):
            mean = ops.so3_compose_rotvec(x, drift * dt, self.corruption.tol)
            return ops.so3_rel_log(mean, x_next) / diffusion
        if isinstance(self.corruption, S.CosineVPSDE):
            return (x_next - (x + drift * dt)) / diffusion
        raise NotImplementedError(f"Update for {type(self.corruption)} not implemented.")
