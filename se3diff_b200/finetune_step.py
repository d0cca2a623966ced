"""One fine-tune step around the sampling path -- host mirror of `bioemu/src/bioemu/finetune.py:291-514` (rollout, chunked
loss + backward, validation loss) and of the observable `bioemu/src/bioemu/observables/folding_stability.py:40-194`.

Division of labour: the rollout is the CUDA path (`euler_maruyama_predictor_finetune`: 200 no-grad
evaluations of the 31 M-parameter score model); the observable is one kernel; the loss side re-evaluates the 0.19 M-parameter
control model WITH gradients on the stored states (`DiGConditionalScoreModel._forward_torch`) and reduces with the path
functionals of `pathwise.py`; in a multi-GPU job the flat gradient is all-reduced right before `optimizer.step()`
(`distributed.allreduce_gradients`, finetune.py:625).

Not reproduced: the reference's `ddr_debug/` dump of every batch to npz + PDB/XTC inside `compute_finetune_loss`
(finetune.py:419-447, developer debugging that needs mdtraj) and `check_protein_valid`.
"""
from __future__ import annotations

import math
from collections import defaultdict
from typing import NamedTuple

import numpy as np
import torch

from . import ops, pathwise
from .chemgraph import Batch, batch_lengths
from .denoiser import DenoisedSDEPath, _dense

K_BOLTZMANN = 0.001987203599772605  # kcal / (mol K), folding_stability.py:19


class FinetuneBundle(NamedTuple):
    """finetune.py:125-139: what `load_finetune_bundle` returns."""

    sdes: dict
    score_model: torch.nn.Module
    finetune_model: torch.nn.Module
    denoiser: object
    h_func: object


# ---- observable ------------------------------------------------------------------------------------------------
def load_reference_ca_coords(ref_path, device=None) -> torch.Tensor:
    """C-alpha coordinates of the first model of a PDB file in nm (folding_stability.py:22-49), parsed from the fixed
    columns of the ATOM records (the reference goes through Bio.PDB, which is not a dependency here)."""
    coords, seen = [], set()
    with open(ref_path) as f:
        for line in f:
            rec = line[:6]
            if rec == "ENDMDL":
                break
            if rec != "ATOM  " or line[12:16].strip() != "CA":
                continue
            key = (line[21], line[22:27])          # chain, residue number + insertion code
            if key in seen or line[16] not in (" ", "A"):
                continue
            seen.add(key)
            coords.append([float(line[30:38]) / 10, float(line[38:46]) / 10, float(line[46:54]) / 10])
    return torch.tensor(np.asarray(coords, dtype=np.float32), device=device)


def compute_folded_proportion(coords, ref_coords, k: float = -24.0, d_0: float = 0.4, tol: float = 1e-7):
    """folding_stability.py:52-81 -> p_folded [B]; one kernel (se3_folded_proportion)."""
    return ops.folded_proportion(coords, ref_coords, k, d_0, tol)


def compute_dG(p_folded, temperature: float = 298.0, tol: float = 1e-7):
    """folding_stability.py:84-101."""
    p = torch.clamp(p_folded.mean(), min=tol, max=1.0 - tol)
    return -K_BOLTZMANN * temperature * torch.log(p / (1.0 - p))


def compute_folded_proportion_from_dG(dG, temperature: float = 298.0):
    """folding_stability.py:104-117."""
    return torch.sigmoid(-dG / (K_BOLTZMANN * temperature))


class FoldingStability:
    """h-function of the fine-tune objective (folding_stability.py:120-194): `h(batch, sequence) -> [B, 1]` folded
    probability per sample.  The reference hard-wires `structures/2vwf_trimmed_SH3.pdb`; here the reference structure is
    given as a path, as coordinates, or per sequence through `ref_paths`."""

    def __init__(self, k: float = -24.0, d_0: float = 0.4, tol: float = 1e-7, ref_path=None, ref_coords=None, ref_paths: dict | None = None):
        self.k, self.d_0, self.tol = k, d_0, tol
        self.ref_path, self.ref_coords, self.ref_paths = ref_path, ref_coords, ref_paths or {}

    def sequence_to_ref_path(self, sequence: str):
        path = self.ref_paths.get(sequence, self.ref_path)
        if path is None:
            raise FileNotFoundError(f"no reference structure configured for sequence {sequence!r}")
        return path

    def __call__(self, batch, sequence: str) -> torch.Tensor:
        lengths = batch_lengths(batch)
        coords = _dense(batch["pos"], batch, lengths)                                   # [B, L, 3]
        ref = self.ref_coords if self.ref_coords is not None else load_reference_ca_coords(self.sequence_to_ref_path(sequence))
        return compute_folded_proportion(coords, ref.to(coords.device), self.k, self.d_0, self.tol).unsqueeze(-1)


# ---- rollout ----------------------------------------------------------------------------------------------------
@torch.no_grad()
def generate_finetune_batch(*, chemgraph, finetune_bundle: FinetuneBundle, batch_size: int, device=None, seed: int | None = None) -> DenoisedSDEPath:
    """finetune.py:291-335 from the context graph on (see sampling_io.generate_chemgraph)."""
    if seed is not None:
        torch.manual_seed(seed)
    batch = Batch.from_data_list([chemgraph for _ in range(batch_size)])
    sdes, score_model, finetune_model, denoiser, _ = finetune_bundle
    return denoiser(batch=batch, sdes=sdes, score_model=score_model, finetune_model=finetune_model, device=device)


# ---- loss -------------------------------------------------------------------------------------------------------
def _chunk_update(batches, timesteps, dts, dWs_batch, int_u_u_dt_sg, hs, h_stars, finetune_model, fields, batch_size, device=None,
                  lambda_: float = 0.1, tol: float = 1e-7):
    """finetune.py:338-393: re-evaluate the control with gradients on a chunk of stored states, form the chunk's share of the
    estimator and back-propagate it (gradients accumulate on `finetune_model`)."""
    us = defaultdict(list)
    ts_host = timesteps.detach().cpu()
    # One differentiable pass over the whole chunk when the stored states share one context (B copies of one sequence): the
    # chunk's K control evaluations are independent, so K launch-bound passes of the 0.19 M-parameter model become one of batch
    # K * B (the reference loops, finetune.py:352-361; equal up to floating-point summation order).
    stacked = None
    if hasattr(finetune_model, "forward_stacked") and not (finetune_model.training and getattr(finetune_model.model_nn, "dropout_p", 0) > 0):
        stacked = finetune_model.forward_stacked(list(batches), timesteps.to(device=device, dtype=torch.float32))
    if stacked is not None:
        lengths = batch_lengths(batches[0])
        for f in fields:
            us[f] = [_dense(stacked[f][i], batches[0], lengths) for i in range(len(batches))]
    else:
        for i, batch in enumerate(batches):
            t = torch.full((batch_size,), float(ts_host[i]), device=device)
            u_t = finetune_model(batch, t)
            lengths = batch_lengths(batch)
            for f in fields:
                us[f].append(_dense(u_t[f], batch, lengths))
    us_flat = {f: torch.stack(us[f], dim=0).flatten(-2, -1) for f in fields}
    dWs_flat = {f: dWs_batch[f].flatten(-2, -1) for f in fields}
    int_dws = sum(pathwise.compute_int_dws(us=us_flat[f], dWs=dWs_flat[f]) for f in fields)
    int_u_u_dt = sum(pathwise.compute_int_u_u_dt(us=us_flat[f], dts=dts) for f in fields)
    loss_ev = pathwise.compute_ev_loss(ws=int_dws, hs=hs, h_stars=h_stars, from_int_dws=True, use_stab=True, tol=tol)
    loss_kl = pathwise.compute_kl_loss(ws=int_dws, int_u_u_dt=int_u_u_dt, int_u_u_dt_sg=int_u_u_dt_sg, from_int_dws=True, use_rloo=True)
    (loss_ev + lambda_ * loss_kl).backward()


def compute_finetune_loss(*, sequence: str, h_stars, finetune_bundle: FinetuneBundle, denoised_sde_path: DenoisedSDEPath, batch_size: int,
                          device=None, for_grad: bool = True, micro_batch_size: int = 1, lambda_: float = 0.1, tol: float = 1e-7):
    """finetune.py:396-514: with `for_grad` the gradient of the fine-tune objective is accumulated on the control model
    chunk by chunk; the return value is the plain (validation) loss E-term + lambda * KL-term."""
    if batch_size < 2:
        raise ValueError("Batch size must be at least 2 for estimating variances.")
    sdes, _, finetune_model, _, h_func = finetune_bundle
    batches, timesteps, us_sg, dWs = denoised_sde_path
    fields = list(sdes.keys())
    with torch.no_grad():
        hs = h_func(batch=batches[-1], sequence=sequence)                                # [B, K]
    h_stars = h_stars.to(hs.device)
    dts = torch.diff(timesteps)
    num_steps = len(dts)
    if micro_batch_size > num_steps:
        raise ValueError(f"micro_batch_size ({micro_batch_size}) must be less than or equal to num_steps ({num_steps}).")
    int_u_u_dt_sg = sum(pathwise.compute_int_u_u_dt(us=us_sg[f].flatten(-2, -1), dts=dts) for f in fields)
    if for_grad:
        with torch.enable_grad():
            for i in range(math.ceil(num_steps / micro_batch_size)):
                lo, hi = i * micro_batch_size, min((i + 1) * micro_batch_size, num_steps)
                _chunk_update(batches=batches[lo:hi], timesteps=timesteps[lo:hi], dts=dts[lo:hi], dWs_batch={f: dWs[f][lo:hi] for f in fields},
                              int_u_u_dt_sg=int_u_u_dt_sg, hs=hs, h_stars=h_stars, finetune_model=finetune_model, fields=fields,
                              batch_size=batch_size, device=device, lambda_=lambda_, tol=tol)
    ws = torch.ones_like(int_u_u_dt_sg)
    loss_ev = pathwise.compute_ev_loss(ws=ws, hs=hs, h_stars=h_stars, from_int_dws=False, use_stab=False, tol=tol)
    loss_kl = pathwise.compute_kl_loss(ws=ws, int_u_u_dt=int_u_u_dt_sg, int_u_u_dt_sg=int_u_u_dt_sg, from_int_dws=False, use_rloo=False)
    return loss_ev + lambda_ * loss_kl


def finetune_step(*, sequence: str, chemgraph, h_stars, finetune_bundle: FinetuneBundle, optimizer, batch_size: int, device=None,
                  micro_batch_size: int = 1, lambda_: float = 0.1, tol: float = 1e-7, seed: int | None = None):
    """Body of the training loop for one sequence (finetune.py:599-626): rollout, loss + backward, gradient exchange across
    ranks (if a process group is up), optimizer step.  Returns the detached validation loss."""
    from . import distributed

    optimizer.zero_grad()
    path = generate_finetune_batch(chemgraph=chemgraph, finetune_bundle=finetune_bundle, batch_size=batch_size, device=device, seed=seed)
    loss = compute_finetune_loss(sequence=sequence, h_stars=h_stars, finetune_bundle=finetune_bundle, denoised_sde_path=path,
                                 batch_size=batch_size, device=device, for_grad=True, micro_batch_size=micro_batch_size, lambda_=lambda_, tol=tol)
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        distributed.allreduce_gradients(finetune_bundle.finetune_model.parameters())
    optimizer.step()
    return loss.detach()
