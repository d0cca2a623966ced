"""The step either side of the sampling path: batch assembly from a sequence's embeddings and the
on-disk result format with its resume rule -- host mirror of `bioemu/src/bioemu/sample.py:143-183, 186-236, 288-308` and
`bioemu/src/bioemu/utils.py:13-28`.  The ColabFold embedding retrieval in front of it (network, external tools) is out of
scope: the embeddings come in as arrays or as the `.npy` files the reference caches.
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import torch

from .chemgraph import Batch, ChemGraph, complete_graph_edge_index


def format_npz_samples_filename(start_id: int, num_samples: int) -> str:
    """utils.py:13-17: `batch_<start:07d>_<end:07d>.npz`, end exclusive."""
    return f"batch_{start_id:07d}_{start_id + num_samples:07d}.npz"


def count_samples_in_output_dir(output_dir) -> int:
    """utils.py:20-28: samples already on disk, from the file names alone (the resume rule of sample.py:288-289)."""
    total = 0
    for p in Path(output_dir).glob("batch_*.npz"):
        lo, hi = p.stem.split("_")[1:]
        total += int(hi) - int(lo)
    return total


def _load(x):
    if isinstance(x, (str, os.PathLike)):
        x = np.load(x)
    return x if torch.is_tensor(x) else torch.from_numpy(np.asarray(x))


def generate_chemgraph(*, sequence: str, single_embeds, pair_embeds) -> ChemGraph:
    """sample.py:143-183 from the embeddings on: `single_embeds [L, 384]`, `pair_embeds [L, L, 128]` (arrays, tensors or
    the cached .npy paths) -> ChemGraph with NaN frames, the row-major complete graph and the flattened pair features."""
    seq_len = len(sequence)
    single, pair = _load(single_embeds), _load(pair_embeds)
    if single.shape[0] != seq_len or tuple(pair.shape[:2]) != (seq_len, seq_len):
        raise ValueError(f"embeddings {tuple(single.shape)} / {tuple(pair.shape)} do not match a sequence of length {seq_len}")
    return ChemGraph(pos=torch.full((seq_len, 3), float("nan")), node_orientations=torch.full((seq_len, 3, 3), float("nan")),
                     edge_index=complete_graph_edge_index(seq_len), single_embeds=single,
                     pair_embeds=pair.reshape(seq_len**2, pair.shape[-1]))


@torch.no_grad()
def generate_batch(*, chemgraph: ChemGraph, bundle, batch_size: int, device=None, seed: int | None = None) -> dict:
    """sample.py:186-236: seed, B copies of the context graph, one denoiser call, frames back on the host as
    `pos [B, L, 3]`, `node_orientations [B, L, 3, 3]`.  `bundle` = (sdes, score_model, denoiser) as `load_bundle` returns."""
    if seed is not None:
        torch.manual_seed(seed)
    batch = Batch.from_data_list([chemgraph for _ in range(batch_size)])
    sdes, score_model, denoiser = bundle
    out = denoiser(batch=batch, sdes=sdes, score_model=score_model, device=device)
    graphs = out.to_data_list()
    return {"pos": torch.stack([g.pos for g in graphs]).to("cpu"),
            "node_orientations": torch.stack([g.node_orientations for g in graphs]).to("cpu")}


def sample_to_dir(*, sequence: str, chemgraph: ChemGraph, output_dir, num_samples: int, bundle, batch_size: int, device=None) -> list:
    """The batch loop of `sample()` (sample.py:288-308): resumes after the samples already in `output_dir`, seeds every
    batch with its global sample offset, writes one npz per batch with keys `pos`, `node_orientations`, `sequence`.
    Returns the paths written by this call."""
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    existing = count_samples_in_output_dir(output_dir)
    written = []
    for seed in range(existing, num_samples, batch_size):
        n = min(batch_size, num_samples - seed)
        path = output_dir / format_npz_samples_filename(seed, n)
        if path.exists():
            raise ValueError(f"Not sure why {path} already exists when so far only {existing} samples have been generated.")
        batch = generate_batch(chemgraph=chemgraph, bundle=bundle, batch_size=n, device=device, seed=seed)
        np.savez(path, **{k: v.cpu().numpy() for k, v in batch.items()}, sequence=sequence)
        written.append(path)
    return written


def load_samples(output_dir, sequence: str | None = None):
    """Concatenated ensemble of a results directory, as sample.py:310-320 reads it back."""
    files = sorted(Path(output_dir).glob("batch_*.npz"))
    seqs = {np.load(f)["sequence"].item() for f in files}
    if sequence is not None and seqs != {sequence}:
        raise ValueError(f"Expected all sequences to be {sequence}, but got {seqs}")
    pos = torch.tensor(np.concatenate([np.load(f)["pos"] for f in files]))
    rot = torch.tensor(np.concatenate([np.load(f)["node_orientations"] for f in files]))
    return pos, rot
