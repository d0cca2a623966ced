"""Frames -> backbone atoms and the physicality filter: host mirror of `bioemu/src/bioemu/convert_chemgraph.py:139-293`
(`get_atom37_from_frames`, `compute_backbone`, `_adjust_oxygen_pos`) and `:296-395` (`_filter_unphysical_traj_masks`,
`_get_physical_traj_indices`), batched over the ensemble on the GPU (the reference loops over samples in Python and filters
through mdtraj on the CPU).  Writing PDB / XTC files (`save_pdb_and_xtc`, `_write_pdb`: mdtraj, modelcif) stays outside.
"""
from __future__ import annotations

import torch

from . import ops

RESTYPES = "ARNDCQEGHILKMFPSTWYV"                      # openfold/np/residue_constants.py: restypes
_ORDER = {c: i for i, c in enumerate(RESTYPES)}
_BB_SLOTS = 5                                          # atom37 order starts N, CA, C, CB, O


def sequence_to_aatype(sequence: str, device=None) -> torch.Tensor:
    """`restype_order.get(x, 0)` per residue (convert_chemgraph.py:173-175): unknown letters map to alanine."""
    return torch.tensor([_ORDER.get(x, 0) for x in sequence], device=device)


def backbone_atoms_batch(pos: torch.Tensor, node_orientations: torch.Tensor, sequence: str) -> torch.Tensor:
    """pos [B, L, 3] in the unit the atoms are wanted in relative to Angstrom-scaled ideal geometry (the reference passes
    Angstrom), node_orientations [B, L, 3, 3] -> [B, L, 5, 3] = N, CA, C, CB, O."""
    return ops.backbone_atoms(pos, node_orientations, sequence_to_aatype(sequence, pos.device))


def get_atom37_from_frames(pos: torch.Tensor, node_orientations: torch.Tensor, sequence: str):
    """convert_chemgraph.py:139-186 for one structure: (atom_37 [L, 37, 3] Angstrom, atom_37_mask [L, 37], aatype [L])."""
    assert pos.dim() == 2 and pos.shape[1] == 3 and tuple(node_orientations.shape[1:]) == (3, 3)
    assert len(sequence) == pos.shape[0] == node_orientations.shape[0]
    aatype = sequence_to_aatype(sequence, pos.device)
    bb = ops.backbone_atoms(pos[None], node_orientations[None], aatype)[0]
    atom_37 = torch.zeros(pos.shape[0], 37, 3, device=pos.device)
    atom_37[:, :_BB_SLOTS] = bb
    mask = torch.zeros(pos.shape[0], 37, dtype=torch.bool, device=pos.device)
    # compute_backbone takes `any(pos != 0)` BEFORE the oxygen is imputed: the oxygen of the zero-torsion construction is a
    # generic non-zero point, the glycine CB is exactly zero (convert_chemgraph.py:201)
    mask[:, :_BB_SLOTS] = torch.any(bb != 0, dim=-1)
    mask[:, 4] = True
    return atom_37, mask, aatype


def physicality_statistics(pos_nm: torch.Tensor, node_orientations: torch.Tensor, sequence: str) -> torch.Tensor:
    """[B, 3] in Angstrom: max sequential CA-CA, max sequential C-N, min heavy-atom distance between residues >= 3 apart."""
    aatype = sequence_to_aatype(sequence, pos_nm.device)
    pos = pos_nm * 10.0
    pos = pos - pos.mean(dim=1, keepdim=True)            # save_pdb_and_xtc centres every structure (convert_chemgraph.py:427-428)
    return ops.physicality(ops.backbone_atoms(pos, node_orientations, aatype), aatype)


def filter_unphysical_masks(pos_nm, node_orientations, sequence: str, max_ca_seq_distance: float = 4.5, max_cn_seq_distance: float = 2.0,
                            clash_distance: float = 1.0):
    """The three per-sample masks of `_filter_unphysical_traj_masks` (convert_chemgraph.py:296-345)."""
    s = physicality_statistics(pos_nm, node_orientations, sequence)
    return s[:, 0] < max_ca_seq_distance, s[:, 1] < max_cn_seq_distance, s[:, 2] > clash_distance


def get_physical_sample_indices(pos_nm, node_orientations, sequence: str, max_ca_seq_distance: float = 4.5, max_cn_seq_distance: float = 2.0,
                                clash_distance: float = 1.0, strict: bool = False) -> torch.Tensor:
    """`_get_physical_traj_indices` (convert_chemgraph.py:348-370): indices of the samples passing all three criteria."""
    a, b, c = filter_unphysical_masks(pos_nm, node_orientations, sequence, max_ca_seq_distance, max_cn_seq_distance, clash_distance)
    ok = a & b & c
    if strict:
        assert int(ok.sum()) > 0, "Ended up with empty trajectory"
    return torch.nonzero(ok).flatten()


def backbone_trajectory(pos_nm, node_orientations, sequence: str, filter_samples: bool = True):
    """What `save_pdb_and_xtc` hands to mdtraj (convert_chemgraph.py:404-458) before superposition: heavy backbone atoms of
    every (kept) sample in nm, centred, as [B', n_atoms, 3] plus the kept indices.  Atom order per residue N, CA, C, CB, O
    with the glycine CB left out."""
    aatype = sequence_to_aatype(sequence, pos_nm.device)
    pos = pos_nm * 10.0
    pos = pos - pos.mean(dim=1, keepdim=True)
    atoms = ops.backbone_atoms(pos, node_orientations, aatype)
    keep = torch.arange(pos.shape[0], device=pos.device)
    if filter_samples:
        s = ops.physicality(atoms, aatype)
        keep = torch.nonzero((s[:, 0] < 4.5) & (s[:, 1] < 2.0) & (s[:, 2] > 1.0)).flatten()
    present = torch.ones(len(sequence), _BB_SLOTS, dtype=torch.bool, device=pos.device)
    present[aatype == _ORDER["G"], 3] = False
    return atoms[keep][:, present] * 0.1, keep
