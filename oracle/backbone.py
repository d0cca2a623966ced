"""TEST INFRASTRUCTURE -- numpy restatement of the physicality criteria of `convert_chemgraph.py:296-345`, which the
reference evaluates through mdtraj (absent here, so this part is PARITY UNPINNED against mdtraj itself; it follows the
documented semantics: `compute_contacts(scheme="ca")` on sequential pairs, C-N distances of sequential pairs, and
`compute_contacts()` defaults = closest heavy-atom distance over all residue pairs at least three apart)."""
from __future__ import annotations

import numpy as np

GLY = 7


def physicality_statistics(atoms: np.ndarray, aatype: np.ndarray) -> np.ndarray:
    """atoms [B, L, 5, 3] (N, CA, C, CB, O) -> [B, 3]: max seq CA-CA, max seq C-N, min heavy distance for |i-j| >= 3."""
    B, L = atoms.shape[:2]
    out = np.zeros((B, 3))
    present = np.ones((L, 5), dtype=bool)
    present[aatype == GLY, 3] = False
    for b in range(B):
        a = atoms[b].astype(np.float64)
        out[b, 0] = np.linalg.norm(a[:-1, 1] - a[1:, 1], axis=-1).max()
        out[b, 1] = np.linalg.norm(a[:-1, 2] - a[1:, 0], axis=-1).max()
        best = np.inf
        for i in range(L):
            for j in range(i + 3, L):
                d = np.linalg.norm(a[i][present[i]][:, None] - a[j][present[j]][None], axis=-1)
                best = min(best, d.min())
        out[b, 2] = best
    return out
