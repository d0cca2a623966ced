"""TEST INFRASTRUCTURE -- imports the UNMODIFIED reference in place (build container only).

`/root/reference` does not exist on the GPU box; nothing that runs there may import this module.
It is used by `oracle/gen_golden.py` (to mint tests/golden/*.npz) and by the `live reference`
tests, which skip when the reference tree is absent.
"""
from __future__ import annotations

import os
import sys

REF_ROOT = os.environ.get("SE3DIFF_REFERENCE", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_pyg_shim")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "bioemu", "src", "bioemu"))


def load():
    """Returns a namespace with the reference modules (torch_geometric replaced by the shim)."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    for p in (os.path.join(REF_ROOT, "bioemu", "src"), REF_ROOT, _SHIM):
        if p not in sys.path:
            sys.path.insert(0, p)
    import types

    import bioemu.chemgraph as chemgraph
    import bioemu.denoiser as denoiser
    import bioemu.models as models
    import bioemu.sde_lib as sde_lib
    import bioemu.so3_sde as so3_sde
    import bioemu.structure_module as structure_module
    from torch_geometric.data import Batch

    ns = types.SimpleNamespace(
        chemgraph=chemgraph, denoiser=denoiser, models=models, sde_lib=sde_lib, so3_sde=so3_sde,
        structure_module=structure_module, Batch=Batch, root=REF_ROOT,
    )
    try:
        import se3diff.finetune as toy_finetune
        import se3diff.models as toy_models
        import se3diff.train as toy_train

        ns.toy_models, ns.toy_train, ns.toy_finetune = toy_models, toy_train, toy_finetune
    except Exception:  # ppft etc. are optional for the goldens
        pass
    return ns


def make_batch(ns, single, pair, lengths, pos=None, rot=None, extra=None):
    """Batch of ChemGraphs as sample.py:143-183,223 builds it.  single: list of [L,384]; pair: list of
    [L,L,128]."""
    import torch

    from .score_model import make_edge_index

    graphs = []
    o = 0
    for g, n in enumerate(lengths):
        kw = dict(
            pos=torch.full((n, 3), float("nan")) if pos is None else pos[o : o + n],
            node_orientations=torch.full((n, 3, 3), float("nan")) if rot is None else rot[o : o + n],
            edge_index=make_edge_index(n),
            single_embeds=single[g],
            pair_embeds=pair[g].reshape(n * n, -1),
        )
        if extra is not None:
            for k, v in extra.items():
                kw[k] = v[o : o + n]
        graphs.append(ns.chemgraph.ChemGraph(**kw))
        o += n
    return ns.Batch.from_data_list(graphs)


def extract_functions(rel_path: str, names, namespace: dict) -> dict:
    """Executes the named top-level function definitions of a reference file, verbatim and in memory, inside `namespace`
    -- for modules whose import chain needs packages that are absent here (hydra, mdtraj, Bio): the function bodies still
    are the reference's own code.  Returns the namespace."""
    import ast

    path = os.path.join(REF_ROOT, rel_path)
    tree = ast.parse(open(path).read(), filename=path)
    picked = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise RuntimeError(f"{rel_path}: functions not found: {sorted(missing)}")
    for n in picked:
        n.decorator_list = []
    exec(compile(ast.Module(body=picked, type_ignores=[]), path, "exec"), namespace)
    return namespace
