"""Batch container at the drop-in boundary (reference: bioemu/src/bioemu/chemgraph.py:12-31 on top of
torch_geometric.data.Data/Batch, third-party, pinned ==2.6.1 in environment.yml:261).

When torch_geometric is importable the reference's own `ChemGraph(Data)` / PyG `Batch` objects work
unchanged with this package: the samplers and the score model only use the duck-typed surface listed
in SURVEY.md section 8b (`pos, node_orientations, edge_index, single_embeds, pair_embeds, batch, ptr /
num_graphs, __getitem__/__setitem__/__contains__, replace(**kw), to(device), to_data_list()`).
This module provides that same surface without PyG, for boxes where it is absent.
"""
from __future__ import annotations

from typing import Any

import torch

_NODE_KEYS = ("pos", "node_orientations", "single_embeds", "pos_is_known")


class ChemGraph:
    """Attribute bag with `replace()` shallow-copy semantics (chemgraph.py:21-31)."""

    def __init__(self, **fields: Any):
        object.__setattr__(self, "_fields", dict(fields))

    # -- mapping / attribute protocol --------------------------------------------------------------
    def __getattr__(self, key):
        f = object.__getattribute__(self, "_fields")
        if key in f:
            return f[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self._fields[key] = value

    def __getitem__(self, key):
        return self._fields[key]

    def __setitem__(self, key, value):
        self._fields[key] = value

    def __contains__(self, key):
        return key in self._fields

    def keys(self):
        return list(self._fields.keys())

    def items(self):
        return list(self._fields.items())

    def _clone_meta(self, fields):
        out = self.__class__.__new__(self.__class__)
        object.__setattr__(out, "_fields", fields)
        for k, v in self.__dict__.items():
            if k != "_fields":
                object.__setattr__(out, k, v)
        return out

    def replace(self, **kwargs: Any):
        f = dict(self._fields)
        f.update(kwargs)
        return self._clone_meta(f)

    def to(self, device, non_blocking: bool = False):
        return self._clone_meta({k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
                                 for k, v in self._fields.items()})

    @property
    def num_nodes(self) -> int:
        for k in _NODE_KEYS:
            v = self._fields.get(k)
            if torch.is_tensor(v):
                return int(v.shape[0])
        return 0


class _Replicated:
    """Placeholder for a concatenated field of a batch that holds B references to ONE graph: the concatenation is only built
    when somebody reads the field on the host.  sample.py:223 and finetune.py:325 build such a batch per call and hand it
    straight to the denoiser, which moves it to the device -- there the graph is shipped once and replicated on the GPU, so the
    B-fold host copy (231 MB per fine-tune step at L = 84, B = 64) is never needed."""

    __slots__ = ("src", "copies", "is_edge_index", "nodes")

    def __init__(self, src, copies, is_edge_index, nodes):
        self.src, self.copies, self.is_edge_index, self.nodes = src, copies, is_edge_index, nodes

    def build(self):
        if self.is_edge_index:
            e = self.src.shape[1]
            return self.src.repeat(1, self.copies) + (torch.arange(self.copies, device=self.src.device) * self.nodes).repeat_interleave(e)
        return self.src.repeat(self.copies, *([1] * (self.src.dim() - 1)))


class Batch(ChemGraph):
    """Concatenated graphs with PyG's `batch`/`ptr` bookkeeping (Batch.from_data_list as used at
    sample.py:223 and finetune.py:325)."""

    # -- lazily concatenated fields of a replicated batch -------------------------------------------------------------------
    def _materialise(self, key, value):
        if not isinstance(value, _Replicated):
            return value
        t = value.build()
        self._fields[key] = t
        rep = self.__dict__.get("_replica")
        if rep is not None:
            rep[1][key] = (t, t._version, value.src)      # `.to(cuda)` still ships the single graph while this copy is unmodified
        return t

    def __getattr__(self, key):
        f = object.__getattribute__(self, "_fields")
        if key in f:
            return self._materialise(key, f[key])
        raise AttributeError(key)

    def __getitem__(self, key):
        return self._materialise(key, self._fields[key])

    def items(self):
        return [(k, self._materialise(k, v)) for k, v in list(self._fields.items())]

    @property
    def num_nodes(self) -> int:
        own = self.__dict__.get("_lengths")
        return sum(own) if own is not None else super().num_nodes

    @classmethod
    def from_data_list(cls, graphs):
        lengths = [g.num_nodes for g in graphs]
        offsets = [0]
        for n in lengths:
            offsets.append(offsets[-1] + n)
        fields = {}
        # B references to ONE graph (sample.py:223 and finetune.py:325 build their batch exactly so): the concatenated tensors are
        # plain replications of the single-graph tensors, so they are not built here (see _Replicated) and `.to(cuda)` ships the
        # graph once (3.9 MB instead of 988 MB at L = 84, B = 256) and replicates on the device
        replicated = len(graphs) > 1 and all(g is graphs[0] for g in graphs)
        for k in graphs[0].keys():
            vals = [g[k] for g in graphs]
            if not torch.is_tensor(vals[0]):
                fields[k] = vals
            elif replicated:
                fields[k] = _Replicated(vals[0], len(graphs), k == "edge_index", lengths[0])
            elif k == "edge_index":
                fields[k] = torch.cat([v + o for v, o in zip(vals, offsets[:-1])], dim=1)
            else:
                fields[k] = torch.cat(vals, dim=0)
        fields["batch"] = torch.repeat_interleave(torch.arange(len(graphs)), torch.tensor(lengths))
        fields["ptr"] = torch.tensor(offsets, dtype=torch.long)
        out = cls(**fields)
        object.__setattr__(out, "_lengths", lengths)
        object.__setattr__(out, "_edges", [int(g["edge_index"].shape[1]) if "edge_index" in g else 0 for g in graphs])
        if replicated:      # field -> (built tensor, its version, source), filled in as fields are materialised on the host
            object.__setattr__(out, "_replica", (len(graphs), {}))
        return out

    def _replica_source(self, key, value):
        """The single-graph tensor `value` (field `key`) is B copies of, or None (field replaced / modified / not a replica)."""
        if isinstance(value, _Replicated):
            return value.src if value.src.device.type == "cpu" else None
        rep = self.__dict__.get("_replica")
        if rep is None or not torch.is_tensor(value):
            return None
        ent = rep[1].get(key)
        if ent is None or value is not ent[0] or value._version != ent[1] or value.device.type != "cpu":
            return None
        return ent[2]

    def h2d_nbytes(self) -> int:
        """Bytes `.to(cuda)` copies from the host for this batch."""
        n = 0
        for k, v in self._fields.items():
            if isinstance(v, _Replicated):
                n += v.src.numel() * v.src.element_size() if v.src.device.type == "cpu" else 0
            elif torch.is_tensor(v) and v.device.type == "cpu":
                src = self._replica_source(k, v)
                n += (src if src is not None else v).numel() * v.element_size()
        return n

    def to(self, device, non_blocking: bool = False):
        dev = torch.device(device)
        if dev.type != "cuda" or self.__dict__.get("_replica") is None:
            return self._clone_meta({k: (v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else
                                         (_Replicated(v.src.to(device, non_blocking=non_blocking), v.copies, v.is_edge_index, v.nodes) if isinstance(v, _Replicated) else v))
                                     for k, v in self._fields.items()})
        B = self.__dict__["_replica"][0]
        n, e = self.__dict__["_lengths"][0], self.__dict__["_edges"][0]
        out = {}
        for k, v in self._fields.items():
            src = self._replica_source(k, v)
            if src is None:
                out[k] = v.to(dev, non_blocking=non_blocking) if torch.is_tensor(v) else (v.build().to(dev) if isinstance(v, _Replicated) else v)
                continue
            one = src.to(dev, non_blocking=non_blocking)
            if k == "edge_index":
                out[k] = one.repeat(1, B) + (torch.arange(B, device=dev) * n).repeat_interleave(e)
            else:
                out[k] = one.repeat(B, *([1] * (one.dim() - 1)))
        return self._clone_meta(out)

    @property
    def num_graphs(self) -> int:
        return len(self.__dict__["_lengths"])

    @property
    def lengths(self):
        return list(self.__dict__["_lengths"])

    def to_data_list(self):
        lengths, edges = self.__dict__["_lengths"], self.__dict__["_edges"]
        n_total, e_total = sum(lengths), sum(edges)
        out, o, eo = [], 0, 0
        for g, (n, e) in enumerate(zip(lengths, edges)):
            f = {}
            for k, v in self._fields.items():
                if k in ("batch", "ptr"):
                    continue
                if isinstance(v, _Replicated):
                    f[k] = v.src
                elif not torch.is_tensor(v):
                    f[k] = v[g] if isinstance(v, list) and len(v) == len(lengths) else v
                elif k == "edge_index":
                    f[k] = v[:, eo:eo + e] - o
                elif v.shape[0] == n_total:
                    f[k] = v[o:o + n]
                elif e_total and v.shape[0] == e_total:
                    f[k] = v[eo:eo + e]
                else:
                    f[k] = v
            out.append(ChemGraph(**f))
            o, eo = o + n, eo + e
        return out


def complete_graph_edge_index(seq_len: int) -> torch.Tensor:
    """Row-major complete graph exactly as sample.py:165-171 builds it (bit-exact integer part)."""
    return torch.cat([
        torch.arange(seq_len).repeat_interleave(seq_len).view(1, seq_len**2),
        torch.arange(seq_len).repeat(seq_len).view(1, seq_len**2),
    ], dim=0)


def batch_lengths(batch) -> list[int]:
    """Per-graph residue counts of any PyG-like batch (one host sync unless `ptr` is on the host)."""
    own = getattr(batch, "__dict__", {}).get("_lengths")
    if own is not None:
        return list(own)
    ptr = batch["ptr"] if "ptr" in batch else None
    if ptr is not None:
        p = ptr.tolist()
        return [p[i + 1] - p[i] for i in range(len(p) - 1)]
    return torch.bincount(batch["batch"]).tolist()
