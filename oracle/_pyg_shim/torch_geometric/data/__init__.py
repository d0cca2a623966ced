"""TEST INFRASTRUCTURE ONLY -- see ../__init__.py."""
from __future__ import annotations

import copy

import torch


class _Store(dict):
    """Attribute store.  PyG keeps a ``_parent`` back-reference that chemgraph.py:29 rewrites."""

    _parent = None

    def __copy__(self):
        out = _Store(self)
        out._parent = self._parent
        return out


class Data:
    def __init__(self, **kwargs):
        self.__dict__["_store"] = _Store()
        self._store._parent = self
        for k, v in kwargs.items():
            self._store[k] = v

    # attribute / item access -------------------------------------------------------------
    def __getattr__(self, key):
        if key.startswith("__"):
            raise AttributeError(key)
        store = self.__dict__.get("_store")
        if store is not None and key in store:
            return store[key]
        raise AttributeError(key)

    def __setattr__(self, key, value):
        if key in ("_store",):
            self.__dict__[key] = value
        else:
            self._store[key] = value

    def __getitem__(self, key):
        return self._store[key]

    def __setitem__(self, key, value):
        self._store[key] = value

    def __contains__(self, key):
        return key in self._store

    def keys(self):
        return list(self._store.keys())

    def items(self):
        return list(self._store.items())

    def to(self, device, *args, **kwargs):
        out = copy.copy(self)
        out.__dict__["_store"] = copy.copy(self._store)
        out._store._parent = out
        for k, v in list(out._store.items()):
            if torch.is_tensor(v):
                out._store[k] = v.to(device, *args, **kwargs)
        return out

    @property
    def num_nodes(self):
        for k in ("pos", "x", "node_orientations", "single_embeds"):
            if k in self._store and torch.is_tensor(self._store[k]):
                return self._store[k].shape[0]
        for v in self._store.values():
            if torch.is_tensor(v) and v.dim() > 0:
                return v.shape[0]
        return 0


_NODE_LEVEL_SKIP = {"edge_index", "batch", "ptr"}


class Batch(Data):
    _cls_cache: dict = {}

    @classmethod
    def from_data_list(cls, data_list):
        elem_cls = type(data_list[0])
        # PyG builds a dynamic subclass so that the batch is also an instance of the element class.
        if elem_cls is Data or elem_cls is Batch:
            dyn = Batch
        else:
            dyn = Batch._cls_cache.get(elem_cls)
            if dyn is None:
                dyn = type(f"{elem_cls.__name__}Batch", (elem_cls, Batch), {})
                Batch._cls_cache[elem_cls] = dyn
        out = dyn.__new__(dyn)
        out.__dict__["_store"] = _Store()
        out._store._parent = out

        num_nodes = [d.num_nodes for d in data_list]
        offsets = [0]
        for n in num_nodes:
            offsets.append(offsets[-1] + n)
        keys = data_list[0].keys()
        for k in keys:
            vals = [d[k] for d in data_list]
            if not torch.is_tensor(vals[0]):
                out._store[k] = vals
            elif k == "edge_index":
                out._store[k] = torch.cat([v + off for v, off in zip(vals, offsets[:-1])], dim=1)
            else:
                out._store[k] = torch.cat(vals, dim=0)
        out._store["batch"] = torch.cat(
            [torch.full((n,), i, dtype=torch.long) for i, n in enumerate(num_nodes)]
        )
        out._store["ptr"] = torch.tensor(offsets, dtype=torch.long)
        out.__dict__["_num_graphs"] = len(data_list)
        out.__dict__["_elem_cls"] = elem_cls
        out.__dict__["_num_nodes_list"] = num_nodes
        out.__dict__["_num_edges_list"] = [
            d["edge_index"].shape[1] if "edge_index" in d else 0 for d in data_list
        ]
        return out

    @property
    def num_graphs(self):
        return self.__dict__["_num_graphs"]

    def to(self, device, *args, **kwargs):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v
        out.__dict__["_store"] = copy.copy(self._store)
        out._store._parent = out
        for k, v in list(out._store.items()):
            if torch.is_tensor(v):
                out._store[k] = v.to(device, *args, **kwargs)
        return out

    def to_data_list(self):
        elem_cls = self.__dict__["_elem_cls"]
        ptr = self._store["ptr"].tolist()
        n_nodes_total = ptr[-1]
        e_off = [0]
        for n in self.__dict__["_num_edges_list"]:
            e_off.append(e_off[-1] + n)
        outs = []
        for g in range(self.num_graphs):
            kw = {}
            for k, v in self._store.items():
                if k in ("batch", "ptr"):
                    continue
                if not torch.is_tensor(v):
                    kw[k] = v[g] if isinstance(v, list) and len(v) == self.num_graphs else v
                elif k == "edge_index":
                    kw[k] = v[:, e_off[g] : e_off[g + 1]] - ptr[g]
                elif v.shape[0] == n_nodes_total:
                    kw[k] = v[ptr[g] : ptr[g + 1]]
                elif e_off[-1] > 0 and v.shape[0] == e_off[-1]:
                    kw[k] = v[e_off[g] : e_off[g + 1]]
                else:
                    kw[k] = v
            outs.append(elem_cls(**kw))
        return outs
