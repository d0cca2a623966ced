"""TEST INFRASTRUCTURE ONLY -- see ../__init__.py."""
import torch


def _counts(batch, batch_size=None):
    b = int(batch.max().item()) + 1 if batch_size is None else batch_size
    return torch.bincount(batch, minlength=b), b


def to_dense_batch(x, batch=None, fill_value=0.0, max_num_nodes=None, batch_size=None):
    if batch is None:
        mask = torch.ones(1, x.shape[0], dtype=torch.bool, device=x.device)
        return x.unsqueeze(0), mask
    counts, b = _counts(batch, batch_size)
    lmax = int(counts.max().item()) if max_num_nodes is None else max_num_nodes
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    pos = torch.arange(x.shape[0], device=x.device) - ptr[batch]
    out = x.new_full((b * lmax,) + tuple(x.shape[1:]), fill_value)
    idx = batch * lmax + pos
    out[idx] = x
    mask = torch.zeros(b * lmax, dtype=torch.bool, device=x.device)
    mask[idx] = True
    return out.view((b, lmax) + tuple(x.shape[1:])), mask.view(b, lmax)


def to_dense_adj(edge_index, batch=None, edge_attr=None, max_num_nodes=None, batch_size=None):
    if batch is None:
        n = int(edge_index.max().item()) + 1
        batch = edge_index.new_zeros(n)
    counts, b = _counts(batch, batch_size)
    lmax = int(counts.max().item()) if max_num_nodes is None else max_num_nodes
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    g = batch[edge_index[0]]
    i = edge_index[0] - ptr[g]
    j = edge_index[1] - ptr[g]
    flat = g * lmax * lmax + i * lmax + j
    if edge_attr is None:
        edge_attr = torch.ones(edge_index.shape[1], device=edge_index.device)
    out = edge_attr.new_zeros((b * lmax * lmax,) + tuple(edge_attr.shape[1:]))
    out.index_add_(0, flat, edge_attr)  # scatter-add semantics, as PyG
    return out.view((b, lmax, lmax) + tuple(edge_attr.shape[1:]))
