"""CPU ORACLE (test infrastructure) -- DiG score network forward, functional, from a state_dict.

Restates bioemu/src/bioemu/models.py (embeddings, sparse->dense, masks) and
bioemu/src/bioemu/structure_module.py (DiG-flavoured invariant point attention, FFN, DiffHead) in
eval mode (dropout off).  Faithful to the reference's cost profile too: the pair representation,
``pair_bias`` and ``pair_value`` are recomputed per sample, per layer, per call, exactly as
structure_module.py:179,209 does -- this is what `bench.py`'s cpu_baseline times.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def relative_position_bucket(rel: torch.Tensor, num_buckets: int, max_distance: int) -> torch.Tensor:
    """T5-style signed log buckets (models.py:94-125); integer-exact restatement."""
    nb = num_buckets // 2
    ret = (rel < 0).to(rel) * nb
    rel = torch.abs(rel)
    max_exact = nb // 2
    large = max_exact + (
        torch.log(rel / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)
    ).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return ret + torch.where(rel < max_exact, rel, large)


def sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    """models.py:50-69 with min_input 0, max_input 1000 (identity rescale)."""
    half = dim // 2
    t = (t - 0.0) * 1000.0 / (1000.0 - 0.0)
    f = torch.exp(torch.arange(half) * (-math.log(10000) / (half - 1)))
    e = t[:, None] * f[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def to_dense(x, lengths):
    """to_dense_batch (PyG 2.6.1 semantics: zero fill, mask)."""
    b, lmax = len(lengths), max(lengths)
    out = x.new_zeros((b, lmax) + tuple(x.shape[1:]))
    mask = torch.zeros(b, lmax, dtype=torch.bool)
    o = 0
    for g, n in enumerate(lengths):
        out[g, :n] = x[o : o + n]
        mask[g, :n] = True
        o += n
    return out, mask


class ScoreModelOracle:
    """DiGConditionalScoreModel.forward (models.py:359-384) on flat [N,.] inputs.

    single_embeds [N, 384]; pair_embeds: list over graphs of [L_g, L_g, 128] (the complete-graph edge
    list of sample.py:165-171 is row-major (i, j), so to_dense_adj == reshape); lengths list[int].
    """

    def __init__(self, state_dict: dict, num_heads: int, num_buckets: int = 64, max_distance: int = 128, dtype=torch.float32):
        # dtype: float32 restates the reference bit for bit; float64 evaluates the SAME expressions in double precision
        # (inputs are widened, outputs returned in float32) -- used by the tests to measure the fp32 noise floor of the
        # reference itself, i.e. how far two correct fp32 evaluations may differ.
        self.dtype = dtype
        self.p = {k.removeprefix("model_nn."): v.detach().to(dtype) for k, v in state_dict.items()}
        self.h = num_heads
        self.num_buckets, self.max_distance = num_buckets, max_distance
        self.n_layer = 1 + max(int(k.split(".")[3]) for k in self.p if k.startswith("st_module.encoder.layers."))
        self.d_model = self.p["x1d_proj.1.weight"].shape[0]
        self.context = None

    def set_context(self, single_embeds, pair_embeds, lengths, pos_is_known=None):
        self.single, self.lengths = single_embeds, list(lengths)
        lmax = max(lengths)
        pr = torch.zeros(len(lengths), lmax, lmax, pair_embeds[0].shape[-1], dtype=self.dtype)
        for g, n in enumerate(lengths):
            pr[g, :n, :n] = pair_embeds[g]
        self.pair = pr
        self.pos_is_known = pos_is_known
        return self

    # -- blocks -------------------------------------------------------------------------------
    def _ln(self, x, pre):
        return F.layer_norm(x, (x.shape[-1],), self.p[pre + ".weight"], self.p[pre + ".bias"])

    def _ipa(self, x1d, x2d, T, R, bias, pre):
        """SAAttention.forward (structure_module.py:109-220); R here is the rotation (not inverse)."""
        p, H = self.p, self.h
        lead = x1d.shape[:-1]
        q = F.linear(x1d, p[pre + "scalar_query.weight"]).reshape(*lead, H, -1)
        k = F.linear(x1d, p[pre + "scalar_key.weight"]).reshape(*lead, H, -1)
        v = F.linear(x1d, p[pre + "scalar_value.weight"]).reshape(*lead, H, -1)
        dk = q.shape[-1]
        sw = 1.0 / math.sqrt(3 * dk)
        pw = 1.0 / math.sqrt(3 * 4 * 9 / 2)
        s_attn = torch.einsum("bihc,bjhc->bhij", q * sw, k)

        def glob(x):  # apply_affine, structure_module.py:145-160
            return torch.matmul(R[:, :, None, None], x.unsqueeze(-1)).squeeze(-1) + T[:, :, None, None]

        qp = glob(F.linear(x1d, p[pre + "point_query.weight"]).reshape(*lead, H, -1, 3))
        kp = glob(F.linear(x1d, p[pre + "point_key.weight"]).reshape(*lead, H, -1, 3))
        vp = glob(F.linear(x1d, p[pre + "point_value.weight"]).reshape(*lead, H, -1, 3))
        d = torch.norm(qp.unsqueeze(2) - kp.unsqueeze(1), dim=-1)  # un-squared norm, :170
        w = pw * F.softplus(p[pre + "trained_point_weight"])
        p_attn = -0.5 * w[:, None, None] * torch.sum(d, dim=-1).permute(0, 3, 1, 2)
        pair_attn = (1.0 / math.sqrt(3)) * F.linear(x2d, p[pre + "pair_bias.weight"]).permute(0, 3, 1, 2)
        attn = torch.softmax(s_attn + p_attn + pair_attn + bias, dim=-1)

        o_s = torch.einsum("bhij,bjhc->bihc", attn, v).reshape(*lead, -1)
        o_pg = torch.einsum("bhij,bjhcp->bihcp", attn.to(self.dtype), vp.to(self.dtype))   # explicit fp32 in the reference (:193-196)
        o_pl = torch.matmul(R.transpose(-1, -2)[:, :, None, None],
                            (o_pg - T[:, :, None, None]).unsqueeze(-1)).squeeze(-1)
        o_n = torch.norm(o_pl, dim=-1).reshape(*lead, -1)
        o_pl = o_pl.reshape(*lead, -1)
        v_pair = F.linear(x2d, p[pre + "pair_value.weight"]).reshape(*x2d.shape[:-1], H, -1)
        o_pair = torch.einsum("bhij,bijhc->bihc", attn, v_pair).reshape(*lead, -1)
        feat = torch.cat([o_s, o_pl, o_pair, o_n], dim=-1)
        return F.linear(feat, p[pre + "fc_out.weight"], p[pre + "fc_out.bias"])

    def __call__(self, pos, rot, t):
        """-> (pos_out [N,3], rot_out [N,3]); t [B] in [0,1] (scaled by 1000 as models.py:365)."""
        p, lengths = self.p, self.lengths
        dt = self.dtype
        T, mask = to_dense(pos.to(dt), lengths)
        R, _ = to_dense(rot.to(dt), lengths)  # rotation; the wrapper transposes and IPA transposes back
        single, _ = to_dense(self.single, lengths)
        if self.pos_is_known is not None:
            known = to_dense(self.pos_is_known, lengths)[0].bool()
            attn_mask = ~(mask & known)
        else:
            attn_mask = ~mask
        te = t.to(dt) * 1000
        x1d = F.linear(self._ln(single.to(dt), "x1d_proj.0"), p["x1d_proj.1.weight"]) + sinusoid(te, self.d_model).to(dt)[:, None]
        x2d = F.linear(self._ln(self.pair.to(dt), "x2d_proj.0"), p["x2d_proj.1.weight"])
        seq = torch.arange(T.shape[1])
        rel = seq.unsqueeze(1) - seq.unsqueeze(0)
        bucket = relative_position_bucket(rel, self.num_buckets, self.max_distance)
        x2d = x2d + F.embedding(bucket, p["rp_proj.relative_attention_bias.weight"])[None]
        z = (~attn_mask).long().sum(-1, keepdims=True)
        filled = attn_mask.masked_fill(z == 0, False)
        bias = filled.to(dt).masked_fill(filled, float("-inf"))[:, None, :, None].permute(0, 3, 1, 2)
        for n in range(self.n_layer):
            pre = f"st_module.encoder.layers.{n}."
            x1d = x1d + self._ipa(self._ln(x1d, pre + "norm1"), x2d, T, R, bias, pre + "attn.")
            y = self._ln(x1d, pre + "norm2")
            y = F.linear(F.gelu(F.linear(y, p[pre + "ffn.ff.0.weight"], p[pre + "ffn.ff.0.bias"])),
                         p[pre + "ffn.ff.3.weight"], p[pre + "ffn.ff.3.bias"])
            x1d = x1d + y

        def head(name):
            pre = f"st_module.diff_head.{name}."
            y = self._ln(x1d, pre + "0")
            y = F.relu(F.linear(y, p[pre + "1.weight"], p[pre + "1.bias"]))
            return F.linear(y, p[pre + "3.weight"], p[pre + "3.bias"])

        T_eps, IR_eps = head("fc_t"), head("fc_eps")
        # models.py:305 -- IR_perturbed^T = R
        T_out = torch.matmul(R, T_eps.unsqueeze(-1)).squeeze(-1)
        return T_out[mask].float(), IR_eps[mask].float()


def make_edge_index(seq_len: int) -> torch.Tensor:
    """Row-major complete graph (sample.py:165-171)."""
    return torch.cat([
        torch.arange(seq_len).repeat_interleave(seq_len).view(1, seq_len**2),
        torch.arange(seq_len).repeat(seq_len).view(1, seq_len**2),
    ], dim=0)


def init_state_dict(dim_model=512, dim_pair=256, num_layers=8, num_heads=32, dim_hidden=1024,
                    num_buckets=64, seed=0) -> dict:
    """Random-init weights with the reference's parameter names and shapes,
    drawn with torch default initialisers in a fixed order under ``seed``.  Synthetic-weights
    generator for benchmarks: NOT the reference's RNG order (that needs the reference class)."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i, bias=True, pre=""):
        bound = 1.0 / math.sqrt(i)
        d = {pre + "weight": (torch.rand(o, i, generator=g) * 2 - 1) * bound}
        if bias:
            d[pre + "bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
        return d

    def ln(n, pre):
        return {pre + "weight": torch.ones(n), pre + "bias": torch.zeros(n)}

    sd = {}
    sd.update(ln(384, "x1d_proj.0."))
    sd.update(lin(dim_model, 384, False, "x1d_proj.1."))
    sd.update(ln(128, "x2d_proj.0."))
    sd.update(lin(dim_pair, 128, False, "x2d_proj.1."))
    sd["rp_proj.relative_attention_bias.weight"] = torch.randn(num_buckets, dim_pair, generator=g)
    sd["step_emb.dummy"] = torch.empty(0)
    H = num_heads
    for n in range(num_layers):
        pre = f"st_module.encoder.layers.{n}."
        sd.update(ln(dim_model, pre + "norm1."))
        sd.update(ln(dim_model, pre + "norm2."))
        a = pre + "attn."
        sd.update(lin(dim_model, dim_model, False, a + "scalar_query."))
        sd.update(lin(dim_model, dim_model, False, a + "scalar_key."))
        sd.update(lin(dim_model, dim_model, False, a + "scalar_value."))
        sd.update(lin(H, dim_pair, False, a + "pair_bias."))
        sd.update(lin(H * 12, dim_model, False, a + "point_query."))
        sd.update(lin(H * 12, dim_model, False, a + "point_key."))
        sd.update(lin(H * 24, dim_model, False, a + "point_value."))
        sd[a + "trained_point_weight"] = torch.rand(H, generator=g)
        sd.update(lin(dim_model, dim_pair, False, a + "pair_value."))
        sd.update(lin(dim_model, dim_model * 2 + H * 32, True, a + "fc_out."))
        sd.update(lin(dim_hidden, dim_model, True, pre + "ffn.ff.0."))
        sd.update(lin(dim_model, dim_hidden, True, pre + "ffn.ff.3."))
    for name in ("fc_t", "fc_eps"):
        pre = f"st_module.diff_head.{name}."
        sd.update(ln(dim_model, pre + "0."))
        sd.update(lin(dim_model, dim_model, True, pre + "1."))
        sd.update(lin(3, dim_model, True, pre + "3."))
    return {"model_nn." + k: v for k, v in sd.items()}
