"""CPU tests of the host side of se3diff_b200: library export table, schedule scalars against the
reference-generated goldens, container semantics, state_dict surface.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden


def test_library_loads_and_exports_every_declared_symbol():
    from se3diff_b200 import _lib
    from se3diff_b200.build import build

    path = build()
    assert os.path.exists(path)
    header = open(os.path.join(ROOT, "include", "se3diff_b200.h")).read()
    declared = set(re.findall(r"\b(se3_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    h = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(h, name), f"{name} declared in include/se3diff_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    lib = _lib.lib()
    assert lib.se3_abi_version() == _lib.ABI_VERSION == 5


def test_packed_pair_operand_sizes_follow_the_documented_layout():
    """se3_ipa_tc_packed_pair_bytes (host-side, no GPU): the packed pair bias is [H][L][pitch] with pitch a whole number of 8-element
    chunks that covers L; for L <= 128 (query-major rows read with one 16-byte shared-memory load per logit step) the chunk count
    is odd where 8 * chunks <= 128 allows it -- eight consecutive rows then start in eight different bank groups -- and the slab fits
    the 256 * round_up(L, 16) bytes of the probability operand it lands in; the pair values are [L][H][round_up(L,16)/8][16][8]."""
    from se3diff_b200 import ops

    for L in list(range(1, 140)) + [200, 256, 257, 300, 511, 512]:
        for H in (1, 4, 32):
            bias_bytes, value_bytes = ops._packed_pair_sizes(L, H)
            assert bias_bytes % (2 * H * L) == 0
            pitch = bias_bytes // (2 * H * L)
            assert pitch == ops.ipa_tc_bias_pitch(L, H) and pitch % 8 == 0 and L <= pitch <= L + 15
            if L <= 128:
                chunks = pitch // 8
                assert chunks % 2 == 1 or pitch == 128, (L, pitch)
                assert L * pitch * 2 <= 256 * ((L + 15) // 16 * 16), (L, pitch)
                words = pitch // 2                                   # 4-byte words between consecutive rows
                if chunks % 2 == 1:                                  # eight consecutive rows: eight distinct groups of four banks
                    assert len({(words * r % 32) // 4 for r in range(8)}) == 8, (L, pitch)
            else:
                assert pitch == (L + 7) // 8 * 8
            assert value_bytes == L * H * ((L + 15) // 16 * 16) * 16 * 2


def test_struct_layouts_match_header():
    from se3diff_b200 import _lib

    header = open(os.path.join(ROOT, "include", "se3diff_b200.h")).read()
    for cname, cls in (("se3_em_scalars", _lib.EmScalars), ("se3_dpm_scalars", _lib.DpmScalars),
                       ("se3_heun_scalars", _lib.HeunScalars), ("se3_ipa_shape", _lib.IpaShape)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(float|int32_t)\s+", "", decl)
            names += [n.strip() for n in decl.split(",")]
        assert names == [f[0] for f in cls._fields_], cname


def test_ops_reject_cpu_tensors():
    from se3diff_b200 import _lib, ops

    with pytest.raises(_lib.Se3LibraryError):
        ops.so3_exp(torch.zeros(4, 3))


def test_schedule_scalars_match_reference_goldens():
    """schedule.py against tests/golden/schedules.npz (minted by the reference SDE objects)."""
    from se3diff_b200 import schedule
    from se3diff_b200.sdes import CosineVPSDE

    g = load_golden("schedules.npz")
    tab = load_golden("so3_tables.npz")
    cols = list(g["columns"])

    class So3Stub:  # schedule only needs _marginal_std, beta, tol and the two CPU tables
        tol = 1e-7
        sigma_min, sigma_max = 0.02, 2.33

        class score_function:
            sigma_grid = torch.from_numpy(tab["small_sigma_grid"])
            score_scaling = torch.from_numpy(tab["small_score_scaling"])

        def _marginal_std(self, t):
            return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

        def beta(self, t):
            return self._marginal_std(t) * np.sqrt(2.0 * np.log(self.sigma_max / self.sigma_min))

    r3, so3 = CosineVPSDE(0.008), So3Stub()
    d = g["dpm"]
    c = {n: i for i, n in enumerate(cols)}
    steps = schedule.dpm_schedule(r3, so3, 50, 0.99, 0.001)
    f32 = lambda x: float(np.float32(x))
    for i, st in enumerate(steps):
        s = st.scalars
        assert st.t == d[i, c["t"]] and st.t_lambda == d[i, c["t_lambda"]]
        assert s.pos_std_t == d[i, c["std_t"]] and s.pos_std_lam == d[i, c["std_lambda"]]
        assert s.rot_g_t == d[i, c["so3_g_t"]] and s.rot_g_lam == d[i, c["so3_g_lambda"]]
        assert s.rot_scale_t == d[i, c["score_scaling_t"]] and s.rot_scale_lam == d[i, c["score_scaling_lambda"]]
        assert s.dt == d[i, c["dt"]]
        assert s.pos_c_x_mid == f32(np.float32(d[i, c["alpha_lambda"]]) / np.float32(d[i, c["alpha_t"]]))
        assert s.pos_c_x_fin == f32(np.float32(d[i, c["alpha_next"]]) / np.float32(d[i, c["alpha_t"]]))
        assert s.dt_mid == f32(np.float32(d[i, c["t_lambda"]]) - np.float32(d[i, c["t"]]))
    e = g["em"]
    for i, st in enumerate(schedule.em_schedule(r3, so3, 200, 0.99, 0.001)):
        s = st.scalars
        assert st.t == e[i, c["t"]] and s.dt == e[i, c["dt"]] and s.pos_beta == e[i, c["beta_t"]]
        assert s.pos_std == e[i, c["std_t"]] and s.rot_g == e[i, c["so3_g_t"]] and s.rot_scale == e[i, c["score_scaling_t"]]
        assert s.score_weight == 1.0 and s.noise_weight == 1.0
    hs = schedule.heun_schedule(r3, so3, 100, 0.99, 0.001, 0.5)
    hgold = g["heun"]
    assert hs[0].t_hat == hs[0].t and hs[0].scalars.churn_dt == 0.0
    for i, st in enumerate(hs):
        assert st.t == hgold[i, c["t"]] and st.t_next == hgold[i, c["t_next"]]
        assert st.scalars.next_pos_std == hgold[i, c["std_next"]]
        if i > 0:
            assert st.t_hat > st.t and st.scalars.churn_dt > 0
        assert st.correct


def test_chemgraph_container_semantics():
    from se3diff_b200.chemgraph import Batch, ChemGraph, batch_lengths, complete_graph_edge_index

    gs = []
    for n in (3, 5):
        gs.append(ChemGraph(pos=torch.randn(n, 3), node_orientations=torch.eye(3).repeat(n, 1, 1),
                            edge_index=complete_graph_edge_index(n), single_embeds=torch.randn(n, 384),
                            pair_embeds=torch.randn(n * n, 128), system_id=f"g{n}"))
    b = Batch.from_data_list(gs)
    assert b.num_graphs == 2 and batch_lengths(b) == [3, 5]
    assert b.batch.tolist() == [0] * 3 + [1] * 5 and b.ptr.tolist() == [0, 3, 8]
    assert torch.equal(b.edge_index[:, 9:], complete_graph_edge_index(5) + 3)
    b2 = b.replace(pos=None)
    assert b2.pos is None and b.pos is not None and "pos" in b2 and b2.single_embeds is b.single_embeds
    b["pos"] = torch.zeros(8, 3)
    assert b.pos.abs().sum() == 0
    back = b.to_data_list()
    assert [x.num_nodes for x in back] == [3, 5] and torch.equal(back[1].edge_index, complete_graph_edge_index(5))
    assert torch.equal(back[1].pair_embeds, gs[1].pair_embeds) and back[0].system_id == "g3"
    # edge_index layout is the reference's (sample.py:165-171)
    assert complete_graph_edge_index(3).tolist() == [[0, 0, 0, 1, 1, 1, 2, 2, 2], [0, 1, 2, 0, 1, 2, 0, 1, 2]]


def test_state_dict_surface_matches_reference_golden():
    """Same keys and shapes as the reference checkpoint layout (tests/state_dict.ptkeep re-exported into
    score_model_tiny.npz) and bit-exact relative-position buckets."""
    import yaml

    from se3diff_b200.models import DiGConditionalScoreModel, RelativePositionBias

    g = load_golden("score_model_tiny.npz")
    cfg = yaml.safe_load(str(g["cfg_json"]))
    m = DiGConditionalScoreModel(**cfg)
    sd = {k[4:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd::")}
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd)
    full = DiGConditionalScoreModel()
    assert sum(p.numel() for p in full.parameters()) == 31_284_486
    assert sum(p.numel() for p in full.model_nn.st_module.encoder.layers[0].parameters()) == 3_813_408
    ft = DiGConditionalScoreModel(dim_model=64, dim_pair=32, num_layers=2, num_heads=4, dim_hidden=256)
    assert sum(p.numel() for p in ft.parameters()) == 193_806
    b = RelativePositionBias._relative_position_bucket(torch.arange(40), 64, 128).tolist()
    assert b == list(range(16)) + [16, 16, 16, 17, 17, 18, 18, 18, 19, 19, 19, 20, 20, 20, 20, 21, 21, 21, 21, 22, 22, 22, 22, 22]


def test_samplers_refuse_cpu():
    from se3diff_b200 import shortcuts
    from se3diff_b200.chemgraph import Batch, ChemGraph

    b = Batch.from_data_list([ChemGraph(pos=torch.zeros(2, 3), node_orientations=torch.zeros(2, 3, 3))])
    with pytest.raises(RuntimeError):
        shortcuts.dpm_solver(batch=b, sdes={"pos": shortcuts.CosineVPSDE(), "node_orientations": None},
                             score_model=lambda x, t: x, num_steps=2, max_t=0.99, min_t=0.001, device="cpu")


def test_batch_assembly_and_result_format(tmp_path):
    """sampling_io: the graph of sample.py:143-183, the npz naming / resume rule of utils.py:13-28 and the batch loop of
    sample.py:288-308, driven with a stub denoiser (the real ones need a GPU)."""
    import numpy as np

    from se3diff_b200 import sampling_io as sio

    L, seq = 7, "ACDEFGH"
    g = torch.Generator().manual_seed(0)
    single, pair = torch.randn(L, 384, generator=g), torch.randn(L, L, 128, generator=g)
    np.save(tmp_path / "single.npy", single.numpy())
    cg = sio.generate_chemgraph(sequence=seq, single_embeds=tmp_path / "single.npy", pair_embeds=pair.numpy())
    assert cg.pair_embeds.shape == (L * L, 128) and torch.equal(cg.pair_embeds[2 * L + 3], pair[2, 3])
    assert torch.equal(cg.edge_index[:, 2 * L + 3], torch.tensor([2, 3])) and torch.isnan(cg.pos).all()
    with pytest.raises(ValueError):
        sio.generate_chemgraph(sequence=seq + "A", single_embeds=single, pair_embeds=pair)
    assert sio.format_npz_samples_filename(30, 10) == "batch_0000030_0000040.npz"

    calls = []

    def stub_denoiser(*, batch, sdes, score_model, device=None):
        calls.append((batch.num_graphs, torch.initial_seed()))
        n = batch["pos"].shape[0]
        return batch.replace(pos=torch.randn(n, 3), node_orientations=torch.eye(3).expand(n, 3, 3).clone())

    out = tmp_path / "samples"
    kw = dict(sequence=seq, chemgraph=cg, output_dir=out, bundle=(None, None, stub_denoiser), batch_size=4)
    first = sio.sample_to_dir(num_samples=6, **kw)
    assert [p.name for p in first] == ["batch_0000000_0000004.npz", "batch_0000004_0000006.npz"]
    assert calls == [(4, 0), (2, 4)]                       # batch sizes and per-batch seeds = global sample offsets
    assert sio.count_samples_in_output_dir(out) == 6
    second = sio.sample_to_dir(num_samples=11, **kw)       # resume: only the missing 5
    assert [p.name for p in second] == ["batch_0000006_0000010.npz", "batch_0000010_0000011.npz"]
    z = np.load(second[0])
    assert set(z.files) == {"pos", "node_orientations", "sequence"} and z["pos"].shape == (4, L, 3)
    assert z["node_orientations"].shape == (4, L, 3, 3) and z["sequence"].item() == seq and z["pos"].dtype == np.float32
    pos, rot = sio.load_samples(out, seq)
    assert pos.shape == (11, L, 3) and rot.shape == (11, L, 3, 3)
    torch.manual_seed(4)
    assert torch.equal(pos[4:6], torch.randn(2 * L, 3).view(2, L, 3))   # batch seeded with its offset (sample.py:298-306)


def test_path_functionals_match_the_pinned_restatement():
    """se3diff_b200.pathwise (mirror of bioemu/ppft.py) against oracle.toy's functionals, which tests/golden/toy.npz pins bit
    for bit to the reference's own ppft through the toy fine-tune loss."""
    from oracle import toy
    from se3diff_b200 import pathwise as P

    g = torch.Generator().manual_seed(3)
    Tn, B = 7, 9
    us = torch.randn(Tn, B, 5, 3, generator=g).flatten(-2, -1).requires_grad_(True)
    dWs = torch.randn(Tn, B, 15, generator=g) * 0.1
    dts = -torch.rand(Tn, generator=g) * 0.02
    hs, h_stars = torch.rand(B, 2, generator=g), torch.tensor([0.4, 0.7])
    assert torch.allclose(P.riemannian_ito_integral(us, dWs), toy.ito_integral(us, dWs), rtol=1e-6, atol=1e-7)
    assert torch.allclose(P.riemannian_quadratic_covariation(us, us, dts), toy.quadratic_covariation(us, us, dts), rtol=1e-6, atol=1e-8)
    w = P.compute_int_dws(us=us, dWs=dWs)
    uu = P.compute_int_u_u_dt(us=us, dts=dts)
    sg = uu.detach() * 1.3
    a = P.compute_ev_loss(ws=w, hs=hs, h_stars=h_stars) + 0.1 * P.compute_kl_loss(ws=w, int_u_u_dt=uu, int_u_u_dt_sg=sg)
    b = toy.compute_ev_loss(toy.ito_integral(us, -dWs), hs, h_stars) + 0.1 * toy.compute_kl_loss(toy.ito_integral(us, -dWs), toy.quadratic_covariation(us, us, -dts), sg)
    assert torch.allclose(a, b, rtol=1e-5)
    ga, gb = torch.autograd.grad(a, us, retain_graph=True)[0], torch.autograd.grad(b, us)[0]
    assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-8)
    assert torch.allclose(P.compute_ws(us=us, dWs=dWs, dts=dts), torch.ones(B))          # exp(0) at the current parameters
    assert torch.allclose(P.rloo_baseline(uu), toy.rloo_baseline(uu))


def test_pdb_ca_parser_on_the_reference_structures():
    """load_reference_ca_coords against the CA counts SURVEY.md quotes (56 / 84); needs the reference tree, skipped elsewhere."""
    from se3diff_b200.finetune_step import load_reference_ca_coords

    root = "/root/reference/structures"
    if not os.path.isdir(root):
        pytest.skip("reference tree not mounted")
    sh3 = load_reference_ca_coords(os.path.join(root, "2vwf_trimmed_SH3.pdb"))
    pdz = load_reference_ca_coords(os.path.join(root, "1be9_trimmed.pdb"))
    assert sh3.shape == (56, 3) and pdz.shape == (84, 3)
    d = (sh3[1:] - sh3[:-1]).norm(dim=-1)
    assert 0.36 < d.min() and d.max() < 0.40        # consecutive C-alpha atoms are 0.38 nm apart


def test_batch_of_one_repeated_graph_tracks_replication():
    """sample.py:223 batches B references to one graph: the batch must know which fields are plain replications (so that the
    device transfer can ship the graph once), and must forget it for fields that were replaced or written to."""
    import torch
    from se3diff_b200.chemgraph import Batch, ChemGraph, complete_graph_edge_index

    L, B = 5, 3
    g = ChemGraph(pos=torch.randn(L, 3), node_orientations=torch.randn(L, 3, 3), edge_index=complete_graph_edge_index(L),
                  single_embeds=torch.randn(L, 4), pair_embeds=torch.randn(L * L, 2))
    full = lambda b: sum(v.numel() * v.element_size() for _, v in b.items() if torch.is_tensor(v))
    one = sum(v.numel() * v.element_size() for _, v in g.items())
    b = Batch.from_data_list([g] * B)
    small = b["batch"].numel() * 8 + b["ptr"].numel() * 8
    assert b.h2d_nbytes() == one + small < full(b)
    b2 = b.replace(pos=torch.zeros(B * L, 3))                      # a replaced field travels whole
    assert b2.h2d_nbytes() == b.h2d_nbytes() + (B - 1) * L * 3 * 4
    b["single_embeds"].add_(1.0)                                   # so does one that was written to in place
    assert b.h2d_nbytes() == one + small + (B - 1) * L * 4 * 4
    distinct = Batch.from_data_list([g, g.replace(pos=torch.randn(L, 3)), g])
    assert distinct.h2d_nbytes() == full(distinct)
    assert torch.equal(b.to("cpu")["pair_embeds"], b["pair_embeds"])


def test_ipa_backward_shape_gate_and_cpu_refusal():
    """ops.ipa_bwd_supported mirrors the shared-memory bound of se3_ipa_attention_bwd (include/se3diff_b200.h), and the
    differentiable operator has no CPU path."""
    from se3diff_b200 import _lib, ops

    ok = lambda B, n, H, dk: ops.ipa_bwd_supported(ops.ipa_shape(B, n, H, dk, 1, head_major=False))
    assert ok(1280, 84, 4, 16) and ok(64, 56, 4, 16) and ok(2, 128, 2, 8) and ok(2, 128, 32, 16)
    assert ok(2, 129, 4, 16) and ok(2, 512, 4, 16)    # tiled two-kernel edition (keys of a (sample, head) no longer fit one CTA)
    assert ok(2, 128, 4, 32)                          # 235 KB of shared memory for the resident edition: tiled edition
    assert ok(2, 96, 4, 32) and ok(2, 512, 4, 32)
    assert not ok(2, 513, 4, 16)                      # SE3_IPA_BWD_MAX_LEN
    assert not ok(0, 84, 4, 16) and not ok(70000, 84, 4, 16) and not ok(2, 84, 4, 12)
    sh = ops.ipa_shape(1, 8, 2, 4, 1, head_major=False)
    x = torch.zeros(8, sh.proj_stride)
    with pytest.raises(_lib.Se3LibraryError):
        ops.IpaAttention.apply(x, torch.zeros(8, 9), torch.zeros(8, 3), torch.zeros(1, 2, 8, 8), torch.zeros(1, 8, 8, 8), None,
                               torch.zeros(2), 0.5, sh)
    with pytest.raises(_lib.Se3LibraryError):
        ops.r3_update_dpm(torch.zeros(4, 3), torch.zeros(4, 3), _lib.DpmScalars(), final_half=False)


def test_bench_stdout_carries_only_the_result_line():
    """bench.claim_stdout(): whatever native code prints on file descriptor 1 during the run (NCCL's version banner under
    torchrun) goes to stderr; the one JSON line goes to the real stdout."""
    import subprocess
    import sys

    code = ("import os, sys, json; sys.path.insert(0, %r); import bench; out = bench.claim_stdout(); "
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n'); print('python-level chatter'); "
            "print(json.dumps({'metric': 'm', 'value': 1.0}), file=out, flush=True)") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert lines == ['{"metric": "m", "value": 1.0}'], r.stdout
    assert "NCCL version" in r.stderr and "python-level chatter" in r.stderr


def test_einsum_attention_formulation_equals_the_oracle_on_cpu():
    """`DistributionalGraphormer._ipa_torch` (the autograd formulation the backward kernel is tested against, fp64, on the
    GPU) evaluated on CPU tensors against `ScoreModelOracle._ipa`, the restatement pinned bit-exactly to the reference's
    SAAttention.forward (structure_module.py:109-220): values in fp32 and fp64, and fp64 gradients of a linear functional."""
    from oracle.score_model import ScoreModelOracle
    from se3diff_b200.models import DistributionalGraphormer, SAAttention

    torch.manual_seed(3)
    B, n, H, dk, dp = 2, 19, 4, 16, 32
    D = H * dk
    a = SAAttention(D, dp, H, dropout=0.0)
    x1d, x2d = torch.randn(B, n, D), torch.randn(B, n, n, dp)
    T = torch.randn(B, n, 3) * 2.0
    R = torch.linalg.qr(torch.randn(B, n, 3, 3))[0]
    R = R * torch.sign(torch.linalg.det(R))[..., None, None]
    bias = torch.zeros(B, 1, 1, n)
    bias[1, ..., n - 3:] = float("-inf")
    pre = "st_module.encoder.layers.0.attn."
    sd = {"model_nn." + pre + k: v for k, v in a.state_dict().items()}
    sd["model_nn.x1d_proj.1.weight"] = torch.zeros(D, 4)
    with torch.no_grad():
        want = ScoreModelOracle(sd, num_heads=H)._ipa(x1d, x2d, T, R, bias, pre)
        got = DistributionalGraphormer._ipa_torch(a, x1d, x2d, T, R, bias)
    assert (got - want).abs().max() <= 2e-6 * want.abs().max()
    # fp64: values and gradients (autograd through both formulations)
    a64 = SAAttention(D, dp, H, dropout=0.0).double()
    a64.load_state_dict({k: v.double() for k, v in a.state_dict().items()})
    orc = ScoreModelOracle(sd, num_heads=H)
    orc.p = {k: v.double().requires_grad_(True) for k, v in orc.p.items()}
    x1, x2 = x1d.double().requires_grad_(True), x2d.double().requires_grad_(True)
    w = torch.randn(B, n, D, dtype=torch.float64)
    (orc._ipa(x1, x2, T.double(), R.double(), bias.double(), pre) * w).sum().backward()
    g_want = (x1.grad.clone(), x2.grad.clone(), orc.p[pre + "trained_point_weight"].grad.clone(), orc.p[pre + "pair_value.weight"].grad.clone())
    x1.grad = x2.grad = None
    (DistributionalGraphormer._ipa_torch(a64, x1, x2, T.double(), R.double(), bias.double()) * w).sum().backward()
    g_got = (x1.grad, x2.grad, a64.trained_point_weight.grad, a64.pair_value.weight.grad)
    for u, v in zip(g_got, g_want):
        assert (u - v).abs().max() <= 1e-6 * v.abs().max()      # the reference aggregates the points in fp32 (:193-196) even in an fp64 run


def test_tc_operator_truth_function_equals_the_oracle_on_cpu():
    """`tests/ipa_tc_reference.ref` -- the fp64 einsum the tensor-core attention kernels are tested against on the GPU -- evaluated on
    CPU tensors against `ScoreModelOracle._ipa`, the restatement pinned bit-exactly to the reference's SAAttention.forward
    (structure_module.py:109-220), with per-sample pair tensors, a key mask and 10-nm translations."""
    import math

    import torch.nn.functional as F
    from ipa_tc_reference import ref
    from oracle.score_model import ScoreModelOracle
    from se3diff_b200.models import SAAttention

    torch.manual_seed(5)
    for B, n, H, dk, dp, shared in ((2, 19, 4, 16, 32, False), (3, 33, 8, 16, 24, True)):
        D = H * dk
        a = SAAttention(D, dp, H, dropout=0.0)
        x1d = torch.randn(B, n, D)
        x2d = torch.randn(1 if shared else B, n, n, dp)
        T = torch.randn(B, n, 3) * 10.0
        R = torch.linalg.qr(torch.randn(B, n, 3, 3))[0]
        R = R * torch.sign(torch.linalg.det(R))[..., None, None]
        key_bias = torch.zeros(B, n)
        key_bias[1, n - 3:] = float("-inf")
        pre = "st_module.encoder.layers.0.attn."
        sd = {"model_nn." + pre + k: v for k, v in a.state_dict().items()}
        sd["model_nn.x1d_proj.1.weight"] = torch.zeros(D, 4)
        with torch.no_grad():
            want = ScoreModelOracle(sd, num_heads=H)._ipa(x1d, x2d.expand(B, -1, -1, -1), T, R, key_bias[:, None, None, :], pre)
            proj = F.linear(x1d, a.fused_projection_weight()).reshape(B * n, -1)
            pair_bias = (a.pair_weight * a.pair_bias(x2d)).permute(0, 3, 1, 2)
            hw = -0.5 * a.point_weight * F.softplus(a.trained_point_weight)
            feat = ref(proj, R.reshape(B * n, 9), T.reshape(B * n, 3), pair_bias, a.pair_value(x2d), hw, B, n, heads=H, d_k=dk, key_bias=key_bias)
            got = F.linear(feat.float().view(B, n, -1), a.fc_out.weight, a.fc_out.bias)
        assert a.point_weight == 1.0 / math.sqrt(54)
        assert (got - want).abs().max() <= 3e-6 * want.abs().max(), (got - want).abs().max() / want.abs().max()


def test_split_perm_host_function_equals_its_restatement():
    """se3_ipa_split_perm (a host function of the C ABI): the row index sets that turn the reference's fused projection weight
    [q | k | v | q_pt | k_pt | v_pt] (structure_module.py:56-107, parameter order) into head-major records -- integer work, exact."""
    from se3diff_b200 import ops

    for heads, dk in ((32, 16), (4, 16), (3, 8)):
        hd, ar = heads * dk, torch.arange
        sc, pt, qpos = [], [], []
        for h in range(heads):
            qpos.append(h * 3 * dk + ar(dk))
            sc += [h * dk + ar(dk), hd + h * dk + ar(dk), 2 * hd + h * dk + ar(dk)]
            pt += [3 * hd + h * 12 + ar(12), 3 * hd + 12 * heads + h * 12 + ar(12), 3 * hd + 24 * heads + h * 24 + ar(24)]
        got = ops.ipa_split_perms(heads, dk)
        for g, w in zip(got, (torch.cat(sc), torch.cat(pt), torch.cat(qpos))):
            assert torch.equal(g, w)
        assert sorted(torch.cat(got[:2]).tolist()) == list(range(3 * hd + 48 * heads))       # a permutation of all rows


def _instantiate(node):
    """What hydra.utils.instantiate does with the two YAML idioms the reference uses (sample.py:120-138, finetune.py:150-188):
    `_target_` = dotted name of a callable, remaining keys = keyword arguments (nested nodes first), `_partial_: true` = return
    functools.partial instead of calling."""
    import functools
    import importlib

    if isinstance(node, dict) and "_target_" in node:
        kw = {k: _instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_")}
        mod, name = node["_target_"].rsplit(".", 1)
        fn = getattr(importlib.import_module(mod), name)
        return functools.partial(fn, **kw) if node.get("_partial_") else fn(**kw)
    if isinstance(node, dict):
        return {k: _instantiate(v) for k, v in node.items()}
    return node


def test_plugin_yaml_targets_resolve():
    """The drop-in mechanism itself: the YAMLs shipped under se3diff_b200/config/ are the reference's
    config/denoiser/*.yaml and checkpoints/bioemu-v1.0/config.yaml with `_target_` pointed at se3diff_b200.shortcuts.  Resolved
    the way sample.py:120-138 resolves them: every denoiser becomes a partial carrying the reference's own settings, the score
    models are constructed with the reference's parameter count, the SDE nodes bind to classes that accept the reference's keywords."""
    import inspect

    import yaml
    from se3diff_b200 import shortcuts

    cfg_dir = os.path.join(ROOT, "se3diff_b200", "config")
    want = {"dpm": ("dpm_solver", 50, None), "euler_maruyama": ("euler_maruyama_predictor", 200, None),
            "euler_maruyama_finetune": ("euler_maruyama_predictor_finetune", 200, None), "heun": ("heun_denoiser", 100, 0.5),
            "heun_finetune": ("heun_denoiser_finetune", 100, 0.5), "sde_dpm_finetune": ("sde_dpm_solver_finetune", 50, None)}
    for name, (fn, steps, noise) in want.items():
        node = yaml.safe_load(open(os.path.join(cfg_dir, "denoiser", name + ".yaml")))
        d = _instantiate(node)
        assert d.func is getattr(shortcuts, fn)
        assert d.keywords["num_steps"] == steps and d.keywords["max_t"] == 0.99 and d.keywords["min_t"] == 0.001
        assert d.keywords.get("noise") == noise
        # the call site binds exactly these four more (sample.py:227-232; finetune.py adds finetune_model)
        params = inspect.signature(d.func).parameters
        assert {"batch", "sdes", "score_model", "device"} <= set(params) and all(p.kind is p.KEYWORD_ONLY for p in params.values())
    node = yaml.safe_load(open(os.path.join(cfg_dir, "bioemu-v1.0", "config.yaml")))
    model = _instantiate(node["score_model"])
    assert isinstance(model, shortcuts.DiGConditionalScoreModel)
    assert sum(p.numel() for p in model.parameters()) == 31_284_486
    control = _instantiate(node["finetune_model"])
    assert sum(p.numel() for p in control.parameters()) == 193_806
    pos_sde = _instantiate(node["sdes"]["pos"])
    assert isinstance(pos_sde, shortcuts.CosineVPSDE) and pos_sde.s == 0.008
    so3 = node["sdes"]["node_orientations"]
    assert list(node["sdes"]) == ["node_orientations", "pos"]                             # key order fixes the RNG order (denoiser.py:233)
    assert so3["_target_"] == "se3diff_b200.shortcuts.DiGSO3SDE"
    inspect.signature(shortcuts.DiGSO3SDE.__init__).bind(None, **{k: v for k, v in so3.items() if k != "_target_"})
    # shadowing bioemu.shortcuts exposes every name the reference's alias module does
    for name in ("CosineVPSDE", "DiGConditionalScoreModel", "DiGSO3SDE", "dpm_solver", "heun_denoiser", "euler_maruyama_predictor",
                 "euler_maruyama_predictor_finetune", "heun_denoiser_finetune", "sde_dpm_solver_finetune"):
        assert hasattr(shortcuts, name), name


def test_so3_table_cache_written_by_the_reference_is_read_back():
    """npz cache interop (so3_sde.py:914-990; file names :1098, 1354, 1607).  tests/golden/so3_cache/ holds the three files the
    UNMODIFIED reference's `DiGSO3SDE.__init__` wrote through `SO3LookupCache.save_cache` for a tiny table set
    (oracle/gen_golden.py::so3_cache).  This package must find them under the same names, load them without a GPU (no table
    build), and hold exactly their contents in buffers of the reference's names."""
    import numpy as np
    from oracle.gen_golden import CACHE_SDE
    from se3diff_b200 import sdes as S

    d = os.path.join(ROOT, "tests", "golden", "so3_cache")
    files = sorted(os.listdir(d))
    assert files == ["cache_igso3_s0.020-2.330-8_l64_o64-3.npz", "cache_score-scaling_s0.020-2.330-8_l65_o64-3.npz",
                     "cache_uso3_s0.020-2.330-8_o64-3.npz"]
    sde = S.DiGSO3SDE(**CACHE_SDE, cache_dir=d, overwrite_cache=False)         # would raise on this CPU-only box if it had to build
    ig, us, sc = (np.load(os.path.join(d, f)) for f in files[:1] + files[2:] + files[1:2])
    assert sde.igso3._get_cache_name() == files[0] and sde.uso3._get_cache_name() == files[2]
    assert torch.equal(sde.igso3.cdf_igso3, torch.from_numpy(ig["cdf_igso3"])) and torch.equal(sde.igso3.omega_grid, torch.from_numpy(ig["omega_grid"]))
    assert torch.equal(sde.uso3.cdf_igso3, torch.from_numpy(us["cdf_igso3"])) and torch.equal(sde.uso3.omega_grid, torch.from_numpy(us["omega_grid"]))
    assert torch.equal(sde.score_function.score_scaling, torch.from_numpy(sc["score_scaling"]))
    assert sde.igso3.cdf_igso3.dtype == torch.float32 and tuple(sde.igso3.cdf_igso3.shape) == (8, 64) and tuple(sde.uso3.cdf_igso3.shape) == (1, 64)
    assert sorted(os.listdir(d)) == files                                         # nothing rewritten
