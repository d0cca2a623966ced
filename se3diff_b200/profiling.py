"""Live per-kernel timing for bench.py's `roofline` objects: CUDA events recorded on the launching
stream around individual launches of this library's kernels (never under a profiler)."""
from __future__ import annotations

import json
import os

import torch

from . import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measured_peaks():
    """MEASURED_PEAKS.json (driver-written) or the fallback stated in B200_PROFILING.md."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), tensor_burst=float(d["bf16_tflops"]),
                    tensor_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


class _Hook:
    def __init__(self, name):
        self.name, self.orig, self.events = name, getattr(ops, name), []

    def __enter__(self):
        def wrapped(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = self.orig(*a, **k)
            e1.record()
            self.events.append((e0, e1))
            return r

        setattr(ops, self.name, wrapped)
        return self

    def __exit__(self, *exc):
        setattr(ops, self.name, self.orig)
        return False

    def mean_ms(self):
        torch.cuda.synchronize()
        ts = [a.elapsed_time(b) for a, b in self.events]
        return sum(ts) / max(1, len(ts)), len(ts)


def _time_alone(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


HBM_NOMINAL_GBS = 8000.0      # north_star's "~8 TB/s" (DGX B200 figure; HGX: 7.7 TB/s) -- reported next to the measured copy bandwidth
SMS, XU_LANES_PER_CLK_SM = 148, 16      # MUFU: 4 lanes / clock / SM sub-partition (measured: 8 cycles per warp instruction, scripts/microbench)
FP32_LANES_PER_CLK_SM, FP64_LANES_PER_CLK_SM = 128, 64     # B200: 128 FFMA and 64 DFMA lanes per SM and clock (~37 TFLOP/s FP64 chip-wide)


def _hbm_line(name, n, bytes_per, ms, pk, **extra):
    gbs = bytes_per * n / ms / 1e6
    return {"kernel": name, "bound": "hbm", "n": n, "bytes_per_unit": bytes_per, "ms": ms, "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s",
            "frac": gbs / pk["hbm"], "peak_source": pk["source"], "peak_nominal": HBM_NOMINAL_GBS, "frac_nominal": gbs / HBM_NOMINAL_GBS,
            "traffic": None, **extra}


def series_rooflines(device="cuda", sm_mhz=None):
    """K1a / K1b (so3_sde.py:1731-1940, 1131-1187, 1637-1696): the truncated IGSO(3) series and the table builds.  They are
    arithmetic-bound, not HBM-bound (one element = up to 2001 terms from 8 bytes of input): reported as series TERMS per second
    against the instruction rate of the pipe that evaluates a term.  A term of the (f, df) pair is one exp and three sin in the
    working precision -- libm-accurate, because the tables and scores must round like torch's (the fast MUFU forms are 2 ulp
    off) -- plus ~12 multiply-adds: ~100 FP32-pipe instructions in fp32, ~200 FP64-pipe instructions in fp64; a table row shares
    its exponential factors across the omega grid (~50 per term for the density, ~150 with the derivative).  The term count is
    the number actually evaluated: terms whose exponential underflows to exactly zero are skipped (the sum is unchanged bit
    for bit), so it depends on sigma."""
    import math

    clk = (sm_mhz or 1965) * 1e6
    g = torch.Generator(device=device).manual_seed(1)
    n, l_max = 1_000_000, 2000
    om = torch.rand(n, generator=g, device=device) * math.pi
    sg = 0.02 * (2.33 / 0.02) ** torch.rand(n, generator=g, device=device)

    def terms(sigma, tiny):      # l runs while exp(-l(l+1) sigma^2 / 2) > 0 in the working precision
        lcut = torch.sqrt(2.0 * (-math.log(tiny)) / sigma.double() ** 2)
        return torch.clamp(lcut.ceil(), max=l_max + 1)

    def line(name, n_terms, ms, lanes, per_term, what, **extra):
        peak = SMS * lanes * clk / per_term
        return {"kernel": name, "bound": what, "terms_evaluated": n_terms, "ms": ms, "achieved": n_terms / ms / 1e6, "peak": peak / 1e9,
                "unit": "Gterm/s", "frac": n_terms / ms / 1e6 / (peak / 1e9),
                "peak_source": f"148 SMs x {lanes} {what} lanes x {clk / 1e6:.0f} MHz / ~{per_term} instructions per term (estimate, see profiling.py)", **extra}

    t32 = terms(sg, 1.4e-45).sum().item()
    res = [line("se3_igso3_series_f32 (f, df, dlog)", t32, _time_alone(lambda: ops.igso3_series(om, sg, l_max, want=("f", "df", "dlog"))),
                FP32_LANES_PER_CLK_SM, 100, "fp32", n=n, l_max=l_max)]
    rv = torch.randn(n, 3, generator=g, device=device)
    res.append(line("se3_igso3_score", t32, _time_alone(lambda: ops.igso3_score(rv, sg, l_max)), FP32_LANES_PER_CLK_SM, 100, "fp32", n=n, l_max=l_max))
    n64 = 200_000
    om64, sg64 = om[:n64].double(), sg[:n64].double()
    t64 = terms(sg64, 4.9e-324).sum().item()
    res.append(line("se3_igso3_series_f64 (f, df, dlog)", t64, _time_alone(lambda: ops.igso3_series(om64, sg64, l_max, want=("f", "df", "dlog")), iters=5),
                    FP64_LANES_PER_CLK_SM, 200, "fp64", n=n64, l_max=l_max))
    # table builds of the shipped configuration (config.yaml:23-35): 1000 sigma rows x 2001 omega points x <= 2001 terms, fp64
    sig_grid = 0.02 * (2.33 / 0.02) ** torch.linspace(0.001, 1.0, 1000, device=device)
    om_pts = (torch.linspace(0.0, 1, 2001, device=device, dtype=torch.float64) ** 3 * math.pi)
    tt = (terms(sig_grid, 4.9e-324) * 2001).sum().item()
    res.append(line("se3_igso3_build_cdf (1000 x 2000 table)", tt, _time_alone(lambda: ops.igso3_build_cdf(sig_grid, om_pts, 2000), iters=3, warmup=1),
                    FP64_LANES_PER_CLK_SM, 50, "fp64", note="once per process (or read from the npz cache)"))
    res.append(line("se3_igso3_build_score_scaling (1000 rows)", tt, _time_alone(lambda: ops.igso3_build_score_scaling(sig_grid, om_pts, 2000), iters=3, warmup=1),
                    FP64_LANES_PER_CLK_SM, 150, "fp64", note="density and derivative series per (sigma, omega)"))
    return res


def elementwise_rooflines(n: int = 10_000_000, device="cuda"):
    """BASELINE config 3: SO(3) exp/log/compose, IGSO3 noising and the fused frame updates at n = 1e7
    (inputs 120-1200 MB, i.e. larger than L2).  Algorithmic bytes per unit from SURVEY.md 8(d)."""
    from . import _lib as L

    pk = measured_peaks()
    g = torch.Generator(device=device).manual_seed(0)
    v = torch.randn(n, 3, generator=g, device=device)
    r = ops.so3_exp(v)
    w = torch.randn(n, 3, generator=g, device=device) * 0.1
    z1, z2 = torch.randn(n, 3, generator=g, device=device), torch.randn(n, 3, generator=g, device=device)
    pos = torch.randn(n, 3, generator=g, device=device)
    out_r, out_p = torch.empty_like(r), torch.empty_like(pos)
    em = L.EmScalars(-0.02, 0.1414, 1.0, 1.0, 0.67, 4.0, 3.1, 1.76, 0.7, 1e-7)
    dp = L.DpmScalars(0.7, 1.01, 0.02, 0.69, 1.02, 0.04, 4.0, 4.1, 0.67, 0.65, -0.01, -0.02, 1e-7)
    cases = [
        ("se3_so3_exp", 48, lambda: ops.so3_exp(v)),
        ("se3_so3_log", 48, lambda: ops.so3_log(r)),
        ("se3_so3_compose_rotvec", 84, lambda: ops.so3_compose_rotvec(r, w, out=out_r)),
        ("se3_frame_update_em", 144, lambda: ops.frame_update_em(r, pos, w, v, z1, z2, em, rot_out=out_r, pos_out=out_p)),
        ("se3_frame_update_dpm_mid", 120, lambda: ops.frame_update_dpm_mid(r, pos, w, v, dp, rot_out=out_r, pos_out=out_p)),
        ("se3_frame_update_dpm_final", 132, lambda: ops.frame_update_dpm_final(r, pos, w, z1, v, dp, rot_out=out_r, pos_out=out_p)),
        # translation updates alone: pos + score + noise in, pos out = 48 B/residue; DPM: pos + score in, pos out = 36
        ("se3_r3_update_em", 48, lambda: ops.r3_update_em(pos, v, z2, em, pos_out=out_p)),
        ("se3_r3_update_dpm", 36, lambda: ops.r3_update_dpm(pos, v, dp, final_half=False, pos_out=out_p)),
    ]
    # IGSO3 noising (sample_marginal, so3_sde.py:249-288) with the full-size table (8 MB, L2-resident): 36 B in + 36 B out
    # + 16 B of passed-in noise = 88 B/rotation; 72 B with in-kernel Philox
    sig_grid = 0.02 * (2.33 / 0.02) ** torch.linspace(0.001, 1.0, 1000, device=device)
    om = (torch.linspace(0.0, 1, 2001, device=device, dtype=torch.float64) ** 3 * 3.141592653589793)
    cdf = ops.igso3_build_cdf(sig_grid, om, 2000)
    omg = om[1:].float()
    cidx = ops.igso3_build_cdf_index(cdf)
    sig = 0.02 * (2.33 / 0.02) ** torch.rand(n, generator=g, device=device)
    uu = torch.rand(n, generator=g, device=device)
    cases += [
        ("se3_igso3_sample(noise passed in)", 88, lambda: ops.igso3_sample(cdf, omg, n, sigma=sig, sigma_grid=sig_grid, normals=z1, u=uu, x=r, cdf_index=cidx)),
        ("se3_igso3_sample(philox)", 72, lambda: ops.igso3_sample(cdf, omg, n, sigma=sig, sigma_grid=sig_grid, seed=1, x=r, cdf_index=cidx)),
    ]
    return [_hbm_line(name, n, bytes_per, _time_alone(fn), pk) for name, bytes_per, fn in cases]


def kernel_rooflines(step_fn, L: int, B: int, heads: int = 32, sm_mhz=None, with_elementwise: bool = True):
    """(dominant-kernel roofline, other kernels).  The dominant kernel of this library inside the sampling step is the
    IPA attention operator (both passes of the tensor-core edition are timed together: one C-ABI call).

    Which roof binds: per launch the operator has 4608*L^2*B tensor-eligible flop (QK^T + P.V
    scalar/point/pair) = a few microseconds at the bf16 tensor peak, and these algorithmic bytes -- projections in,
    frames in, concat features out, the shared pair tensors once -- = tens of microseconds at the HBM peak.  Of the
    two roofs HBM is the binding one, so `bound` = "hbm"; the tensor-pipe figure is kept next to it."""
    pk = measured_peaks()
    prev = os.environ.get("SE3DIFF_B200_CUDA_GRAPH")
    prev_m = os.environ.get("SE3DIFF_B200_MODEL_GRAPH")
    os.environ["SE3DIFF_B200_CUDA_GRAPH"] = "0"      # the event hooks live in the eager launch path
    os.environ["SE3DIFF_B200_MODEL_GRAPH"] = "0"
    names = ("ipa_attention_fwd", "ipa_attention_tc_fwd")
    hooks = [_Hook(n) for n in names]
    try:
        for h in hooks:
            h.__enter__()
        step_fn()
        timed = {h.name: h.mean_ms() for h in hooks}
    finally:
        for h in hooks:
            h.__exit__()
        for key, val in (("SE3DIFF_B200_CUDA_GRAPH", prev), ("SE3DIFF_B200_MODEL_GRAPH", prev_m)):
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val
    name = max(names, key=lambda n: timed[n][1])
    ms, n = timed[name]
    rows, width_out = B * L, heads * (2 * 16 + 4 * 8)
    edition, proj_bytes, out_bytes = {
        "ipa_attention_fwd": ("fp32 SIMT", rows * heads * 96 * 4, rows * width_out * 4),
        "ipa_attention_tc_fwd": ("tcgen05 two-pass (bf16 scalar and point records from one projection)", rows * heads * 48 * (2 + 2), rows * width_out * 2),
    }[name]
    pair_el = 4 if name == "ipa_attention_fwd" else 2
    nbytes = proj_bytes + rows * 48 + out_bytes + heads * L * L * pair_el + L * L * heads * 16 * pair_el
    flops = 4608.0 * L * L * B
    gbs, tf = nbytes / ms / 1e6, flops / ms / 1e9
    # the unit that actually binds the operator: 4 sqrt per (i, j, head) for the un-squared point distances + 1 exp, on the XU pipe
    xu_ops = 5.0 * L * L * heads * B
    xu_peak = SMS * XU_LANES_PER_CLK_SM * (sm_mhz or 1965) * 1e6
    roof = {"kernel": "se3_" + name, "bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"],
            "peak_nominal": HBM_NOMINAL_GBS, "frac_nominal": gbs / HBM_NOMINAL_GBS,
            "xu": {"ops_per_launch": xu_ops, "achieved_gops": xu_ops / ms / 1e6, "peak_gops": xu_peak / 1e9, "frac": xu_ops / ms / 1e6 / (xu_peak / 1e9),
                   "peak_source": f"148 SMs x 16 MUFU lanes x {sm_mhz or 1965} MHz (SM clock {'sampled under load' if sm_mhz else 'maximum'})"},
            "traffic": _ncu_traffic(name, L, B), "launches_timed": n, "ms_per_launch": ms, "algorithmic_bytes_per_launch": nbytes,
            "peak_source": pk["source"], "edition": edition,
            "tensor": {"algorithmic_flops_per_launch": flops, "achieved_tflops": tf, "peak_tflops": pk["tensor_sustained"],
                       "frac": tf / pk["tensor_sustained"]},
            "note": "HBM is the binding roof of the two (bytes/peak >> flops/peak); the kernel itself is issue/MUFU-bound: 128 sqrt + "
                    "32 exp per (i,j) pair per layer are not a contraction (structure_module.py:170).  `traffic` is the ncu "
                    "dram read+write of both passes per call, including the un-normalised probability tiles pass 1 hands to pass 2."}
    return roof, (elementwise_rooflines() + series_rooflines(sm_mhz=sm_mhz) if with_elementwise else [])


def _ncu_traffic(name: str, L: int, B: int):
    """DRAM bytes per call from the committed `ncu --set full` capture of this kernel at this shape (profiles/), else None."""
    import json

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ipa_traffic.json")
    try:
        with open(path) as f:
            for rec in json.load(f):
                if rec["kernel"] == name and rec["L"] == L and rec["B"] == B:
                    return rec["dram_bytes_per_call"]
    except (OSError, ValueError, KeyError):
        pass
    return None
