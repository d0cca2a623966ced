"""IGSO3 sampler at n = 1e7: how much of its time is the guide-record working set falling out of L2?  (developer diagnostics)
Variants: full table (1000 sigma rows: 32 MB of records) with / without the persisting-L2 window, a 100-row table (3 MB), one shared sigma."""
import math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from se3diff_b200 import ops
from se3diff_b200.profiling import _time_alone, measured_peaks
dev, n = "cuda", 10_000_000
g = torch.Generator(device=dev).manual_seed(0)
x = ops.so3_exp(torch.randn(n, 3, generator=g, device=dev))
z = torch.randn(n, 3, generator=g, device=dev); uu = torch.rand(n, generator=g, device=dev)
om = (torch.linspace(0.0, 1, 2001, device=dev, dtype=torch.float64) ** 3 * math.pi)
pk = measured_peaks()["hbm"]
for rows in (1000, 100):
    grid = 0.02 * (2.33 / 0.02) ** torch.linspace(0.001, 1.0, rows, device=dev)
    cdf = ops.igso3_build_cdf(grid, om, 2000); idx = ops.igso3_build_cdf_index(cdf); omg = om[1:].float()
    for name, sig in (("random sigma", 0.02 * (2.33 / 0.02) ** torch.rand(n, generator=g, device=dev)), ("one sigma", torch.full((n,), 0.5, device=dev))):
        for mode, kw in (("noise passed in", dict(normals=z, u=uu)), ("philox", dict(seed=1))):
            ms = _time_alone(lambda: ops.igso3_sample(cdf, omg, n, sigma=sig, sigma_grid=grid, x=x, cdf_index=idx, **kw))
            b = 88 if "noise" in mode else 72
            print(f"rows={rows:5d} {name:13s} {mode:16s} L2_WINDOW={os.environ.get('SE3DIFF_B200_L2_WINDOW', '1')}: {ms:.3f} ms  {b * n / ms / 1e6:7.1f} GB/s = {b * n / ms / 1e6 / pk:.3f} of measured")
