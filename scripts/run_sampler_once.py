"""One launch each of the IGSO3 sampler (noise passed in, Philox) at n = 1e7 on the full-size table (ncu target)."""
import math, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from se3diff_b200 import ops
dev, n = "cuda", 10_000_000
g = torch.Generator(device=dev).manual_seed(0)
x = ops.so3_exp(torch.randn(n, 3, generator=g, device=dev))
z = torch.randn(n, 3, generator=g, device=dev); uu = torch.rand(n, generator=g, device=dev)
om = (torch.linspace(0.0, 1, 2001, device=dev, dtype=torch.float64) ** 3 * math.pi)
grid = 0.02 * (2.33 / 0.02) ** torch.linspace(0.001, 1.0, 1000, device=dev)
cdf = ops.igso3_build_cdf(grid, om, 2000); idx = ops.igso3_build_cdf_index(cdf); omg = om[1:].float()
sig = 0.02 * (2.33 / 0.02) ** torch.rand(n, generator=g, device=dev)
for _ in range(2):
    ops.igso3_sample(cdf, omg, n, sigma=sig, sigma_grid=grid, x=x, cdf_index=idx, normals=z, u=uu)
    ops.igso3_sample(cdf, omg, n, sigma=sig, sigma_grid=grid, x=x, cdf_index=idx, seed=1)
torch.cuda.synchronize()
print("ok")
