"""One pass over the elementwise kernels at n = 1e7 (ncu target)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from se3diff_b200.profiling import elementwise_rooflines
for r in elementwise_rooflines():
    print(r["kernel"], round(r["ms"], 4), "ms", round(r["frac"], 3))
