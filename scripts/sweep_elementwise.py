"""BASELINE config 3: IGSO3 noising + SO(3) exp / log / compose + fused frame updates, achieved HBM GB/s for n = 1e4 .. 1e7
rotations (each kernel alone, CUDA events, algorithmic bytes of SURVEY.md 8d).  Writes a markdown table to stdout."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from se3diff_b200.profiling import elementwise_rooflines, measured_peaks
sizes = [10_000, 100_000, 1_000_000, 10_000_000]
res = {n: elementwise_rooflines(n) for n in sizes}
pk = measured_peaks()
print(f"| kernel | B/unit | " + " | ".join(f"n = {n:.0e}: GB/s (frac of {res[n][0]['peak']:.0f})" for n in sizes) + " |")
print("|---|---:|" + "---:|" * len(sizes))
for k in range(len(res[sizes[0]])):
    r0 = res[sizes[0]][k]
    print(f"| `{r0['kernel']}` | {r0['bytes_per_unit']} | " + " | ".join(f"{res[n][k]['achieved']:.0f} ({res[n][k]['frac']:.2f})" for n in sizes) + " |")
