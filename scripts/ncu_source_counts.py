"""Dynamic instruction mix of the kernels in an ncu report taken with `--set full --import-source on` (developer diagnostics).
usage: python scripts/ncu_source_counts.py report.ncu-rep [kernel-name substring]
Per kernel instance: warp instructions per warp, the executed-opcode histogram (instructions per warp), the ten SASS lines with the
most stall samples."""
import csv, io, subprocess, sys
from collections import Counter

rep, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kernels, cur = [], None
for r in csv.reader(io.StringIO(txt)):
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], rows=[]); kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
for k in kernels:
    if pat not in k["name"]:
        continue
    h = k["hdr"]; ie, ss = h.index("Instructions Executed"), h.index("# Samples")
    warps = int(k["rows"][0][ie])                           # the first instruction is executed once by every warp
    total = sum(int(r[ie]) for r in k["rows"])
    print(f"== {k['name'][:90]}\n   {len(k['rows'])} SASS lines, {warps} warps, {total / warps:.1f} warp instructions per warp")
    mix, smp = Counter(), Counter()
    for r in k["rows"]:
        f = r[1].split()
        op = (f[1] if f[0].startswith("@") else f[0]).split(".")[0]
        mix[op] += int(r[ie]) / warps; smp[op] += int(r[ss])
    print("   opcode: instructions per warp (stall samples)")
    print("   " + "  ".join(f"{op} {n:.1f} ({smp[op]})" for op, n in mix.most_common(24)))
    print("   lines with the most stall samples:")
    for r in sorted(k["rows"], key=lambda r: -int(r[ss]))[:10]:
        print(f"     {int(r[ss]):6d}  x{int(r[ie]) / warps:4.2f}  {r[1].strip()[:80]}")
