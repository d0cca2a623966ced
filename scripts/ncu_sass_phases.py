"""Dynamic instruction count of a kernel from an ncu report's SASS page, cut into segments at barriers / tensor-core
instructions, plus an opcode histogram per segment (developer diagnostics).
usage: python scripts/ncu_sass_phases.py report.ncu-rep kernel_regex"""
import collections, csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
lines = txt.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
seg, segs, total = 0, collections.defaultdict(lambda: [0, 0, collections.Counter(), ""]), 0
CUT = ("BAR.SYNC", "UTCHMMA", "SYNCS.PHASECHK", "UTCBAR")
for r in rows:
    src = r["Source"].strip()
    try:
        n = int(r["Instructions Executed"]); samples = int(r["# Samples"])
    except (ValueError, KeyError, TypeError):
        continue
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    if any(c in src for c in CUT):
        seg += 1
        segs[seg][3] = src[:60]
    s = segs[seg]
    parts = op.split(".")
    name = parts[0] + ("." + parts[1] if parts[0] in ("MUFU", "LDS", "STS", "LDG") and len(parts) > 1 else "")
    s[0] += n; s[1] += samples; s[2][name] += n
    total += n
print("total warp-instructions", total)
for k in sorted(segs):
    n, smp, hist, tag = segs[k]
    if n < total * 0.004:
        continue
    print(f"seg {k:3d} after [{tag}]: {n:>11d} inst ({100*n/total:4.1f}%)  samples {smp:>7d}   top:", ", ".join(f"{o} {c*100//n}%" for o, c in hist.most_common(9)))
