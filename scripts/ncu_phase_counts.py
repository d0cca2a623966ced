"""Dynamic warp instructions and stall samples per PHASE of k_ipa_tc_pass1, from the source page of an `ncu --set full --import-source on`
capture: the kernel's clock64 phase stamps (SE3_STAMP -> `CS2R ..., SR_CLOCKLO` in SASS) delimit the phases.
usage: python scripts/ncu_phase_counts.py <report.ncu-rep> [items per launch, default 8192 = 256 samples x 32 heads]"""
import collections, csv, re, subprocess, sys

rep = sys.argv[1]
items = float(sys.argv[2]) if len(sys.argv) > 2 else 8192.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:pass1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
si, ei, st, bi = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("stall_barrier")
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                                  # first captured launch only
    if len(r) > bi:
        body.append((r[si].strip(), int(r[ei]), int(r[st]), int(r[bi])))
print(rows[0][1][:110])
tot = sum(b[1] for b in body)
print(f"{len(body)} SASS lines, {tot / items:.0f} warp instructions per item (four warps per CTA)")
stamps = [i for i, b in enumerate(body) if "CS2R" in b[0] and "SR_CLOCKLO" in b[0]]
names = ["prologue", "staging / prefetch", "", "", "frame transforms", "", "MMA 1 issue / wait", "logit pass", "exponential pass", "MMA 2 issue / wait", "epilogue", "", "tail"]
prev = 0
for k, n in enumerate(stamps + [len(body)]):
    seg = body[prev:n]
    inst, stall, barrier = sum(b[1] for b in seg), sum(b[2] for b in seg), sum(b[3] for b in seg)
    ops = collections.Counter()
    for s, e, _, _ in seg:
        m = re.match(r"(@!?U?P\d\s+)?([A-Z0-9_]+)", s)
        if m:
            ops[m.group(2)] += e
    name = names[k] if k < len(names) else ""
    if inst / items >= 20:
        print(f"{name:22s} {inst / items:8.1f} inst/item ({100 * inst / tot:4.1f} %)  stall samples {stall:5d} (at barriers {barrier:5d})   "
              + ", ".join(f"{o} {c / items:.0f}" for o, c in ops.most_common(8)))
    prev = n
