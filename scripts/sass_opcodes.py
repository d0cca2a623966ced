"""Blackwell evidence: counts of the tensor-core / TMEM / TMA SASS mnemonics per kernel of the built library
(`cuobjdump -sass`; the PTX names never appear in SASS: tcgen05.mma -> UTCHMMA, tcgen05.ld/st -> LDTM/STTM,
cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS/UTCALLOC).
usage: python scripts/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "se3diff_b200", "_lib", "libse3diff_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
ops = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU.SQRT", "MUFU.EX2", "LDGSTS", "HMMA")
per, name, arch = collections.OrderedDict(), None, set()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0]
        per[name] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if name:
        for o in ops:
            if re.search(r"\b" + re.escape(o) + r"\b", line):
                per[name][o] += 1
total = collections.Counter()
print(f"# {os.path.relpath(lib, root)}: SASS mnemonic counts per kernel (cuobjdump -sass), arch {sorted(arch)}")
print("# " + "  ".join(f"{o:>9s}" for o in ops) + "  kernel")
for k, c in per.items():
    total.update(c)
    if any(c[o] for o in ops[:6]):
        print("  " + "  ".join(f"{c[o]:9d}" for o in ops) + "  " + k[:110])
print("  " + "  ".join(f"{total[o]:9d}" for o in ops) + "  TOTAL (all %d kernels)" % len(per))
assert total["HMMA"] == 0, "legacy mma.sync tensor path present"
