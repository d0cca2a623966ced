"""ncu launch-list CSV -> markdown share table (profiles/)."""
import collections, csv, gzip, shutil, sys
src, tag, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[hdr_i], rows[hdr_i + 1:]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    agg[r[ki]][0] += 1; agg[r[ki]][1] += v
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if '<unnamed>' in k or 'anonymous' in k)
with open(f'profiles/{tag}_launch_list_summary.md', 'w') as f:
    f.write(f"# {tag}: ncu launch list\n\nCommand (B200; the same command exited 0 without ncu first):\n\n```\n{cmd}\n```\n\n")
    f.write(f"{sum(v[0] for v in agg.values())} launches, {tot/1e3:.1f} ms of kernel time (cold-cache, serialised by ncu: compare SHARES). "
            f"Kernels of libse3diff_b200.so: {100*ours/tot:.1f}% of the time.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:28]:
        f.write(f"| `{k[:120]}` | {v[0]} | {v[1]:.1f} | {100*v[1]/tot:.1f}% |\n")
with open(src, 'rb') as fi, gzip.open(f'profiles/{tag}_launch_list.csv.gz', 'wb') as fo:
    shutil.copyfileobj(fi, fo)
print(open(f'profiles/{tag}_launch_list_summary.md').read()[:3000])
