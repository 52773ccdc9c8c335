"""Summarise the source page of an ncu report: stall totals, hottest SASS lines, samples per BAR.SYNC segment.
usage: python scripts/ncu_stalls.py gpurun_out/prof_x.ncu-rep [n_lines]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        h = i
        break
hdr = rows[h]
data = []
for r in rows[h + 1:]:          # multi-kernel reports repeat the header: summarise the first kernel only
    if r and r[0] == "Address":
        break
    if len(r) == len(hdr):
        data.append(r)
isrc, iex, ins = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ins] or 0) for r in data)
print("samples", tot, "warp instructions", sum(int(r[iex] or 0) for r in data))
agg = collections.Counter()
for r in data:
    for i in stall_cols:
        agg[hdr[i]] += int(r[i] or 0)
print({k: v for k, v in agg.most_common(10)})
seg = 0
segs = collections.OrderedDict()
lines = []
for k, r in enumerate(data):
    if "BAR.SYNC" in r[isrc]:
        seg += 1
    d = segs.setdefault(seg, [0, 0])
    d[0] += int(r[ins] or 0)
    d[1] += int(r[iex] or 0)
    lines.append((int(r[ins] or 0), k, seg, r))
for n, k, sg, r in sorted(lines, key=lambda x: -x[0])[:n_lines]:
    st = sorted([(hdr[i][6:], int(r[i] or 0)) for i in stall_cols], key=lambda x: -x[1])[:2]
    print(str(n).rjust(6), f"line {k} seg {sg}", st, r[isrc][:80])
print("segments (samples, %):", {k: (v[0], round(100 * v[0] / max(tot, 1), 1)) for k, v in segs.items()})
