"""Per-SASS-instruction executed counts of one kernel from an ncu report (--import-source on): the instruction mix that
the warps actually issued (dynamic, not static), grouped by opcode, plus the hottest address ranges.
usage: python scripts/ncu_hot_sass.py report.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["Instructions Executed"]) for r in rows)
print(f"{len(rows)} SASS instructions, {tot} warp instructions executed")
by = collections.Counter()
for r in rows:
    src = r["Source"].strip()
    parts = src.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    by[op.split(".")[0]] += int(r["Instructions Executed"])
for op, n in by.most_common(top):
    print(f"  {op:12s} {n:12d} {100.0 * n / tot:5.1f} %")
if len(sys.argv) > 3:   # dump every instruction with its count
    for i, r in enumerate(rows):
        print(i, r["Instructions Executed"], r["# Samples"], r["Source"].strip())
