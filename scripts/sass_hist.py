#!/usr/bin/env python
"""cuobjdump -sass opcode histogram of the tensor-core / TMA kernels -> profiles/<tag>_sass_histogram.txt
(evidence for the tcgen05 / TMEM / TMA claims in DESIGN.md; run here, needs no GPU)."""
import collections, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
csrc = os.path.join(root, "neural-qmc-wavefunction_b200", "csrc")
objs = ["nqmc_orbital_lap_pc.o", "nqmc_orbital_fwd_pc.o", "nqmc_orbital_val2_tc.o", "nqmc_orbital_bwd_tc.o", "nqmc_orbital_bwd128.o",
        "nqmc_gram.o"]
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMAPF", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG",
       "BAR", "UTCATOMSWS"]
lines = [f"# cuobjdump -sass opcode counts per kernel (static instruction counts, not executed counts); objects built by __graft_entry__.build()",
         f"# columns: {' '.join(KEY)} | total"]
for o in objs:
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(csrc, o)], capture_output=True, text=True).stdout
    cur, hist = None, None
    res = []
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            if cur:
                res.append((cur, hist))
            cur, hist = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
        if m and cur:
            hist[m.group(1)] += 1
    if cur:
        res.append((cur, hist))
    lines.append(f"\n## {o}")
    for name, h in res:
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        dem = re.sub(r"\((?:int|bool)\)", "", dem.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).split("(")[0]
        if not any(h[k] for k in ("UTCHMMA", "LDTM", "STTM", "UTMALDG")):
            continue
        lines.append(f"{dem[:70]:70s} " + " ".join(f"{k}={h[k]}" for k in KEY if h[k]) + f" | {sum(h.values())}")
open(os.path.join(root, "profiles", f"{tag}_sass_histogram.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
