#!/usr/bin/env python
"""Turn gpurun_out/{launches.csv,prof_*.ncu-rep} into the text summaries committed under profiles/."""
import collections, csv, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
rep = sys.argv[2] if len(sys.argv) > 2 else (f"gpurun_out/{tag}_prof.ncu-rep" if __import__("os").path.exists(f"gpurun_out/{tag}_prof.ncu-rep") else f"gpurun_out/prof_{tag}.ncu-rep")
import os
launches = f"gpurun_out/{tag}_launches.csv" if os.path.exists(f"gpurun_out/{tag}_launches.csv") else "gpurun_out/launches.csv"
rows = list(csv.reader(open(launches)))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[h], rows[h + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    a = agg.setdefault(r[ki].split("(")[0][:90], [0, 0.0])
    a[0] += 1
    a[1] += v
own = lambda k: not any(t in k for t in ('cutlass', 'ffma_peak', 'at::', 'cublas'))   # the bench also measures its roofline denominators
tot = sum(a[1] for k, a in agg.items() if own(k))
with open(f"profiles/{tag}_launches_summary.txt", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  (python bench.py --steps 2 --warmup 3 --burn 3 --no-cpu-baseline)\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES (share = of this repository's kernels; the cuBLAS / FFMA peak probes are marked)\n")
    f.write(f"{'total_us':>12} {'launches':>8} {'us/launch':>10} {'share':>7}  kernel\n")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        share = f"{100 * t / tot:6.1f}%" if own(k) else "  (peak)"
        f.write(f"{t:12.1f} {n:8d} {t / n:10.1f} {share}  {k}\n")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units, body = rr[0], rr[1], rr[2:]
keep = [i for i, n in enumerate(hdr) if any(s in n for s in (
    "Kernel Name", "gpu__time_duration.sum", "sm__throughput.avg.pct", "pipe_fma", "pipe_lsu.avg.pct", "pipe_xu.avg.pct",
    "pipe_alu.avg.pct", "pipe_tensor_cycles_active", "pipe_tensor_op", "inst_executed_pipe_uniform", "issue_active.avg.pct", "warps_active.avg.pct", "registers_per_thread",
    "occupancy_limit", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput", "bank_conflicts_pipe_lsu_mem_shared.sum",
    "wavefronts_mem_shared.sum", "issue_stalled", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "shared_mem_per_block",
    "launch__grid_size", "launch__block_size", "thread_inst_executed_op_ffma", "lts__t_bytes.sum"))]
with open(f"profiles/{tag}_ncu_full_metrics.txt", "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on   report: {rep}\n")
    for i in keep:
        f.write(f"{hdr[i]} [{units[i]}]\n")
        for b in body:
            f.write(f"    {b[hdr.index('Kernel Name')].split('(')[0][:60]:60s} {b[i]}\n")
print("wrote profiles/", tag)
