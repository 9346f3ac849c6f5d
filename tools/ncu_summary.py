#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (read here, without a GPU): python tools/ncu_summary.py report.ncu-rep [out.md]
Prints, per profiled launch, the metrics the roofline discussion needs (duration, DRAM bytes, L2 hit rate, issue / pipe
utilisation, occupancy, registers) and the top stall reasons, and the ten most stalled SASS instructions."""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__inst_executed.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"## {d.get('Kernel Name', '?')[:160]}", file=out)
    for k in WANT:
        if k in d:
            print(f"  {k:72s} {d[k]} {units[hdr.index(k)]}", file=out)
    st = []
    for k, v in d.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            try:
                st.append((float(v), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    print("  stall reasons (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:6]), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[1]
    if "# Samples" in h:
        ia, isamp = h.index("Source"), h.index("# Samples")
        stall = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        data = []
        for r in rows[2:]:
            try:
                data.append((int(r[isamp]), r))
            except (ValueError, IndexError):
                pass
        tot = sum(n for n, _ in data) or 1
        print("## most stalled instructions (share of warp-stall samples, top reason)", file=out)
        for n, r in sorted(data, key=lambda t: -t[0])[:10]:
            top = max(((int(r[i]) if r[i].isdigit() else 0, x) for i, x in stall), default=(0, ""))
            print(f"  {100 * n / tot:5.1f}%  {r[ia].strip()[:70]:70s} {top[1]}", file=out)
