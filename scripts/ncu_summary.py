"""Summarise an .ncu-rep (raw + source pages) into text: python scripts/ncu_summary.py file.ncu-rep [min_exec]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
minex = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg.per_second"]
for h, u, v in zip(hdr, units, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        print(f"{h:86s} {v:>22s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))[2:]
tot = sum(int(r[5]) for r in rows if r[5].isdigit())
hot = [r for r in rows if r[5].isdigit() and int(r[5]) >= minex * 1e6]
print(f"\ninstructions executed total {tot/1e6:.1f} M; hot (>= {minex} M) {sum(int(r[5]) for r in hot)/1e6:.1f} M in {len(hot)} SASS lines")
for r in hot:
    print(f"{r[0][-5:]} {r[1][:72]:72s} ex={int(r[5])/1e6:6.2f}M samp={r[4]:>5s}")
