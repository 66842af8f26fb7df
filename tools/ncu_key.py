"""Prints the handful of ncu raw metrics used to judge the integer kernels.  python tools/ncu_key.py <report.ncu-rep>"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ["gpu__time_duration.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed.sum", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum"]
for r in rows[2:]:
    d = dict(zip(h, r))
    print(d.get("Kernel Name", "?")[:80])
    for k in want:
        if k in d: print(f"  {k} = {d[k]}")
    for k, v in d.items():
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(v or 0) > 0.3: print(f"  {k[34:-23]} = {v}")
