"""Condense an .ncu-rep (one kernel, --set full) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/rNN_ncu_<kernel>.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
col = {k: i for i, k in enumerate(h)}
print("kernel:", v[col["Kernel Name"]])
print("grid x block:", v[col["Grid Size"]], "x", v[col["Block Size"]])
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]
for k in want:
    if k in col:
        print(f"{k}: {v[col[k]]} {u[col[k]]}")
print("stalls per issued instruction (> 0.2):")
for k in h:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
        try:
            x = float(v[col[k]])
        except ValueError:
            continue
        if x > 0.2:
            print("  ", k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], round(x, 2))
