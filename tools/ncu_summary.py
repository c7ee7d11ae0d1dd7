"""Turn an .ncu-rep (ncu --set full) into the short per-launch text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/rNN_x_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
        "lts__t_sector_hit_rate.pct"]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
for li, r in enumerate(rows[2:]):
    print(f"== launch {li}: {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    for k in KEEP:
        if k in col:
            print(f"   {k} = {r[col[k]]} {units[col[k]]}")
