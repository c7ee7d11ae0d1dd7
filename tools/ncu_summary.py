"""Turn an .ncu-rep (ncu --set full) into the short per-launch text summary kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep > profiles/rNN_x_ncu_summary.txt
    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep --stalls --sass 60   # + warp-stall reasons, hottest SASS lines
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
want_stalls = "--stalls" in sys.argv
n_sass = int(sys.argv[sys.argv.index("--sass") + 1]) if "--sass" in sys.argv else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(hdr)}
for li, r in enumerate(rows[2:]):
    print(f"== launch {li}: {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    for k in KEEP:
        if k in col:
            print(f"   {k} = {r[col[k]]} {units[col[k]]}")
    if want_stalls:
        st = [(float(r[i].replace(",", "") or 0), n) for n, i in col.items()
              if n.startswith("smsp__average_warp") and "issue_stalled" in n and n.endswith("_per_issue_active.ratio")
              or n.startswith("smsp__average_warps_issue_stalled") and n.endswith(".ratio")]
        for v, n in sorted(st, reverse=True)[:12]:
            print(f"   stall {n} = {v:.3f}")

if n_sass:
    # hottest SASS instructions by warp-stall samples (source page; needs --import-source on / -lineinfo)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    for idx, r in enumerate(rows):
        if "Source" in r and any("Sampling" in c for c in r):
            h = idx
            break
    if h is not None:
        hd = rows[h]
        c_src = hd.index("Source")
        c_smp = next(i for i, c in enumerate(hd) if c.startswith("Warp Stall Sampling (All"))
        c_ni = next((i for i, c in enumerate(hd) if c.startswith("Warp Stall Sampling (Not")), c_smp)
        c_ex = next((i for i, c in enumerate(hd) if c == "Instructions Executed"), None)
        body = [r for r in rows[h + 1:] if len(r) > c_smp and r[c_smp].replace(",", "").isdigit()]
        tot = sum(int(r[c_smp].replace(",", "")) for r in body) or 1
        print(f"== SASS hot spots ({tot} samples, {len(body)} instructions)")
        ranked = sorted(enumerate(body), key=lambda t: -int(t[1][c_smp].replace(",", "")))[:n_sass]
        for pos, r in sorted(ranked):
            ex = r[c_ex] if c_ex is not None else ""
            print(f"   #{pos:5d} {100.0 * int(r[c_smp].replace(',', '')) / tot:5.2f}%  notissued={r[c_ni]:>7} exec={ex:>9}  {r[c_src][:110]}")
        # cumulative share by position (to map regions = roles)
        acc, step = 0, max(1, len(body) // 40)
        for k in range(0, len(body), step):
            seg = sum(int(r[c_smp].replace(",", "")) for r in body[k:k + step])
            print(f"   region #{k:5d}-{min(len(body), k + step) - 1:5d}: {100.0 * seg / tot:5.2f}%")
