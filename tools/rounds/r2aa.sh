#!/bin/bash
# round 2, visit aa (1 GPU): mixed LDS + FFMA2 micro-benchmark; l1tex / shared-memory counters of the tile scan
mkdir -p gpurun_out
timeout 100 tools/ubench/fp32_lds > gpurun_out/r2aa_ubench_fp32_lds.txt 2>&1; grep -E "mixed|LDS.128" gpurun_out/r2aa_ubench_fp32_lds.txt
NCU="ncu --set full --clock-control none"
timeout 600 $NCU -k regex:tile_scan_kernel -s 3 -c 1 -o gpurun_out/r2aa_prof_tile_k5 -f python bench.py --workload refiner --steps 2 --warmup 2 --refiner-schedule 3 > gpurun_out/r2aa_ncu_tile.log 2>&1; tail -1 gpurun_out/r2aa_ncu_tile.log
ncu -i gpurun_out/r2aa_prof_tile_k5.ncu-rep --page raw --csv > gpurun_out/r2aa_tile_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2aa_tile_raw.csv')))
h, u, r = rows[0], rows[1], rows[2]
for n, uu, v in zip(h, u, r):
    if any(t in n for t in ("shared", "l1tex__throughput", "lsu", "l1tex__data_pipe", "l1tex__f_", "bank_conflict", "pipe_fma", "pipe_lsu", "smsp__inst_executed.sum", "l1tex__lsuin", "breakdown")):
        print(n, '=', v, uu)
PY
rm -f gpurun_out/r2aa_prof_tile_k5.ncu-rep
