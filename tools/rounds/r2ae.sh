#!/bin/bash
# round 2, visit ae (1 GPU): slab scan with row-piece L2 prefetch (0 / 4 / 8 chunks), ncu capture of the slab scan
mkdir -p gpurun_out
for pf in 0 4 8; do
( PG_SLAB_PREFETCH=$pf timeout 300 python bench.py --workload refiner --steps 5 --warmup 3 --refiner-schedule 4 2> gpurun_out/r2ae_stderr.log | tail -1 ) > gpurun_out/r2ae_refiner_k5_pf${pf}.json; python -c "
import json;d=json.load(open('gpurun_out/r2ae_refiner_k5_pf${pf}.json'));print('refiner k5 sched 4 prefetch ${pf}:',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['scan_ms'])"; tail -2 gpurun_out/r2ae_stderr.log
done
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:slab_scan_kernel -s 3 -c 1 -o gpurun_out/r2ae_prof_slab_k5 -f python bench.py --workload refiner --steps 2 --warmup 2 --refiner-schedule 4 > gpurun_out/r2ae_ncu_slab.log 2>&1; tail -1 gpurun_out/r2ae_ncu_slab.log
python tools/ncu_summary.py gpurun_out/r2ae_prof_slab_k5.ncu-rep --stalls --sass 25 > gpurun_out/r2ae_slab_k5_ncu_summary.txt 2>&1
ncu -i gpurun_out/r2ae_prof_slab_k5.ncu-rep --page raw --csv > gpurun_out/r2ae_slab_raw.csv 2>/dev/null; rm -f gpurun_out/r2ae_prof_slab_k5.ncu-rep
head -45 gpurun_out/r2ae_slab_k5_ncu_summary.txt | cut -c1-160
