#!/bin/bash
# round 2, visit z (1 GPU): ncu capture of the refiner tile scan (cfg5, k = 5) with warp-stall reasons and hot SASS lines;
# fused head after the tail fix (tests + timing)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:tile_scan_kernel -s 3 -c 1 -o gpurun_out/r2z_prof_tile_k5 -f python bench.py --workload refiner --steps 2 --warmup 2 --refiner-schedule 3 > gpurun_out/r2z_ncu_tile.log 2>&1; tail -2 gpurun_out/r2z_ncu_tile.log
python tools/ncu_summary.py gpurun_out/r2z_prof_tile_k5.ncu-rep --stalls --sass 40 > gpurun_out/r2z_tile_k5_ncu_summary.txt 2>&1; rm -f gpurun_out/r2z_prof_tile_k5.ncu-rep
cat gpurun_out/r2z_tile_k5_ncu_summary.txt | cut -c1-200
( timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "test_head" 2>&1 | tail -5 ) > gpurun_out/r2z_pytest_head.log; tail -3 gpurun_out/r2z_pytest_head.log
for f in 1 0; do PG_HEAD_FUSED=$f timeout 120 python tools/head_time.py 2>&1 | tail -1; done | tee gpurun_out/r2z_head_time.log
