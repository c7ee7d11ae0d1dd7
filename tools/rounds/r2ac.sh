#!/bin/bash
# round 2, visit ac (1 GPU): refiner slab scan (v5) — parity over all schedules, cfg5 k = 5 and k = 40
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "refiner" 2>&1 | tail -15 ) > gpurun_out/r2ac_pytest_refiner.log; tail -8 gpurun_out/r2ac_pytest_refiner.log
for k in 5 40; do
( timeout 300 python bench.py --workload refiner --steps 5 --warmup 3 --refiner-schedule 4 --refiner-topk $k 2> gpurun_out/r2ac_refiner_k${k}_stderr.log | tail -1 ) > gpurun_out/r2ac_refiner_k${k}_s4.json; python -c "
import json;d=json.load(open('gpurun_out/r2ac_refiner_k${k}_s4.json'));print('refiner k${k} sched 4:',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['fp32_fma_tflops'],d['family_ms_per_step'])"; tail -3 gpurun_out/r2ac_refiner_k${k}_stderr.log
done
