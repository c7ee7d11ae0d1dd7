#!/bin/bash
# round 2, visit w (1 GPU): refiner tile scan (v4) — parity tests over all three schedules, cfg5 A/B against the cell-major kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -x -q -m gpu -k "refiner" 2>&1 | tail -15 ) > gpurun_out/r2w_pytest_refiner.log; tail -15 gpurun_out/r2w_pytest_refiner.log
for sched in 2 3; do
( timeout 300 python bench.py --workload refiner --steps 5 --warmup 3 --refiner-schedule $sched 2> gpurun_out/r2w_refiner_s${sched}_stderr.log | tail -1 ) > gpurun_out/r2w_refiner_k5_s${sched}.json; python -c "
import json;d=json.load(open('gpurun_out/r2w_refiner_k5_s${sched}.json'));print('refiner k5 sched ${sched}:',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['scan_ms'],d['family_ms_per_step'])"; tail -3 gpurun_out/r2w_refiner_s${sched}_stderr.log
done
( timeout 300 python bench.py --workload refiner --steps 5 --warmup 3 --refiner-schedule 3 --refiner-topk 40 2> gpurun_out/r2w_refiner_k40_stderr.log | tail -1 ) > gpurun_out/r2w_refiner_k40_s3.json; python -c "
import json;d=json.load(open('gpurun_out/r2w_refiner_k40_s3.json'));print('refiner k40 sched 3:',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['fp32_fma_tflops'],d['family_ms_per_step'])"; tail -3 gpurun_out/r2w_refiner_k40_stderr.log
