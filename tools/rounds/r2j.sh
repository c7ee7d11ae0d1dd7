#!/bin/bash
# round 2, visit j: full GPU suite with the new tests, bench of the three workloads on one GPU
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2j_pytest_gpu.log; tail -12 gpurun_out/r2j_pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > gpurun_out/r2j_smoke.log; cat gpurun_out/r2j_smoke.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2j_bench_stderr.log | tail -1 ) > gpurun_out/r2j_bench.json; python -c "
import json;d=json.load(open('gpurun_out/r2j_bench.json'));print('infer:',d['value'],d['ms_per_step'],d['kernel_ms_per_step'],d['parity_check'],d['clocks'])"; tail -3 gpurun_out/r2j_bench_stderr.log
for k in 5 40; do
( timeout 600 python bench.py --workload refiner --refiner-topk $k --steps 5 --warmup 3 2> gpurun_out/r2j_refiner_k${k}_stderr.log | tail -1 ) > gpurun_out/r2j_refiner_k$k.json; python -c "
import json;d=json.load(open('gpurun_out/r2j_refiner_k$k.json'));print('refiner k=$k:',d['value'],d['ms_per_step'],d['roofline'],d['family_ms_per_step'])"; tail -3 gpurun_out/r2j_refiner_k${k}_stderr.log
done
( timeout 900 python bench.py --workload train --steps 2 --warmup 1 2> gpurun_out/r2j_train_stderr.log | tail -1 ) > gpurun_out/r2j_train.json; python -c "
import json;d=json.load(open('gpurun_out/r2j_train.json'));print('train:',d['value'],d['ms_per_step'],d['max_mem_gb'],list(d['family_ms_per_step'].items())[:8])"; tail -3 gpurun_out/r2j_train_stderr.log
