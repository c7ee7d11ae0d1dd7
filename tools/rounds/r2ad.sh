#!/bin/bash
# round 2, visit ad (1 GPU): explicit shared-memory accesses (slab scan, fused head, GEMM epilogue staging) — tests, refiner
# cfg5 k = 5 / 40, head timing, one inference bench line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "gemm or head or tile_scan or layernorm" 2>&1 | tail -6 ) > gpurun_out/r2ad_pytest.log; tail -4 gpurun_out/r2ad_pytest.log
for k in 5 40; do
( timeout 300 python bench.py --workload refiner --steps 5 --warmup 3 --refiner-schedule 4 --refiner-topk $k 2> gpurun_out/r2ad_refiner_k${k}_stderr.log | tail -1 ) > gpurun_out/r2ad_refiner_k${k}_s4.json; python -c "
import json;d=json.load(open('gpurun_out/r2ad_refiner_k${k}_s4.json'));print('refiner k${k} sched 4:',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['fp32_fma_tflops'],d['family_ms_per_step'])"; tail -3 gpurun_out/r2ad_refiner_k${k}_stderr.log
done
for f in 1 0; do PG_HEAD_FUSED=$f timeout 120 python tools/head_time.py 2>&1 | tail -1; done | tee gpurun_out/r2ad_head_time.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2ad_bench_stderr.log | tail -1 ) > gpurun_out/r2ad_bench_n1.json; python -c "
import json;d=json.load(open('gpurun_out/r2ad_bench_n1.json'));print('infer n1:',d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check'],d['roofline']['frac'],d.get('roofline_attention',{}).get('frac'),d['clocks'],d.get('family_ms_per_step'))"; tail -2 gpurun_out/r2ad_bench_stderr.log
