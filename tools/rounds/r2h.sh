#!/bin/bash
# round 2, visit h: full GPU test suite with the pair attention kernel + LayerNorm-folded GEMMs, bench A/B of the fold
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/r2h_pytest_gpu.log; tail -12 gpurun_out/r2h_pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2h_bench_stderr.log | tail -1 ) > gpurun_out/r2h_bench.json; python -c "
import json;d=json.load(open('gpurun_out/r2h_bench.json'));print('fold   :',d['value'],d['ms_per_step'],d['kernel_ms_per_step'],d['clocks'])"; tail -3 gpurun_out/r2h_bench_stderr.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ln-fold 2> gpurun_out/r2h_bench_nofold_stderr.log | tail -1 ) > gpurun_out/r2h_bench_nofold.json; python -c "
import json;d=json.load(open('gpurun_out/r2h_bench_nofold.json'));print('no fold:',d['value'],d['ms_per_step'],d['kernel_ms_per_step'],d['clocks'])"; tail -3 gpurun_out/r2h_bench_nofold_stderr.log
cat gpurun_out/parity.log 2>/dev/null | tail -8
