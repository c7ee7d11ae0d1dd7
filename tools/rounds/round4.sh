#!/bin/bash
mkdir -p gpurun_out
PG_TEST_TIMEOUT=200 timeout 900 python tools/run_gpu_tests_isolated.py tests/test_gpu_kernels.py -k "attention or gemm" > gpurun_out/k_tests.log 2>&1
grep -E "^(FAIL)|passed|failed|dead-lock|rror" gpurun_out/k_tests.log | tail -20
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python tools/microbench.py 128 > gpurun_out/microbench_r4.log 2>&1; tail -8 gpurun_out/microbench_r4.log; cp gpurun_out/microbench.json gpurun_out/microbench_r4.json
