#!/bin/bash
# A/B of the GEMM variants: correctness (isolated) then microbench with the 2-CTA kernel and with PG_GEMM_1CTA=1.
mkdir -p gpurun_out
PG_TEST_TIMEOUT=200 timeout 900 python tools/run_gpu_tests_isolated.py tests/test_gpu_kernels.py -k "gemm or head" > gpurun_out/gemm_tests_2cta.log 2>&1
grep -E "^(PASS|FAIL)|passed|failed" gpurun_out/gemm_tests_2cta.log | tail -20
timeout 300 python tools/microbench.py 128 > gpurun_out/microbench_2cta.log 2>&1; tail -8 gpurun_out/microbench_2cta.log; cp gpurun_out/microbench.json gpurun_out/microbench_2cta.json
PG_GEMM_1CTA=1 timeout 300 python tools/microbench.py 128 > gpurun_out/microbench_1cta.log 2>&1; tail -8 gpurun_out/microbench_1cta.log; cp gpurun_out/microbench.json gpurun_out/microbench_1cta.json
timeout 600 python -m pytest tests/test_gpu_vit.py tests/test_gpu_golden.py -x -q -m gpu 2>&1 | tail -5
