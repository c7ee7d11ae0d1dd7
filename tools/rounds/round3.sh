#!/bin/bash
mkdir -p gpurun_out
PG_TEST_TIMEOUT=200 timeout 600 python tools/run_gpu_tests_isolated.py tests/test_gpu_kernels.py -k "attention" > gpurun_out/attn_tests.log 2>&1
grep -E "^(PASS|FAIL)|passed|failed|Error|error" gpurun_out/attn_tests.log | tail -20
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 300 python tools/microbench.py 128 > gpurun_out/microbench_2cta.log 2>&1; tail -8 gpurun_out/microbench_2cta.log; cp gpurun_out/microbench.json gpurun_out/microbench_2cta.json
PG_GEMM_1CTA=1 timeout 300 python tools/microbench.py 128 > gpurun_out/microbench_1cta.log 2>&1; tail -8 gpurun_out/microbench_1cta.log; cp gpurun_out/microbench.json gpurun_out/microbench_1cta.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm2_f16_kernel -s 1 -c 3 -o gpurun_out/prof_gemm2 -f python tools/ncu_target.py 64 1 > gpurun_out/ncu_gemm2_stdout.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 1 -c 1 -o gpurun_out/prof_attn2 -f python tools/ncu_target.py 64 1 > gpurun_out/ncu_attn2_stdout.log 2>&1
ls -la gpurun_out/*.ncu-rep
