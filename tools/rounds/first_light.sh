#!/bin/bash
# First GPU contact: isolated kernel tests, then micro-benchmarks.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
PG_TEST_TIMEOUT=300 timeout 1500 python tools/run_gpu_tests_isolated.py tests > gpurun_out/isolated_stdout.log 2>&1
echo "tests exit $?" >> gpurun_out/isolated_stdout.log
tail -60 gpurun_out/isolated_stdout.log
timeout 300 python tools/microbench.py 128 > gpurun_out/microbench.log 2>&1
echo "microbench exit $?" >> gpurun_out/microbench.log
tail -20 gpurun_out/microbench.log
