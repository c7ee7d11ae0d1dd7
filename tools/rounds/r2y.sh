#!/bin/bash
# round 2, visit y (1 GPU): fused head kernel — parity tests, timing against the three-kernel sequence; LDS micro-benchmark
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -x -q -m gpu -k "head or superguessr or golden" 2>&1 | tail -15 ) > gpurun_out/r2y_pytest_head.log; tail -15 gpurun_out/r2y_pytest_head.log
for f in 1 0; do PG_HEAD_FUSED=$f timeout 120 python tools/head_time.py 2>&1 | tail -1; done | tee gpurun_out/r2y_head_time.log
HB=2048 timeout 120 python tools/head_time.py 2>&1 | tail -1 | tee -a gpurun_out/r2y_head_time.log
HB=2048 PG_HEAD_FUSED=0 timeout 120 python tools/head_time.py 2>&1 | tail -1 | tee -a gpurun_out/r2y_head_time.log
timeout 100 tools/ubench/fp32_lds > gpurun_out/r2y_ubench_fp32_lds.txt 2>&1; grep LDS gpurun_out/r2y_ubench_fp32_lds.txt
