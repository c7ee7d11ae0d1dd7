#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 ) > gpurun_out/r2l_pytest_gpu.log; tail -30 gpurun_out/r2l_pytest_gpu.log
