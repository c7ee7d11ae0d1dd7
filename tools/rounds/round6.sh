#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python tools/refiner_bench.py > gpurun_out/refiner_bench3.log 2>&1; grep -E "^\{" gpurun_out/refiner_bench3.log | cut -c1-330
timeout 300 python tools/microbench.py 128 2>&1 | tail -3
