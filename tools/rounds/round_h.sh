#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
nvidia-smi -L | head -2
timeout 600 python -m pytest tests/test_gpu_train_tower.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -8
cat gpurun_out/parity.log
timeout 600 python tools/train_bench.py --samples 128 --steps 2 --warmup 1 2>&1 | tail -22 | tee gpurun_out/train_bench.log
