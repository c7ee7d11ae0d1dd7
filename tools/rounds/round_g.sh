#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -1
timeout 300 python -m pytest tests/test_gpu_train_tower.py -q -m gpu -x -k "attention_backward" 2>&1 | tail -5
timeout 600 python tools/train_bench.py --samples 128 --steps 2 --warmup 1 2>&1 | tail -22 | tee gpurun_out/train_bench.log
