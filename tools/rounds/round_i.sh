#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -1
timeout 200 python tools/attn_ab.py 128 2>&1 | tail -12 | tee gpurun_out/attn_ab2.log
timeout 300 python -m pytest tests/test_gpu_train_tower.py -q -m gpu -x -k "layernorm or tower_backward" 2>&1 | tail -4
timeout 600 python tools/train_bench.py --samples 128 --steps 2 --warmup 1 2>&1 | tail -14 | tee gpurun_out/train_bench.log
