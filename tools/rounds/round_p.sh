#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/mn_probe.py 2>&1 | tail -5
timeout 400 python -m pytest tests/test_gpu_train_tower.py -q -m gpu -x 2>&1 | tail -4
timeout 400 python tools/train_bench.py --samples 128 --steps 2 --warmup 1 --all-trainable --out gpurun_out/train_bench_all_v3.json 2>&1 | grep -E "ms_per_step|rror" | cut -c150-330
timeout 400 python tools/train_bench.py --samples 128 --steps 2 --warmup 1 --out gpurun_out/train_bench_v6.json 2>&1 | grep -E "ms_per_step|rror" | cut -c150-330
