#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -1
timeout 300 python -m pytest tests/test_gpu_train_tower.py -q -m gpu -x -k "attention_backward or tower_backward" 2>&1 | tail -4
for v in 64 32 8w; do
  echo "== PG_ATTN_BWD=$v"
  PG_ATTN_BWD=$v timeout 300 python tools/train_bench.py --samples 128 --steps 2 --warmup 1 --out gpurun_out/train_bench_bwd_$v.json 2>&1 | grep -E "ms_per_step|attention_bwd" | cut -c150-330
done
