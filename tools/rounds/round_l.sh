#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 400 python tools/train_bench.py --samples 128 --steps 3 --warmup 1 --out gpurun_out/train_bench_n1.json 2>&1 | grep -E "ms_per_step" | cut -c150-420
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/train_bench.py --samples 128 --steps 3 --warmup 1 --out gpurun_out/train_bench_n2.json 2>&1 | grep -E "ms_per_step|rror" | cut -c150-460
timeout 400 python tools/train_bench.py --samples 128 --steps 2 --warmup 1 --all-trainable --out gpurun_out/train_bench_all_n1.json 2>&1 | grep -E "ms_per_step|rror" | cut -c1-460
