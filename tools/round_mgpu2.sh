#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 4 --warmup 3 2> gpurun_out/bench_n${N}_stderr.log | tail -1 ) > gpurun_out/bench_n$N.json; cat gpurun_out/bench_n$N.json | cut -c1-1500; tail -3 gpurun_out/bench_n${N}_stderr.log
