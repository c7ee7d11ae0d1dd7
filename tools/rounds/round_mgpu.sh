#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_golden.py -x -q -m gpu -k "refiner or protorefiner" 2>&1 | tail -4
timeout 600 python tools/refiner_bench.py > gpurun_out/refiner_bench2.log 2>&1; grep -E "^\{" gpurun_out/refiner_bench2.log | cut -c1-330
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 4 --warmup 3 2> gpurun_out/bench_n${N}_stderr.log | tail -1 ) > gpurun_out/bench_n$N.json; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n${N}_stderr.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 ) > gpurun_out/bench_ref_n$N.json; cat gpurun_out/bench_ref_n$N.json
