#!/bin/bash
# round 2, visit ai (8 GPUs): BASELINE configs[4] "1 vs 8 GPU": the refiner on a bank sharded by geocell over 8 ranks, slab scan
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731"
( timeout 300 $TR bench.py --gpus 8 --workload refiner --steps 5 --warmup 3 2> gpurun_out/r2ai_refiner_n8_stderr.log | tail -1 ) > gpurun_out/r2ai_refiner_n8.json; python -c "
import json;d=json.load(open('gpurun_out/r2ai_refiner_n8.json'));print('refiner n8:',d['value'],d['ms_per_step'],d['ms_per_step_by_rank'],d['parity_check'],d['roofline']['kernel'],d['family_ms_per_step'])"; tail -2 gpurun_out/r2ai_refiner_n8_stderr.log
