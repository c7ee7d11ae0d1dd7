#!/bin/bash
# round 2, visit p (8 GPUs): BASELINE.json configs[2], [3], [4] at their real shapes
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_vit.py -x -q -m gpu -k "vit_tiny or test_vit_large_336 or chunking" 2>&1 | tail -3 ) > gpurun_out/r2p_pytest_vit.log; tail -2 gpurun_out/r2p_pytest_vit.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711"
( timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 --no-oracle-check 2> gpurun_out/r2p_bench_n8_stderr.log | tail -1 ) > gpurun_out/r2p_bench_n8.json; python -c "
import json;d=json.load(open('gpurun_out/r2p_bench_n8.json'));print('infer n8:',d['value'],d['ms_per_step'],[round(x,1) for x in d['ms_per_step_by_rank']],d['e2e']['value'],d['parity_check'],d['clocks'])"; tail -2 gpurun_out/r2p_bench_n8_stderr.log
( timeout 600 $TR bench.py --gpus 8 --workload refiner --steps 5 --warmup 3 2> gpurun_out/r2p_refiner_n8_stderr.log | tail -1 ) > gpurun_out/r2p_refiner_n8.json; python -c "
import json;d=json.load(open('gpurun_out/r2p_refiner_n8.json'));print('refiner n8:',d['value'],d['ms_per_step'],[round(x,3) for x in d['ms_per_step_by_rank']],d['roofline']['frac'],d['parity_check'],d['family_ms_per_step'])"; tail -2 gpurun_out/r2p_refiner_n8_stderr.log
( timeout 600 $TR bench.py --gpus 8 --workload train --steps 2 --warmup 1 2> gpurun_out/r2p_train_n8_stderr.log | tail -1 ) > gpurun_out/r2p_train_n8.json; python -c "
import json;d=json.load(open('gpurun_out/r2p_train_n8.json'));print('train n8:',d['value'],d['ms_per_step'],[round(x,1) for x in d['ms_per_step_by_rank']])"; tail -2 gpurun_out/r2p_train_n8_stderr.log
( timeout 600 $TR bench.py --gpus 8 --workload train --all-trainable --steps 2 --warmup 1 2> gpurun_out/r2p_train_all_n8_stderr.log | tail -1 ) > gpurun_out/r2p_train_all_n8.json; python -c "
import json;d=json.load(open('gpurun_out/r2p_train_all_n8.json'));print('train all n8:',d['value'],d['ms_per_step'],[round(x,1) for x in d['ms_per_step_by_rank']],d['config']['grad_allreduce_bytes_per_step'])"; tail -2 gpurun_out/r2p_train_all_n8_stderr.log
