#!/bin/bash
# round 2, visit k (2 GPUs): multi-rank parity on NCCL + the three bench workloads at N = 2
mkdir -p gpurun_out
nvidia-smi -L | head -4
( timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_train_tower.py -x -q -m gpu -k "multi_rank or data_parallel or loop_golden or reference_loop" 2>&1 | tail -15 ) > gpurun_out/r2k_pytest_2gpu.log; tail -8 gpurun_out/r2k_pytest_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
( timeout 600 $TR tools/ddp_infer_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -40 ) > gpurun_out/r2k_ddp_infer_check.log; cat gpurun_out/r2k_ddp_infer_check.log
( timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r2k_bench_n2_stderr.log | tail -1 ) > gpurun_out/r2k_bench_n2.json; python -c "
import json;d=json.load(open('gpurun_out/r2k_bench_n2.json'));print('infer n2:',d['value'],d['ms_per_step'],d['ms_per_step_by_rank'],d['e2e']['value'],d['parity_check'])"; tail -3 gpurun_out/r2k_bench_n2_stderr.log
( timeout 900 $TR bench.py --gpus 2 --workload refiner --steps 5 --warmup 3 2> gpurun_out/r2k_refiner_n2_stderr.log | tail -1 ) > gpurun_out/r2k_refiner_n2.json; python -c "
import json;d=json.load(open('gpurun_out/r2k_refiner_n2.json'));print('refiner n2:',d['value'],d['ms_per_step'],d['ms_per_step_by_rank'],d['roofline']['frac'],d['parity_check'],d['family_ms_per_step'])"; tail -3 gpurun_out/r2k_refiner_n2_stderr.log
( timeout 900 $TR bench.py --gpus 2 --workload train --steps 2 --warmup 1 2> gpurun_out/r2k_train_n2_stderr.log | tail -1 ) > gpurun_out/r2k_train_n2.json; python -c "
import json;d=json.load(open('gpurun_out/r2k_train_n2.json'));print('train n2:',d['value'],d['ms_per_step'],d['ms_per_step_by_rank'])"; tail -3 gpurun_out/r2k_train_n2_stderr.log
( timeout 900 $TR bench.py --gpus 2 --workload train --all-trainable --steps 2 --warmup 1 2> gpurun_out/r2k_train_all_n2_stderr.log | tail -1 ) > gpurun_out/r2k_train_all_n2.json; python -c "
import json;d=json.load(open('gpurun_out/r2k_train_all_n2.json'));print('train all n2:',d['value'],d['ms_per_step'],d['ms_per_step_by_rank'],d['config']['grad_allreduce_bytes_per_step'])"; tail -3 gpurun_out/r2k_train_all_n2_stderr.log
