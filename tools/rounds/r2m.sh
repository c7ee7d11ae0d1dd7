#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
( timeout 600 $TR tools/ddp_infer_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -40 ) > gpurun_out/r2m_ddp_infer_check.log; cat gpurun_out/ddp_infer_check_rank*.log
( timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_train_tower.py -q -m gpu -k "multi_rank or data_parallel" 2>&1 | tail -15 ) > gpurun_out/r2m_pytest_2gpu.log; tail -4 gpurun_out/r2m_pytest_2gpu.log
