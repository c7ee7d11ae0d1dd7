#!/bin/bash
# round 2, visit ah (2 GPUs): the two multi-rank GPU tests, bench lines at N = 2 (inference, refiner with the slab scan on a
# cell-sharded bank)
mkdir -p gpurun_out
rm -f gpurun_out/ddp_infer_check_rank*.log
( timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_train_tower.py -x -q -m gpu -k "multi_rank or two_rank or nccl or ddp or two_gpu or averag" 2>&1 | tail -5 ) > gpurun_out/r2ah_pytest_2gpu.log; tail -3 gpurun_out/r2ah_pytest_2gpu.log
cat gpurun_out/ddp_infer_check_rank0.log 2>/dev/null | cut -c1-260
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721"
( timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-oracle-check 2> gpurun_out/r2ah_bench_n2_stderr.log | tail -1 ) > gpurun_out/r2ah_bench_n2.json; python -c "
import json;d=json.load(open('gpurun_out/r2ah_bench_n2.json'));print('infer n2:',d['value'],d['ms_per_step'],[round(x,1) for x in d['ms_per_step_by_rank']],d['e2e']['value'],d['parity_check'],d['clocks'])"; tail -2 gpurun_out/r2ah_bench_n2_stderr.log
( timeout 600 $TR bench.py --gpus 2 --workload refiner --steps 5 --warmup 3 2> gpurun_out/r2ah_refiner_n2_stderr.log | tail -1 ) > gpurun_out/r2ah_refiner_n2.json; python -c "
import json;d=json.load(open('gpurun_out/r2ah_refiner_n2.json'));print('refiner n2:',d['value'],d['ms_per_step'],d['parity_check'],d['roofline']['kernel'],d['roofline']['frac'],d['family_ms_per_step'])"; tail -2 gpurun_out/r2ah_refiner_n2_stderr.log
