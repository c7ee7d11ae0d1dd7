#!/bin/bash
# round 2, visit af (1 GPU): slab scan with producer back-off and the lane-distributed row-piece L2 prefetch (0 / 4 / 8 chunks)
mkdir -p gpurun_out
for pf in 0 4 8; do
( PG_SLAB_PREFETCH=$pf timeout 300 python bench.py --workload refiner --steps 5 --warmup 3 --refiner-schedule 4 2> gpurun_out/r2af_stderr.log | tail -1 ) > gpurun_out/r2af_refiner_k5_pf${pf}.json; python -c "
import json;d=json.load(open('gpurun_out/r2af_refiner_k5_pf${pf}.json'));print('refiner k5 sched 4 prefetch ${pf}:',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['scan_ms'])"; tail -2 gpurun_out/r2af_stderr.log
done
( PG_SLAB_PREFETCH=4 timeout 300 python bench.py --workload refiner --steps 5 --warmup 3 --refiner-schedule 4 --refiner-topk 40 2> gpurun_out/r2af_stderr.log | tail -1 ) > gpurun_out/r2af_refiner_k40_pf4.json; python -c "
import json;d=json.load(open('gpurun_out/r2af_refiner_k40_pf4.json'));print('refiner k40 sched 4 prefetch 4:',d['value'],d['ms_per_step'],d['roofline']['scan_ms'])"
( PG_SLAB_PREFETCH=4 timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tile_scan_equals" 2>&1 | tail -3 )
