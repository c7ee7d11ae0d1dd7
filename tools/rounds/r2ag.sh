#!/bin/bash
# round 2, visit ag (1 GPU): the driver's round-end sequence on the final tree (GPU tests, smoke, both bench arms) + refiner / train lines
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2ag_pytest_gpu.log; tail -3 gpurun_out/r2ag_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > gpurun_out/r2ag_smoke.log; tail -1 gpurun_out/r2ag_smoke.log
( timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r2ag_bench_reference_stderr.log | tail -1 ) > gpurun_out/r2ag_bench_reference.json; cut -c1-200 gpurun_out/r2ag_bench_reference.json
( timeout 900 python bench.py 2> gpurun_out/r2ag_bench_stderr.log | tail -1 ) > gpurun_out/r2ag_bench_n1.json; python -c "
import json;d=json.load(open('gpurun_out/r2ag_bench_n1.json'));print('infer n1:',d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check']['ok'],d['roofline']['frac'],d.get('roofline_attention',{}).get('frac'),d['clocks'],d['cpu_baseline'],d.get('family_ms_per_step'))"; tail -2 gpurun_out/r2ag_bench_stderr.log
( timeout 300 python bench.py --workload refiner --steps 5 --warmup 3 2> gpurun_out/r2ag_refiner_stderr.log | tail -1 ) > gpurun_out/r2ag_refiner_k5.json; python -c "
import json;d=json.load(open('gpurun_out/r2ag_refiner_k5.json'));print('refiner auto:',d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['kernel'])"
( timeout 600 python bench.py --workload train --steps 2 --warmup 1 2> gpurun_out/r2ag_train_stderr.log | tail -1 ) > gpurun_out/r2ag_train.json; python -c "
import json;d=json.load(open('gpurun_out/r2ag_train.json'));print('train:',d['value'],d['ms_per_step'])"; tail -2 gpurun_out/r2ag_train_stderr.log
