#!/bin/bash
# round 2, visit u (1 GPU): the driver's round-end sequence (GPU tests, smoke, both bench arms) + the N = 1 lines of the other
# workloads, the GEMM capture at the bench shape after the residual prefetch, and the launch list of one bench step.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r2u_pytest_gpu.log; tail -3 gpurun_out/r2u_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 ) > gpurun_out/r2u_smoke.log; tail -1 gpurun_out/r2u_smoke.log
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r2u_bench_reference_stderr.log | tail -1 ) > gpurun_out/r2u_bench_reference.json; cut -c1-300 gpurun_out/r2u_bench_reference.json
( timeout 900 python bench.py 2> gpurun_out/r2u_bench_stderr.log | tail -1 ) > gpurun_out/r2u_bench_n1.json; python -c "
import json;d=json.load(open('gpurun_out/r2u_bench_n1.json'));print('infer n1:',d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check']['ok'],d['roofline']['frac'],d.get('roofline_attention',{}).get('frac'),d['clocks'],d.get('family_ms_per_step'))"; tail -2 gpurun_out/r2u_bench_stderr.log
( timeout 600 python bench.py --workload refiner --steps 5 --warmup 3 2> gpurun_out/r2u_refiner_stderr.log | tail -1 ) > gpurun_out/r2u_refiner_k5.json; python -c "
import json;d=json.load(open('gpurun_out/r2u_refiner_k5.json'));print('refiner:',d['value'],d['ms_per_step'],d['roofline']['frac'])"
( timeout 600 python bench.py --workload train --steps 2 --warmup 1 2> gpurun_out/r2u_train_stderr.log | tail -1 ) > gpurun_out/r2u_train.json; python -c "
import json;d=json.load(open('gpurun_out/r2u_train.json'));print('train:',d['value'],d['ms_per_step'])"; tail -2 gpurun_out/r2u_train_stderr.log
( timeout 600 python bench.py --workload train --all-trainable --steps 2 --warmup 1 2> gpurun_out/r2u_train_all_stderr.log | tail -1 ) > gpurun_out/r2u_train_all.json; python -c "
import json;d=json.load(open('gpurun_out/r2u_train_all.json'));print('train all:',d['value'],d['ms_per_step'])"; tail -2 gpurun_out/r2u_train_all_stderr.log
NCU="ncu --set full --clock-control none"
timeout 900 $NCU -k regex:gemm2_f16_kernel -s 5 -c 5 -o gpurun_out/r2u_prof_gemm_bench -f python tools/ncu_target.py 1024 1 > gpurun_out/r2u_ncu_gemm_bench.log 2>&1; tail -1 gpurun_out/r2u_ncu_gemm_bench.log
python tools/ncu_summary.py gpurun_out/r2u_prof_gemm_bench.ncu-rep > gpurun_out/r2u_gemm_bench_ncu_summary.txt 2>&1; rm -f gpurun_out/r2u_prof_gemm_bench.ncu-rep
grep -E "^== launch|duration|dram__bytes|tensor_cycles_active.avg.pct_of_peak_sustained_active" gpurun_out/r2u_gemm_bench_ncu_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kernel" -c 400 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-oracle-check > gpurun_out/r2u_ncu_launches_stdout.log 2>&1; wc -l gpurun_out/r2u_launches.csv
du -sh gpurun_out
