#!/bin/bash
# round 2, visit q (1 GPU): L2 eviction hints in the GEMMs (tests, in-step kernel times, fc2 DRAM bytes at the bench shape);
# stall-reason / SASS hot-spot capture of the pair attention kernel at 64 views.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -x -q -m gpu -k "test_gemm or test_loss" 2>&1 | tail -3 ) > gpurun_out/r2q_pytest.log; tail -2 gpurun_out/r2q_pytest.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2q_bench_stderr.log | tail -1 ) > gpurun_out/r2q_bench_n1.json; python -c "
import json;d=json.load(open('gpurun_out/r2q_bench_n1.json'));print('infer n1:',d['value'],d['ms_per_step'],d['e2e']['value'],d['parity_check'],d.get('family_ms_per_step'))"; tail -2 gpurun_out/r2q_bench_stderr.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:gemm2_f16_kernel -s 5 -c 5 -o gpurun_out/r2q_prof_gemm_bench -f python tools/ncu_target.py 1024 1 > gpurun_out/r2q_ncu_gemm_bench.log 2>&1; tail -1 gpurun_out/r2q_ncu_gemm_bench.log
python tools/ncu_summary.py gpurun_out/r2q_prof_gemm_bench.ncu-rep > gpurun_out/r2q_gemm_bench_ncu_summary.txt 2>&1; rm -f gpurun_out/r2q_prof_gemm_bench.ncu-rep
grep -E "^== launch|duration|dram__bytes" gpurun_out/r2q_gemm_bench_ncu_summary.txt
timeout 900 $NCU -k regex:attention_pair -s 2 -c 1 -o gpurun_out/r2q_prof_attn -f python tools/ncu_target.py 64 2 > gpurun_out/r2q_ncu_attn.log 2>&1; tail -1 gpurun_out/r2q_ncu_attn.log
python tools/ncu_summary.py gpurun_out/r2q_prof_attn.ncu-rep --stalls --sass 80 > gpurun_out/r2q_attn_ncu_summary.txt 2>&1
ncu -i gpurun_out/r2q_prof_attn.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r2q_attn_source.csv.gz
ls -la gpurun_out; du -sh gpurun_out
