#!/bin/bash
# round 2, visit o (1 GPU): remaining ncu evidence (tower kernels by name, all five GEMM shapes at the bench size), launch list of
# one bench step, compute-sanitizer logs.  Reports are summarised on the box (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none"
timeout 900 $NCU -k regex:"im2col|gemm|embed_preln|attention|token_mean" -s 14 -c 14 -o gpurun_out/r2o_prof_tower -f python tools/ncu_target.py 64 2 > gpurun_out/r2o_ncu_tower.log 2>&1; tail -1 gpurun_out/r2o_ncu_tower.log
python tools/ncu_summary.py gpurun_out/r2o_prof_tower.ncu-rep > gpurun_out/r2o_tower_ncu_summary.txt 2>&1; rm -f gpurun_out/r2o_prof_tower.ncu-rep
timeout 900 $NCU -k regex:gemm2_f16_kernel -s 5 -c 5 -o gpurun_out/r2o_prof_gemm_bench -f python tools/ncu_target.py 1024 1 > gpurun_out/r2o_ncu_gemm_bench.log 2>&1; tail -1 gpurun_out/r2o_ncu_gemm_bench.log
python tools/ncu_summary.py gpurun_out/r2o_prof_gemm_bench.ncu-rep > gpurun_out/r2o_gemm_bench_ncu_summary.txt 2>&1; rm -f gpurun_out/r2o_prof_gemm_bench.ncu-rep
grep -E "^== launch|duration|dram__bytes" gpurun_out/r2o_gemm_bench_ncu_summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pg::" -s 393 -c 140 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-oracle-check > gpurun_out/r2o_ncu_launches_stdout.log 2>&1; wc -l gpurun_out/r2o_launches.csv
K="test_gemm or test_layernorm or test_attention or test_head or test_refiner_cell_major or test_refiner_cell_sharded or (test_refiner_vs_oracle and not 100000)"
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$K" 2>&1 | tail -12 ) > gpurun_out/r2o_sanitizer_memcheck.log; tail -4 gpurun_out/r2o_sanitizer_memcheck.log
( timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "test_layernorm or test_head or test_refiner_cell_major or test_attention" 2>&1 | grep -E "RACECHECK SUMMARY|passed|failed|hazard.*in |Race reported" | sort | uniq -c | sort -rn | head -30 ) > gpurun_out/r2o_sanitizer_racecheck.log; cat gpurun_out/r2o_sanitizer_racecheck.log | head -12
du -sh gpurun_out
