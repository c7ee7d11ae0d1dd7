#!/bin/bash
# One GPU visit: driver-style test run, smoke, bench (both arms), ncu launch list, ncu full captures.
mkdir -p gpurun_out
R=${1:-r01}
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 ) > gpurun_out/smoke.log; cat gpurun_out/smoke.log
( timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_stderr.log | tail -1 ) > gpurun_out/bench_$R.json; cat gpurun_out/bench_$R.json; tail -3 gpurun_out/bench_stderr.log
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 ) > gpurun_out/bench_reference_$R.json; cat gpurun_out/bench_reference_$R.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 548 -c 190 --csv --log-file gpurun_out/launches_$R.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_stdout.log 2>&1
  tail -2 gpurun_out/ncu_launches_stdout.log
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_f16_kernel -s 5 -c 4 -o gpurun_out/prof_gemm_$R -f \
      python tools/ncu_target.py 64 2 > gpurun_out/ncu_gemm_stdout.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 2 -c 1 -o gpurun_out/prof_attn_$R -f \
      python tools/ncu_target.py 64 2 > gpurun_out/ncu_attn_stdout.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
timeout 900 python tools/refiner_bench.py > gpurun_out/refiner_bench.log 2>&1; tail -8 gpurun_out/refiner_bench.log
