#!/bin/bash
# round 2, visit a: first light of the pair attention kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.log 2>&1
timeout 300 python tools/attn_ab.py check > gpurun_out/r2a_attn_check.log 2>&1; echo "check rc=$?"; tail -70 gpurun_out/r2a_attn_check.log
if grep -q FAIL gpurun_out/r2a_attn_check.log || ! grep -q "growing logits variant 0 poly 4" gpurun_out/r2a_attn_check.log; then echo "ATTN CHECK FAILED"; exit 0; fi
( timeout 300 python tools/attn_ab.py time 128 2>&1 | tail -20 ) > gpurun_out/r2a_attn_time.log; cat gpurun_out/r2a_attn_time.log
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r2a_pytest_gpu.log; tail -6 gpurun_out/r2a_pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2a_bench_stderr.log | tail -1 ) > gpurun_out/r2a_bench.json; cat gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench_stderr.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_pair_kernel -s 2 -c 1 -o gpurun_out/r2a_prof_attn -f \
    python tools/ncu_target.py 64 2 > gpurun_out/r2a_ncu_attn_stdout.log 2>&1; tail -2 gpurun_out/r2a_ncu_attn_stdout.log
ls -la gpurun_out/*.ncu-rep
