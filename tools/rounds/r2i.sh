#!/bin/bash
# round 2, visit i: split attention kernel (16 softmax warps) first light
mkdir -p gpurun_out
timeout 300 python tools/attn_ab.py check > gpurun_out/r2i_attn_check.log 2>&1; echo "check rc=$?"; grep -c " ok" gpurun_out/r2i_attn_check.log; grep -v " ok" gpurun_out/r2i_attn_check.log | tail -20
if grep -q FAIL gpurun_out/r2i_attn_check.log || ! grep -q "growing logits variant 2 poly 4" gpurun_out/r2i_attn_check.log; then echo "ATTN CHECK FAILED"; exit 0; fi
( timeout 300 python tools/attn_ab.py time 128 2>&1 | tail -20 ) > gpurun_out/r2i_attn_time.log; cat gpurun_out/r2i_attn_time.log
PG_ATTN_VARIANT=split timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_split_kernel -s 2 -c 1 -o gpurun_out/r2i_prof_attn -f \
    python tools/ncu_target.py 64 2 > gpurun_out/r2i_ncu_attn_stdout.log 2>&1; tail -2 gpurun_out/r2i_ncu_attn_stdout.log
