#!/bin/bash
# round 2, visit s (1 GPU): fold attention kernel, second version (row sum as column 64 of the P V MMA, first block in one pass)
mkdir -p gpurun_out
export ATTN_AB_VARIANTS="1:0,0:2,3:0,3:3,3:4"
( timeout 400 python tools/attn_ab.py check 2>&1 | tail -150 ) > gpurun_out/r2s_attn_check.log; grep -c " ok" gpurun_out/r2s_attn_check.log; grep -E "FAIL|rror" gpurun_out/r2s_attn_check.log | head -20
export ATTN_AB_VARIANTS="0:2,3:0,3:2,3:3,3:4"
( timeout 300 python tools/attn_ab.py time 64 2>&1 | tail -12 ) > gpurun_out/r2s_attn_time64.log; tail -5 gpurun_out/r2s_attn_time64.log
( timeout 300 python tools/attn_ab.py time 256 2>&1 | tail -12 ) > gpurun_out/r2s_attn_time256.log; tail -5 gpurun_out/r2s_attn_time256.log
( PG_ATTN_VARIANT=fold timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2s_bench_fold_stderr.log | tail -1 ) > gpurun_out/r2s_bench_n1_fold.json; python -c "
import json;d=json.load(open('gpurun_out/r2s_bench_n1_fold.json'));print('infer n1 fold:',d['value'],d['ms_per_step'],d['parity_check'],d.get('family_ms_per_step'))"; tail -2 gpurun_out/r2s_bench_fold_stderr.log
NCU="ncu --set full --clock-control none --import-source on"
PG_ATTN_VARIANT=fold timeout 600 $NCU -k regex:attention_fold -s 2 -c 1 -o gpurun_out/r2s_prof_attn_fold -f python tools/ncu_target.py 64 2 > gpurun_out/r2s_ncu_attn.log 2>&1; tail -1 gpurun_out/r2s_ncu_attn.log
python tools/ncu_summary.py gpurun_out/r2s_prof_attn_fold.ncu-rep --stalls --sass 60 > gpurun_out/r2s_attn_fold_ncu_summary.txt 2>&1; head -20 gpurun_out/r2s_attn_fold_ncu_summary.txt
ncu -i gpurun_out/r2s_prof_attn_fold.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r2s_attn_fold_source.csv.gz
du -sh gpurun_out
