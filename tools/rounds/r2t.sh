#!/bin/bash
# round 2, visit t (1 GPU): fold attention kernel with two softmax threads per row (poly 10+p) against one thread per row
mkdir -p gpurun_out
export ATTN_AB_VARIANTS="0:2,3:2,3:10,3:12,3:14"
( timeout 400 python tools/attn_ab.py check 2>&1 | tail -150 ) > gpurun_out/r2t_attn_check.log; grep -c " ok" gpurun_out/r2t_attn_check.log; grep -E "FAIL|rror" gpurun_out/r2t_attn_check.log | head -20
export ATTN_AB_VARIANTS="0:2,3:2,3:10,3:12,3:13,3:14"
( timeout 300 python tools/attn_ab.py time 64 2>&1 | tail -12 ) > gpurun_out/r2t_attn_time64.log; tail -6 gpurun_out/r2t_attn_time64.log
( timeout 300 python tools/attn_ab.py time 256 2>&1 | tail -12 ) > gpurun_out/r2t_attn_time256.log; tail -6 gpurun_out/r2t_attn_time256.log
NCU="ncu --set full --clock-control none --import-source on"
ATTN_AB_VARIANTS="3:12" timeout 600 $NCU -k regex:attention_fold -s 6 -c 1 -o gpurun_out/r2t_prof_attn_fold16 -f python tools/attn_ab.py time 64 > gpurun_out/r2t_ncu_attn.log 2>&1; tail -1 gpurun_out/r2t_ncu_attn.log
python tools/ncu_summary.py gpurun_out/r2t_prof_attn_fold16.ncu-rep --stalls --sass 40 > gpurun_out/r2t_attn_fold16_ncu_summary.txt 2>&1; head -20 gpurun_out/r2t_attn_fold16_ncu_summary.txt
ncu -i gpurun_out/r2t_prof_attn_fold16.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/r2t_attn_fold16_source.csv.gz
du -sh gpurun_out
