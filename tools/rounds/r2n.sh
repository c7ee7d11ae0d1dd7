#!/bin/bash
# round 2, visit n (1 GPU): ncu evidence for every kernel family + compute-sanitizer runs
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none"      # (no --import-source: gpurun_out/ is capped at 64 MiB)
# 1. inference tower at 64 views, 2 layers: every forward kernel once (im2col, patch GEMM, embed_preln, LN-folded GEMMs, pair attention, token_mean)
timeout 900 $NCU -s 17 -c 17 -o gpurun_out/r2n_prof_tower -f python tools/ncu_target.py 64 2 > gpurun_out/r2n_ncu_tower.log 2>&1; tail -1 gpurun_out/r2n_ncu_tower.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_pair_kernel -s 2 -c 1 -o gpurun_out/r2n_prof_attn -f python tools/ncu_target.py 64 2 > gpurun_out/r2n_ncu_attn.log 2>&1; tail -1 gpurun_out/r2n_ncu_attn.log
# 2. GEMMs at the BENCH shape (M = 1024 * 577 rows): DRAM traffic per launch for roofline.traffic
timeout 900 $NCU -k regex:gemm2_f16_kernel -s 4 -c 4 -o gpurun_out/r2n_prof_gemm_bench -f python tools/ncu_target.py 1024 1 > gpurun_out/r2n_ncu_gemm_bench.log 2>&1; tail -1 gpurun_out/r2n_ncu_gemm_bench.log
# 3. head + refiner at configs[2] scale, refiner at configs[4] scale, pre-processing, AdamW
timeout 900 $NCU -k regex:"view_mean|softmax_topk|ce_loss|pool_views|pair_hist|cell_scan|pair_scatter|scan_kernel|finalize|gemm_f16_kernel" -s 10 -c 12 -o gpurun_out/r2n_prof_head -f python tools/ncu_misc_target.py head > gpurun_out/r2n_ncu_head.log 2>&1; tail -1 gpurun_out/r2n_ncu_head.log
timeout 900 $NCU -k regex:"cell_major_scan" -s 1 -c 1 -o gpurun_out/r2n_prof_refiner -f python tools/ncu_misc_target.py refiner > gpurun_out/r2n_ncu_refiner.log 2>&1; tail -1 gpurun_out/r2n_ncu_refiner.log
timeout 900 $NCU -k regex:"coeff_kernel|horizontal_kernel|vertical_kernel" -s 3 -c 3 -o gpurun_out/r2n_prof_preprocess -f python tools/ncu_misc_target.py preprocess > gpurun_out/r2n_ncu_preprocess.log 2>&1; tail -1 gpurun_out/r2n_ncu_preprocess.log
timeout 900 $NCU -k regex:"adamw" -s 1 -c 1 -o gpurun_out/r2n_prof_adamw -f python tools/ncu_misc_target.py adamw > gpurun_out/r2n_ncu_adamw.log 2>&1; tail -1 gpurun_out/r2n_ncu_adamw.log
for n in tower gemm_bench head refiner preprocess adamw attn; do python tools/ncu_summary.py gpurun_out/r2n_prof_$n.ncu-rep > gpurun_out/r2n_${n}_ncu_summary.txt 2>&1; done
rm -f gpurun_out/r2n_prof_tower.ncu-rep gpurun_out/r2n_prof_gemm_bench.ncu-rep gpurun_out/r2n_prof_head.ncu-rep gpurun_out/r2n_prof_preprocess.ncu-rep gpurun_out/r2n_prof_adamw.ncu-rep
wc -l gpurun_out/r2n_*_ncu_summary.txt; du -sh gpurun_out
