#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
nvidia-smi -L | head -1
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu.log
cat gpurun_out/parity.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_bwd_kernel -s 2 -c 2 -o gpurun_out/prof_attn_bwd_r01g -f \
    python tools/ncu_train_target.py 64 2 > gpurun_out/ncu_attn_bwd_stdout.log 2>&1
tail -2 gpurun_out/ncu_attn_bwd_stdout.log
ls -la gpurun_out/prof_attn_bwd_r01g.ncu-rep
