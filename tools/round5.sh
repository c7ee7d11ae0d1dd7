#!/bin/bash
mkdir -p gpurun_out
PG_ATTN_VARIANT=32 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vit.py -x -q -m gpu -k "attention or vit" 2>&1 | tail -4
timeout 300 python tools/attn_ab.py 128 2>&1 | tail -6
