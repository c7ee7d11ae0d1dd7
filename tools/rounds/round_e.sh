#!/bin/bash
# round e: new training / pre-processing tests, attention poly A/B, chunk sweep
mkdir -p gpurun_out
nvidia-smi -L | head -2
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_preprocess.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_new.log
timeout 200 python tools/attn_ab.py 128 2>&1 | tail -10 | tee gpurun_out/attn_ab.log
timeout 400 python tools/chunk_sweep.py 1024 2>&1 | tail -12 | tee gpurun_out/chunk_sweep.log
