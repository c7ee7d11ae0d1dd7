#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train_tower.py tests/test_gpu_train.py -q -m gpu -x 2>&1 | tail -12
