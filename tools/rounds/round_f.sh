#!/bin/bash
# round f: tower fine-tune kernels, one test at a time under a timeout (a hung kernel must not eat the budget)
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
nvidia-smi -L | head -1
for t in test_gemm_bf16_operands test_transpose_and_dgelu test_layernorm_backward test_attention_backward_matches_autograd \
         test_tower_backward_matches_reference_autograd test_reference_freeze_policy_and_optimizer_step; do
  echo "=== $t"
  timeout 150 python -m pytest tests/test_gpu_train_tower.py -q -m gpu -x -k "$t" 2>&1 | tail -25
done 2>&1 | tee gpurun_out/pytest_tower.log | tail -120
cat gpurun_out/parity.log 2>/dev/null
