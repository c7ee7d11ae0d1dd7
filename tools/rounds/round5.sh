#!/bin/bash
mkdir -p gpurun_out
PG_ATTN_VARIANT=32 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_vit.py -x -q -m gpu -k "attention or vit" 2>&1 | tail -4
timeout 300 python tools/attn_ab.py 128 2>&1 | tail -6
cat > /tmp/ref_ncu.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
sys.argv=[sys.argv[0]]
import importlib.util
spec = importlib.util.spec_from_file_location("rb", "tools/refiner_bench.py"); rb = importlib.util.module_from_spec(spec); spec.loader.exec_module(rb)
from pigeon_b200 import ops
dev = rb.dev
C,P,D,B,k = 2076, 1_000_000, 768, 8192, 5
bank, sizes = rb.make_bank(C,P,D)
rng = np.random.default_rng(3)
cand = np.stack([rng.choice(C, size=k, replace=False) for _ in range(B)]).astype(np.int64)
probs = -np.sort(-rng.dirichlet(np.ones(k), size=B), axis=1).astype(np.float32)
emb = torch.randn(B,1,D,device=dev)*0.3; init = torch.rand(B,2,device=dev,dtype=torch.float64)*90
for _ in range(3): ops.refiner_forward(bank, emb, init, torch.from_numpy(cand).to(dev), torch.from_numpy(probs).to(dev), k, 1.6, 1000.0)
torch.cuda.synchronize(); print("done")
PY
timeout 400 ncu --set full --clock-control none --import-source on -k regex:cell_major_scan_kernel -s 1 -c 1 -o gpurun_out/prof_refiner -f python /tmp/ref_ncu.py > gpurun_out/ncu_refiner_stdout.log 2>&1; tail -2 gpurun_out/ncu_refiner_stdout.log
