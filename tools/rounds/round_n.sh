#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on \
    -k 'regex:gemm2_f16_kernel<\(int\)5>|gemm2_f16_kernel<\(int\)3>|ln_backward_kernel|attention_delta_kernel|cast_kernel|transpose_kernel' \
    -s 12 -c 14 -o gpurun_out/prof_train_misc_r01h -f python tools/ncu_train_target.py 64 2 > gpurun_out/ncu_train_misc_stdout.log 2>&1
tail -2 gpurun_out/ncu_train_misc_stdout.log
cat > /tmp/pp.py <<'PY'
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from pigeon_b200 import synthetic
from pigeon_b200.preprocess import ClipImageProcessor
imgs = [torch.from_numpy(synthetic.synthetic_photo(480, 640, seed=i)).cuda() for i in range(256)]
proc = ClipImageProcessor(dtype=torch.float16)
for _ in range(3):
    out = proc.preprocess_device(imgs)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    out = proc.preprocess_device(imgs)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
raw = sum(i.numel() for i in imgs)
print(f"preprocess 256 x 640x480 uint8 -> fp16 [256,3,336,336]: {ms:.3f} ms = {256 / ms * 1000:.0f} images/s, "
      f"{(raw + out.numel() * 2) / ms / 1e6:.1f} GB/s of input+output bytes")
PY
timeout 300 python /tmp/pp.py 2>&1 | tail -2 | tee gpurun_out/preprocess_bench.log
ls -la gpurun_out/prof_train_misc_r01h.ncu-rep
