"""Small fixed workload for `ncu --set full` captures: one ViT-L/14-336 forward over 64 views (each kernel family
appears with its production shape: M = 64*577 rows)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import synthetic  # noqa: E402
from pigeon_b200.vit_engine import VitDims, VitEngine  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dims = VitDims(layers=layers)
eng = VitEngine(synthetic.random_vit_state_dict(dims, 0), dims, device="cuda:0", max_views_per_pass=views)
px = torch.randn(views, 3, 336, 336, device="cuda:0").half()
for _ in range(2):
    eng.forward(px)
torch.cuda.synchronize()
print("done")
