"""Small fixed workload for `ncu --set full` captures of the backward kernels: one training forward + backward of a
2-layer ViT-L/14-336-shaped tower over 64 views (every kernel appears with its production shape, M = 64*577 rows)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import CLIPVisionTower, VitDims, synthetic  # noqa: E402
from pigeon_b200.vit_train import TowerTrainer  # noqa: E402

views = int(sys.argv[1]) if len(sys.argv) > 1 else 64
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dims = VitDims(layers=layers)
tower = CLIPVisionTower(dims)
tower.load_state_dict(synthetic.random_vit_state_dict(dims, 0), strict=True)
tower.to("cuda:0")
tr = TowerTrainer(tower, max_views=views)
px = torch.randn(views, 3, 336, 336, device="cuda:0").half()
d_emb = torch.randn(views, dims.hidden, device="cuda:0") * 1e-4
for _ in range(2):
    tr.forward(px)
    tr.backward(d_emb)
torch.cuda.synchronize()
print("done")
