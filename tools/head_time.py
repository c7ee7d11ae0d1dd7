"""Times pg_head_forward at the bench shape (B = 256 four-view samples, D = 1024, C = 2076, top-5) with CUDA events."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pigeon_b200 import ops
ops.head_set_fused(os.environ.get("PG_HEAD_FUSED", "0") == "1")
dev = torch.device("cuda:0")
B, V, D, C, k = int(os.environ.get("HB", 256)), 4, 1024, 2076, 5
g = torch.Generator().manual_seed(0)
emb = (torch.randn(B, V, D, generator=g) * 0.3).to(dev)
lin = torch.nn.Linear(D, C)
W, bias = lin.weight.detach().to(dev), lin.bias.detach().to(dev)
cent = torch.rand(C, 2, dtype=torch.float64).to(dev)
w3 = ops.head_pack_weight(W)
for _ in range(5):
    out = ops.head_forward(emb, w3, bias, cent, k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    out = ops.head_forward(emb, w3, bias, cent, k)
e1.record(); torch.cuda.synchronize()
ref = torch.softmax(emb.mean(1).double() @ W.double().t() + bias.double(), -1)
print(json.dumps({"fused": os.environ.get("PG_HEAD_FUSED", "0"), "B": B, "ms_per_call": e0.elapsed_time(e1) / 50,
                  "probs_rel_err": float((out["probs"].double() - ref).norm() / ref.norm()),
                  "argmax_equal": bool(torch.equal(out["pred_cell"], ref.argmax(-1)))}))
