"""CPU-only: the C-ABI library loads and exports every symbol the header declares; host-side logic
(bank packing, sharding, packed all-gather over gloo with world_size 2, module surface)."""
import os
import re
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pigeon_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pigeon_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 18
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/pigeon_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.pg_abi_version() == 1


def test_no_gpu_fails_loudly():
    """On a box without CUDA the product path must raise, never fall back to CPU arithmetic."""
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from pigeon_b200 import PigeonB200Error, SuperGuessr, ops
    with pytest.raises(PigeonB200Error):
        ops.layernorm_f16(torch.zeros(4, 128), torch.ones(128), torch.zeros(128), 1e-5)
    sg = SuperGuessr(None, panorama=True, geocells=np.zeros((8, 2))).eval()
    with pytest.raises(PigeonB200Error):
        sg(embedding=torch.zeros(2, 4, 1024), labels_clf=torch.zeros(2, dtype=torch.long))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pigeon_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_shard_range_partitions():
    from pigeon_b200.dist import shard_range
    for total in (0, 1, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bank_from_arrays_matches_manual_means():
    from pigeon_b200 import bank as bank_mod
    rng = np.random.default_rng(0)
    n, D = 50, 128
    data = rng.standard_normal((n, 4, D)).astype(np.float32)
    ll = rng.uniform(-90, 90, (n, 2)).astype(np.float32)
    cells = np.array([3, 0, 3, 1, 5])                       # unsorted rows, cell 2 and 4 empty, 6 cells total
    idx = [[1, 2, 3], [4], [10, 11], [], [20, 21, 22, 23]]   # row 3 has no members -> dropped
    b = bank_mod.bank_from_arrays(cells, rng.uniform(-90, 90, (5, 2)), idx, torch.from_numpy(data), ll, num_cells=6)
    assert b["cell_off"].tolist() == [0, 1, 1, 1, 3, 3, 4]
    assert b["proto_count"].tolist() == [1, 3, 2, 4]
    assert b["member_idx"].tolist() == [4, 1, 2, 3, 10, 11, 20, 21, 22, 23]
    m = torch.from_numpy(data).mean(1)
    np.testing.assert_allclose(b["proto_emb"][1], m[[1, 2, 3]].mean(0).numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(b["data_emb"], m.numpy())


def test_module_surface_matches_reference_signatures():
    import inspect
    from pigeon_b200 import CLIPEmbedding, ModelOutput, ProtoRefiner, SuperGuessr
    sg = list(inspect.signature(SuperGuessr.__init__).parameters)[1:12]
    assert sg == ["base_model", "panorama", "hierarchical", "should_smooth_labels", "multi_task", "heading", "yfcc",
                  "serving", "freeze_base", "num_candidates", "embed_dim"]                     # super_guessr.py:31-34
    fw = list(inspect.signature(SuperGuessr.forward).parameters)[1:]
    assert fw == ["pixel_values", "embedding", "heading", "labels", "labels_clf", "labels_multi_task", "labels_climate",
                  "labels_month", "index"]                                                      # :350-353
    pr = list(inspect.signature(ProtoRefiner.__init__).parameters)[1:9]
    assert pr == ["topk", "hedge", "max_refinement", "temperature", "proto_path", "dataset_path", "protos", "verbose"]
    pf = list(inspect.signature(ProtoRefiner.forward).parameters)[1:7]
    assert pf == ["embedding", "geo_tensor", "initial_preds", "candidate_cells", "candidate_probs", "cluster"]
    ce = list(inspect.signature(CLIPEmbedding.__init__).parameters)[1:5]
    assert ce == ["model_name", "device", "load_checkpoint", "panorama"]                        # clip_embedder.py:11-12
    assert ModelOutput._fields == ("loss", "loss_clf", "loss_reg", "loss_climate", "loss_month", "preds_LLH",
                                   "preds_geocell", "preds_mt", "preds_climate", "preds_month", "top5_geocells",
                                   "embedding")                                                 # models/utils.py:7-9


def test_tower_state_dict_uses_hf_key_names_and_roundtrips():
    from pigeon_b200 import CLIPVisionTower, VitDims, synthetic
    dims = VitDims(image_size=56, hidden=256, heads=4, intermediate=512, layers=2)
    sd = synthetic.random_vit_state_dict(dims, seed=3)
    t = CLIPVisionTower(dims)
    t.load_state_dict(sd)
    assert set(t.state_dict()) == set(sd)
    assert "vision_model.pre_layrnorm.weight" in sd and "vision_model.encoder.layers.1.self_attn.q_proj.weight" in sd
    assert len(list(t.vision_model.encoder.layers[:-1].parameters())) == 16                     # freeze policy handle


GLOO_WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from pigeon_b200 import dist as pdist
    rank, world = int(sys.argv[1]), 2
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[2], RANK=str(rank), WORLD_SIZE="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 3
    pack = dict(emb=torch.full((B, 5), float(rank)) + torch.arange(5), idx=torch.arange(B * 2).reshape(B, 2) + 100 * rank,
                val=torch.rand(B, 2).double(), init=torch.full((B, 2), rank + 0.5, dtype=torch.float64))
    out = pdist.all_gather_rows(pack)
    assert out["emb"].shape == (2 * B, 5) and out["emb"].dtype == torch.float32
    assert torch.equal(out["emb"][:B], torch.zeros(B, 5) + torch.arange(5)) and torch.equal(out["emb"][B:], torch.ones(B, 5) + torch.arange(5))
    assert out["idx"].dtype == torch.int64 and out["idx"][B:].min() == 100 and out["idx"][:B].max() == 5
    assert out["init"].dtype == torch.float64 and out["init"][0, 0] == 0.5 and out["init"][-1, 0] == 1.5
    assert torch.equal(out["val"][rank * B:(rank + 1) * B], pack["val"])
    lo, hi = pdist.shard_range(7, rank, world)
    assert (lo, hi) == ((0, 4) if rank == 0 else (4, 7))
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_packed_all_gather_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER % ROOT)
    port = str(29000 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), port], stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0].decode() for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
