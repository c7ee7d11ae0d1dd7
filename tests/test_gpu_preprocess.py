"""pg_preprocess_clip (GPU CLIPProcessor replacement, N2) against the CPU oracle (oracle/preprocess.py, itself pinned
bit-exactly to Pillow and to the HF PIL-backend processor by tests/golden/preprocess.npz).  Bit-exact: uint8 work."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _sha(a):
    import hashlib
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_preprocess_matches_golden_bit_exact(cuda):
    from pigeon_b200 import synthetic
    from pigeon_b200.preprocess import ClipImageProcessor
    G = np.load(os.path.join(GOLD, "preprocess.npz"))
    imgs = [synthetic.synthetic_photo(int(h), int(w), seed=91 + i) for i, (h, w) in enumerate(G["shapes"])]
    px = ClipImageProcessor(device=cuda)(images=imgs, return_tensors="pt")["pixel_values"]      # one mixed-size batch
    assert px.shape == (len(imgs), 3, 336, 336) and px.dtype == torch.float32 and px.is_cuda
    px = px.cpu().numpy()
    for i in range(len(imgs)):
        assert np.array_equal(px[i][:, ::12, ::7], G[f"px{i}_sample"]), i
        assert np.array_equal(_sha(px[i]), G[f"px{i}_sha"]), i


@pytest.mark.parametrize("h,w", [(336, 336), (337, 1200), (1024, 340), (97, 131), (2000, 3000), (336, 3000)])
def test_preprocess_matches_oracle_edge_shapes(cuda, h, w):
    """no-resize, strong anisotropy, up-scaling, large down-scaling factor (ksize 37), long rows."""
    from oracle import preprocess as op
    from pigeon_b200.preprocess import ClipImageProcessor
    rng = np.random.default_rng(h * 7 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = op.clip_preprocess(img)
    proc = ClipImageProcessor(device=cuda)
    got = proc(images=img)["pixel_values"][0].cpu().numpy()
    assert np.array_equal(got, ref)
    half = ClipImageProcessor(device=cuda, dtype=torch.float16)(images=img)["pixel_values"][0].cpu()
    assert torch.equal(half, torch.from_numpy(ref).to(torch.float16))


def test_preprocess_strided_rows_and_errors(cuda):
    from oracle import preprocess as op
    from pigeon_b200 import PigeonB200Error
    from pigeon_b200.preprocess import ClipImageProcessor
    rng = np.random.default_rng(5)
    big = torch.from_numpy(rng.integers(0, 256, (500, 700, 3), dtype=np.uint8)).to(cuda)
    view = big[20:420, 31:631]                                   # row stride 2100 B, unaligned start
    proc = ClipImageProcessor(device=cuda)
    got = proc.preprocess_device([view])[0].cpu().numpy()
    assert np.array_equal(got, op.clip_preprocess(view.cpu().numpy()))
    with pytest.raises(ValueError):
        proc(images=np.zeros((10, 10), dtype=np.uint8))
    with pytest.raises(PigeonB200Error):
        proc(images=np.zeros((8, 40000, 3), dtype=np.uint8))    # row longer than the staged-row limit


def test_preprocess_feeds_the_tower(cuda):
    """uint8 images -> pg_preprocess_clip (fp16) -> SuperGuessr == the same pixels pre-processed by the oracle."""
    from oracle import preprocess as op
    from pigeon_b200 import SuperGuessr, synthetic
    from pigeon_b200.preprocess import ClipImageProcessor
    from pigeon_b200.super_guessr import CLIPVisionTower
    from pigeon_b200.vit_engine import VitDims
    dims = VitDims(image_size=336, patch_size=14, hidden=256, heads=4, intermediate=512, layers=2)
    tower = CLIPVisionTower(dims)
    tower.load_state_dict(synthetic.random_vit_state_dict(dims, seed=4, std=0.05), strict=True)
    cells = synthetic.synthetic_geocells(64, 0)
    sg = SuperGuessr(tower, panorama=True, freeze_base=True, num_candidates=5, geocells=cells).to(cuda).eval()
    imgs = [synthetic.synthetic_photo(400 + 10 * i, 520 - 7 * i, seed=i) for i in range(8)]     # 2 panoramas x 4 views
    px = ClipImageProcessor(device=cuda, dtype=torch.float16)(images=imgs)["pixel_values"].reshape(2, 12, 336, 336)
    ref_px = torch.from_numpy(np.stack([op.clip_preprocess(im) for im in imgs])).reshape(2, 12, 336, 336)
    lab = torch.tensor([1, 2])
    a = sg(pixel_values=px, labels_clf=lab)
    b = sg(pixel_values=ref_px.to(cuda, torch.float16), labels_clf=lab)
    assert torch.equal(a.embedding, b.embedding) and torch.equal(a.preds_geocell, b.preds_geocell)


def test_embed_images_from_raw_uint8(cuda, tmp_path):
    """Bulk-embedding driver (preprocessing/embed.py:45-83) fed RAW images: GPU pre-processing + tower in the loop ==
    embedding the oracle-pre-processed pixels; `.npy` outputs in the reference's format (embed.py:41-43)."""
    from oracle import preprocess as op
    from pigeon_b200 import CLIPEmbedding, synthetic
    from pigeon_b200.loops import embed_images, raw_image_collate
    from pigeon_b200.super_guessr import CLIPVisionTower
    from pigeon_b200.vit_engine import VitDims
    dims = VitDims(image_size=336, patch_size=14, hidden=256, heads=4, intermediate=512, layers=1)
    tower = CLIPVisionTower(dims)
    tower.load_state_dict(synthetic.random_vit_state_dict(dims, seed=6, std=0.05), strict=True)
    model = CLIPEmbedding("unused", device="cuda", clip_model=tower)
    imgs = [synthetic.synthetic_photo(360 + 13 * i, 500 - 11 * i, seed=20 + i) for i in range(5)]

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return len(imgs)

        def __getitem__(self, i):
            return imgs[i], i

    embed_images(model, {"train": DS()}, save_dir=str(tmp_path), collate_fn=raw_image_collate)
    got = np.load(tmp_path / "train.npy")
    idx = np.load(tmp_path / "train_indices.npy")
    assert got.shape == (5, 256) and idx.tolist() == [0, 1, 2, 3, 4]
    ref_px = torch.from_numpy(np.stack([op.clip_preprocess(im) for im in imgs])).to(cuda, torch.float16)
    ref = model(ref_px).cpu().numpy()
    assert np.array_equal(got, ref)
