"""CPU tests of the embedding container (a6): zarr-v2 layout, signature, cache hit, invalidation, resume of partial 3-D
runs, tiled layouts, rank sharding -- the host logic of micro_sam/util.py:684-1212 with a stand-in encoder (the container
code never looks at the values).  Mirrors the reference's own checks in test/test_util.py:106-246 (shapes, zarr layout,
4 tiles for 512^2 / tile 256 / halo 16)."""
import json
import os

import numpy as np
import pytest
import torch

from micro_sam_b200 import util, zarr_store
from micro_sam_b200.sam import ResizeLongestSide


class _FakeSam:
    device = torch.device("cpu")
    image_size = 1024
    calls = 0

    def encode_u8(self, x):   # deterministic function of the batch content
        _FakeSam.calls += x.shape[0]
        m = x.float().mean(dim=(1, 2, 3))
        return m[:, None, None, None] + torch.arange(256, dtype=torch.float32)[None, :, None, None] * torch.ones(1, 1, 64, 64)

    def preprocess(self, x):
        return torch.nn.functional.pad(x.float(), (0, 1024 - x.shape[-1], 0, 1024 - x.shape[-2]))

    def image_encoder(self, x):
        return self.encode_u8(x.permute(0, 2, 3, 1))


class _FakePredictor:
    def __init__(self):
        self.model = _FakeSam()
        self.transform = ResizeLongestSide(1024)
        self.model_type, self.model_name, self._hash = "vit_b", "vit_b", None
        self.device = torch.device("cpu")
        self.reset_image()

    def reset_image(self):
        self.is_image_set, self.features = False, None

    def set_image(self, image):
        x = self.transform.apply_image(image)
        self.original_size, self.input_size = image.shape[:2], x.shape[:2]
        self.features = self.model.encode_u8(torch.from_numpy(x)[None])
        self.is_image_set = True

    def get_image_embedding(self):
        return self.features


def test_2d_container_roundtrip_and_signature(tmp_path):
    pred = _FakePredictor()
    img = np.random.default_rng(0).integers(0, 255, (96, 128)).astype("uint8")
    path = str(tmp_path / "emb.zarr")
    e1 = util.precompute_image_embeddings(pred, img, save_path=path)
    assert e1["features"].shape == (1, 256, 64, 64) and tuple(e1["input_size"]) == (768, 1024) and tuple(e1["original_size"]) == (96, 128)
    # zarr v2 layout on disk (what zarr.open would read)
    assert json.load(open(os.path.join(path, ".zgroup"))) == {"zarr_format": 2}
    za = json.load(open(os.path.join(path, "features", ".zarray")))
    assert za["shape"] == [1, 256, 64, 64] and za["dtype"] == "<f4" and za["compressor"] is None and za["order"] == "C"
    raw = np.fromfile(os.path.join(path, "features", "0.0.0.0"), "<f4").reshape(1, 256, 64, 64)
    assert np.array_equal(raw, e1["features"])
    attrs = json.load(open(os.path.join(path, ".zattrs")))
    for k in ("data_signature", "tile_shape", "halo", "model_type", "model_name", "micro_sam_version", "model_hash", "input_size",
              "original_size"):
        assert k in attrs
    # second call: loaded, not recomputed; predictor is set (util.py:905-912)
    n = _FakeSam.calls
    pred2 = _FakePredictor()
    e2 = util.precompute_image_embeddings(pred2, img, save_path=path)
    assert _FakeSam.calls == n and np.array_equal(e2["features"], e1["features"]) and pred2.is_image_set
    # different data / model type -> RuntimeError (util.py:1077-1102); version mismatch only warns
    with pytest.raises(RuntimeError, match="data_signature"):
        util.precompute_image_embeddings(pred2, img + 1, save_path=path)
    pred2.model_type = "vit_l"
    with pytest.raises(RuntimeError, match="model_type"):
        util.precompute_image_embeddings(pred2, img, save_path=path)
    pred2.model_type, pred2.model_name = "vit_b", "other"
    with pytest.warns(UserWarning):
        util.precompute_image_embeddings(pred2, img, save_path=path)


def test_3d_container_resume_and_lazy(tmp_path):
    pred = _FakePredictor()
    vol = np.random.default_rng(1).integers(0, 255, (5, 64, 64)).astype("uint8")
    path = str(tmp_path / "vol.zarr")
    # a partial run: datasets exist, slices 0 and 3 written, no signature yet
    f = zarr_store.open_group(path)
    ds = f.create_dataset("features", shape=(5, 1, 256, 64, 64), chunks=(1, 1, 256, 64, 64), dtype="float32")
    full = util.precompute_image_embeddings(_FakePredictor(), vol, batch_size=2)["features"]
    ds[0], ds[3] = full[0], full[3]
    n = _FakeSam.calls
    e = util.precompute_image_embeddings(pred, vol, save_path=path, batch_size=2)
    assert _FakeSam.calls - n == 3                      # only slices 1, 2, 4 were computed (util.py:988-992)
    assert e["features"].shape == (5, 1, 256, 64, 64) and np.array_equal(e["features"], full)
    assert "input_size" in zarr_store.open_group(path).attrs
    lazy = util.precompute_image_embeddings(pred, vol, save_path=path, lazy_loading=True)["features"]
    assert isinstance(lazy, zarr_store.Array) and np.array_equal(lazy[2], full[2])
    util.set_precomputed(pred, {"features": lazy, "input_size": e["input_size"], "original_size": e["original_size"]}, i=4)
    assert np.array_equal(pred.features.numpy(), full[4])
    bad = zarr_store.open_group(str(tmp_path / "bad.zarr"))
    bad.create_dataset("features", shape=(4, 1, 256, 64, 64), chunks=(1, 1, 256, 64, 64), dtype="float32")
    with pytest.raises(RuntimeError, match="Invalid partial"):
        util.precompute_image_embeddings(pred, vol, save_path=str(tmp_path / "bad.zarr"))


def test_tiled_containers_and_rank_shards(tmp_path):
    """test/test_util.py:179-246: 512^2, tile 256, halo 16 -> datasets "0".."3" with per-tile attrs; 3-D: (Z,1,256,64,64)."""
    pred = _FakePredictor()
    img = np.random.default_rng(2).integers(0, 255, (512, 512)).astype("uint8")
    path = str(tmp_path / "tiled.zarr")
    e = util.precompute_image_embeddings(pred, img, save_path=path, tile_shape=(256, 256), halo=(16, 16), batch_size=3)
    feats = e["features"]
    assert sorted(feats.keys()) == ["0", "1", "2", "3"] and e["input_size"] is None
    assert feats.attrs["tile_shape"] == [256, 256] and feats.attrs["halo"] == [16, 16] and feats.attrs["shape"] == [512, 512]
    assert feats["1"].shape == (1, 256, 64, 64) and feats["1"].attrs["original_size"] == [272, 272]
    mem = util.precompute_image_embeddings(pred, img, tile_shape=(256, 256), halo=(16, 16), batch_size=2)["features"]
    for t in "0123":
        assert np.array_equal(feats[t][:], mem[t][:])
    util.set_precomputed(pred, e, tile_id=2)
    assert tuple(pred.original_size) == (272, 272) and np.array_equal(pred.features.numpy(), mem["2"][:])
    n = _FakeSam.calls
    e2 = util.precompute_image_embeddings(pred, img, save_path=path, tile_shape=(256, 256), halo=(16, 16))
    assert _FakeSam.calls == n and sorted(e2["features"].keys()) == ["0", "1", "2", "3"]
    with pytest.raises(RuntimeError, match="halo"):
        util.precompute_image_embeddings(pred, img, save_path=path, tile_shape=(256, 256), halo=(8, 8))
    # 3-D tiled, two "ranks" filling one container (static block partition of the (z, tile) order, no collective)
    vol = np.random.default_rng(3).integers(0, 255, (3, 300, 300)).astype("uint8")
    p3 = str(tmp_path / "tiled3d.zarr")
    for rank in (1, 0):
        util.precompute_image_embeddings(_FakePredictor(), vol, save_path=p3, tile_shape=(256, 256), halo=(16, 16), batch_size=4,
                                         rank=rank, world_size=2)
    p3b = str(tmp_path / "tiled3d_single.zarr")   # single rank, three batches (regression: the batch list must survive dataset creation)
    util.precompute_image_embeddings(_FakePredictor(), vol, save_path=p3b, tile_shape=(256, 256), halo=(16, 16), batch_size=5)
    got = zarr_store.open_group(p3)["features"]
    for t in "0123":
        assert np.array_equal(zarr_store.open_group(p3b)["features"][t][:], got[t][:])
    ref = util.precompute_image_embeddings(pred, vol, tile_shape=(256, 256), halo=(16, 16), batch_size=5)["features"]
    assert sorted(got.keys()) == ["0", "1", "2", "3"]
    for t in "0123":
        assert got[t].shape == (3, 1, 256, 64, 64) and got[t].chunks == (1, 1, 256, 64, 64)
        assert np.array_equal(got[t][:], ref[t][:])
    assert "input_size" in zarr_store.open_group(p3).attrs
    # masked tiles (util.py:749-762)
    mask = np.zeros((512, 512), bool)
    mask[300:, 300:] = True
    em = util.precompute_image_embeddings(pred, img, save_path=str(tmp_path / "m.zarr"), tile_shape=(256, 256), halo=(16, 16), mask=mask)
    assert sorted(em["features"].keys()) == ["3"] and em["features"].attrs["tiles_in_mask"] == [3]
