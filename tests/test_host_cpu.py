"""CPU-only tests (`-m "not gpu"`): host logic against the oracle / golden vectors, C-ABI surface, loud failure without
a GPU, and the world_size-2 gloo path of the multi-GPU helpers."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(os.path.dirname(__file__), "golden")


def test_abi_exports_every_declared_symbol():
    from micro_sam_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "msam_b200.h")).read()
    names = sorted(set(re.findall(r"\b(msam_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15, names
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    L.msam_last_error.restype = ctypes.c_char_p
    assert isinstance(L.msam_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_fails_loudly_without_gpu():
    from micro_sam_b200 import _lib, util
    h = ctypes.c_void_p()
    cfg = _lib.MsamConfig(768, 12, 12, (ctypes.c_int32 * 8)(2, 5, 8, 11, -1, -1, -1, -1), 14, 1024, 16, 256, 1, 64)
    assert _lib.lib().msam_create(ctypes.byref(cfg), 0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in _lib.lib().msam_last_error()
    with pytest.raises(RuntimeError):
        util.get_sam_model("vit_b", state_dict={"x": torch.zeros(1)})
    with pytest.raises(RuntimeError):
        util.get_sam_model("vit_b", device="cpu", state_dict={"x": torch.zeros(1)})


def _unpack(packed, shape):
    n, h, w = shape
    return np.unpackbits(packed, axis=-1)[..., :w].astype(bool).reshape(n, h, w)


def test_host_helpers_match_reference_golden():
    from micro_sam_b200 import _amg_utils as A, util
    z = np.load(os.path.join(G, "util.npz"))
    for k in ("gray_f32", "gray_u16", "one_ch", "two_ch", "rgb_u8", "const"):
        assert np.array_equal(util._to_image(z[f"in_{k}"]), z[f"out_{k}"]), k
    with pytest.raises(ValueError):
        util._to_image(np.zeros((2, 2, 2, 2)))
    v = np.load(os.path.join(G, "vendored.npz"))
    for tag in "abc":
        m = _unpack(v[f"masks_{tag}"], v[f"shape_{tag}"])
        rles = A.mask_to_rle(m)
        assert np.array_equal(np.concatenate([np.asarray(r["counts"]) for r in rles]), v[f"rle_counts_{tag}"])
        for r, mm in zip(rles, m):
            assert np.array_equal(A.rle_to_mask(r), mm) and A.area_from_rle(r) == mm.sum()
    assert A.mask_to_rle(np.zeros((0, 4, 4), bool)) == []


def test_blocking_matches_reference_expectations():
    """test/test_util.py:179-208: 512^2 image, tile 256, halo 16 -> 4 tiles, row-major ids, halo clipped to the image."""
    from micro_sam_b200._amg_utils import Blocking
    b = Blocking([0, 0], [512, 512], [256, 256])
    assert b.number_of_blocks == 4 and b.blocks_per_axis == [2, 2]
    t = b.get_block_with_halo(1, [16, 16])
    assert (t.inner_block.begin, t.inner_block.end) == ([0, 256], [256, 512])
    assert (t.outer_block.begin, t.outer_block.end) == ([0, 240], [272, 512])
    assert t.inner_block_local.begin == [0, 16] and t.outerBlock is t.outer_block
    assert b.coordinates_to_block_id([300, 10]) == 2 and b.block_grid_position(3) == [1, 1]
    r = Blocking([0, 0], [500, 700], [256, 256])  # ragged border tiles
    assert r.number_of_blocks == 6 and r.get_block(5).shape == [244, 188]


def test_segmentation_assembly_matches_oracle():
    from oracle import amg_ref
    from micro_sam_b200 import util
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[:96, :128]
    recs = []
    for k in range(14):
        cy, cx, r = rng.integers(0, 96), rng.integers(0, 128), rng.integers(4, 25)
        m = (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
        if k % 5 == 0:  # disconnected mask -> connected components must split it
            m |= (yy - (cy + 40) % 96) ** 2 + (xx - (cx + 50) % 128) ** 2 < 9
        recs.append({"segmentation": m, "area": int(m.sum())})
    for kw in (dict(), dict(merge_exclusively=False), dict(with_background=True), dict(min_object_size=30),
               dict(merge_exclusively=False, with_background=True, min_object_size=10)):
        a = util.mask_data_to_segmentation(recs, **kw)
        b = amg_ref.mask_data_to_segmentation(recs, **kw)
        assert a.dtype == np.uint32 and np.array_equal(a == 0, b == 0)
        pairs = np.unique(np.stack([a.ravel(), b.ravel()], 1), axis=0)
        assert len(pairs) == len(np.unique(a)) == len(np.unique(b)), kw  # equal up to relabelling
        assert a.max() == len(np.unique(a)) - 1                           # consecutive ids


def test_maskdata_and_point_grid():
    from micro_sam_b200 import _amg_utils as A
    from oracle import amg_ref
    assert np.array_equal(A.build_point_grid(32), amg_ref.build_point_grid(32))
    assert A.generate_crop_boxes((300, 500), 1, 512 / 1500) == amg_ref.generate_crop_boxes((300, 500), 1, 512 / 1500)
    d = A.MaskData(a=torch.arange(6), b=list("abcdef"), c=np.arange(6) * 2)
    d.filter(torch.tensor([True, False, True, True, False, False]))
    assert d["a"].tolist() == [0, 2, 3] and d["b"] == ["a", "c", "d"] and d["c"].tolist() == [0, 4, 6]
    d.cat(A.MaskData(a=torch.tensor([9]), b=["z"], c=np.array([1])))
    d.filter(torch.tensor([3, 0]))
    assert d["a"].tolist() == [9, 0] and d["b"] == ["z", "a"]
    assert [len(x[0]) for x in A.batch_iterator(4, list(range(10)))] == [4, 4, 2]


def _dist_worker(rank, world, port, ret):
    import torch.distributed as dist
    from micro_sam_b200 import distributed as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        lo, hi = D.shard_range(11, rank, world)
        table = torch.arange(lo, hi, dtype=torch.float32)[:, None] * torch.ones(1, 3)
        table[:, 1] = rank
        full, counts = D.all_gather_tables(table)
        offs = D.exclusive_id_offsets(torch.arange(lo, hi) + 1)
        # the instance-table exchange of the multi-rank tiled AMG: several tensors of different trailing shapes / dtypes
        tabs = {"gbox": torch.arange(lo, hi, dtype=torch.int32)[:, None].repeat(1, 4), "tile": torch.full((hi - lo,), rank, dtype=torch.int32),
                "low": torch.arange(lo, hi, dtype=torch.float32)[:, None, None] * torch.ones(1, 2, 3)}
        g, c = D.gather_instance_tables(tabs)
        assert c == counts and g["gbox"].shape == (11, 4) and g["low"].shape == (11, 2, 3) and g["tile"].dtype == torch.int32
        assert g["gbox"][:, 0].tolist() == list(range(11)) and g["low"][:, 1, 2].tolist() == list(map(float, range(11)))
        assert g["tile"].tolist() == [0] * counts[0] + [1] * counts[1]
        # a rank that owns no instance at all still takes part in the exchange
        e, ce = D.gather_instance_tables({"low": torch.zeros(3 if rank == 1 else 0, 2, 3), "tile": torch.zeros(3 if rank == 1 else 0, dtype=torch.int32)})
        assert ce == [0, 3] and e["low"].shape == (3, 2, 3) and e["tile"].shape == (3,)
        # the gradient exchange of the data-parallel training step: in-place average of tensors of different shapes
        gs = [torch.full((3, 2), float(rank + 1)), torch.arange(5, dtype=torch.float32) * (rank + 1)]
        assert D.allreduce_average_(gs) == 11
        assert torch.allclose(gs[0], torch.full((3, 2), 1.5)) and torch.allclose(gs[1], torch.arange(5, dtype=torch.float32) * 1.5)
        ret[rank] = (lo, hi, full.tolist(), counts, offs.tolist())
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharding_and_instance_table_gather():
    import socket
    import torch.multiprocessing as mp
    from micro_sam_b200 import distributed as D
    for n in (0, 1, 7, 256):
        cov = []
        for r in range(4):
            lo, hi = D.shard_range(n, r, 4)
            cov += list(range(lo, hi))
            assert hi - lo in (n // 4, n // 4 + 1)
        assert cov == list(range(n))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dist_worker, args=(2, port, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    (lo0, hi0, full0, counts0, offs0), (lo1, hi1, full1, counts1, offs1) = ret[0], ret[1]
    assert (lo0, hi0, lo1, hi1) == (0, 5, 5, 11) and counts0 == counts1 == [5, 6]
    assert full0 == full1 and [row[0] for row in full0] == list(map(float, range(11)))
    assert [row[1] for row in full0] == [0.0] * 5 + [1.0] * 6
    scan = np.cumsum(np.arange(1, 12)) - np.arange(1, 12)
    assert offs0 == scan[:5].tolist() and offs1 == scan[5:].tolist()
    # without a process group the helpers are the identity
    t = torch.ones(3, 2)
    assert D.all_gather_tables(t)[1] == [3]


def test_stitch_segmentation_matches_oracle():
    """inference._stitch_segmentation / _merge_segmentations (inference.py:315-356): earlier tiles are preserved."""
    from micro_sam_b200 import inference
    from micro_sam_b200._amg_utils import Blocking
    from oracle import amg_ref
    shape, tile_shape, halo = (300, 420), (160, 224), (24, 24)
    tiling, otiling = Blocking([0, 0], shape, tile_shape), amg_ref.Blocking([0, 0], shape, tile_shape)
    ids = [0, 1, 3]
    segs = [np.random.default_rng(i).integers(0, 4, tuple(tiling.get_block_with_halo(i, list(halo)).outer_block.shape)).astype("uint32")
            for i in ids]
    a = inference._stitch_segmentation([s.copy() for s in segs], ids, tiling, halo, shape)
    b = amg_ref.stitch_segmentation([s.copy() for s in segs], ids, otiling, halo, shape)
    assert a.dtype == np.uint32 and np.array_equal(a, b)
    assert tiling.coordinates_to_block_id([200, 300]) == amg_ref.coordinates_to_block_id(otiling, [200, 300]) == 3


def test_tiled_prompt_routing_matches_oracle_loop():
    """inference._route_prompts_to_tiles (vectorised) == the reference's per-prompt loop as restated in the oracle
    (inference.py:424-470), for boxes and for points."""
    from micro_sam_b200 import inference
    from micro_sam_b200._amg_utils import Blocking
    from micro_sam_b200.sample_data import random_boxes
    from oracle import amg_ref
    shape, tile_shape, halo = (300, 420), (160, 224), (24, 24)
    tiling, ot = Blocking([0, 0], shape, tile_shape), amg_ref.Blocking([0, 0], shape, tile_shape)
    boxes = random_boxes(60, shape, seed=3)
    ids, b2t, _, _ = inference._route_prompts_to_tiles(tiling, halo, boxes, None, None)
    ref = {}
    for box in boxes:
        c = np.array([(box[1] + box[3]) / 2, (box[0] + box[2]) / 2]).round().astype("int").tolist()
        tid = amg_ref.coordinates_to_block_id(ot, c)
        t = ot.get_block_with_halo(tid, list(halo)).outer_block
        b = np.array([max(box[1] - t.begin[0], 0), max(box[0] - t.begin[1], 0), min(box[3] - t.begin[0], t.shape[0]),
                      min(box[2] - t.begin[1], t.shape[1])])[None]
        ref[tid] = np.concatenate([ref[tid], b]) if tid in ref else b
    assert ids == sorted(ref) and all(np.array_equal(b2t[t], ref[t]) for t in ids)
    pts = (np.random.default_rng(0).random((50, 1, 2)) * [419, 299]).astype(np.float64)
    lbl = np.ones((50, 1), dtype=np.int64)
    ids, _, p2t, l2t = inference._route_prompts_to_tiles(tiling, halo, None, pts, lbl)
    refp = {}
    for k in range(50):
        tid = amg_ref.coordinates_to_block_id(ot, pts[k, 0][::-1].round().astype("int").tolist())
        t = ot.get_block_with_halo(tid, list(halo)).outer_block
        pin = (pts[k, 0] - np.array(t.begin)[::-1])[None, None]
        refp[tid] = np.concatenate([refp[tid], pin]) if tid in refp else pin
    assert ids == sorted(refp) and all(np.array_equal(p2t[t], refp[t]) and len(l2t[t]) == len(refp[t]) for t in ids)


def test_checkpoint_loading_tolerates_missing_training_classes(tmp_path):
    """util._load_checkpoint (util.py:246-290): torch_em checkpoints pickle trainer objects whose classes need not be
    importable; only the weights are wanted.  `sam.` / `module.` prefixes are stripped."""
    import sys
    import types
    import warnings
    from micro_sam_b200 import util
    mod = types.ModuleType("ghost_trainer_mod")
    sys.modules["ghost_trainer_mod"] = mod

    class Trainer:
        def __init__(self):
            self.lr = 1e-4
    Trainer.__module__, Trainer.__qualname__ = "ghost_trainer_mod", "Trainer"
    mod.Trainer = Trainer
    p = str(tmp_path / "best.pt")
    torch.save({"model_state": {"sam.image_encoder.pos_embed": torch.ones(2), "module.sam.x": torch.zeros(1)}, "trainer": Trainer(),
                "epoch": 3}, p)
    del sys.modules["ghost_trainer_mod"]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        state, model_state = util._load_checkpoint(p)
    assert sorted(model_state) == ["image_encoder.pos_embed", "x"] and state["epoch"] == 3
    assert any("ghost_trainer_mod" in str(x.message) for x in w)
    torch.save({"image_encoder.pos_embed": torch.ones(2)}, p)
    assert sorted(util._load_checkpoint(p)[1]) == ["image_encoder.pos_embed"]


def test_tiled_mask_nms_without_the_dense_canvas_equals_the_dense_rule():
    """util._tiled_mask_nms_sparse (large tiled images: intersections on box-overlap windows only) keeps exactly the set the dense
    greedy rule keeps (descending score, keep iou <= thresh), for IoU and intersection-over-min."""
    from micro_sam_b200 import util
    rng = np.random.default_rng(0)
    H, W, tile = 96, 128, 64
    preds, dense = [], []
    for k in range(40):
        ty, tx = int(rng.integers(0, H - tile + 1)), int(rng.integers(0, W - tile + 1))       # tile origin in the image
        m = np.zeros((tile, tile), bool)
        cy, cx, r = rng.integers(8, tile - 8), rng.integers(8, tile - 8), rng.integers(3, 9)
        yy, xx = np.mgrid[:tile, :tile]
        m[(yy - cy) ** 2 + (xx - cx) ** 2 < r * r] = True
        ys, xs = np.where(m)
        bbox = [int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)]
        gb = [bbox[0] + tx, bbox[1] + ty, bbox[2], bbox[3]]
        preds.append({"segmentation": m, "bbox": bbox, "global_bbox": gb, "predicted_iou": float(rng.random()), "stability_score": 1.0})
        full = np.zeros((H, W), bool)
        full[ty:ty + tile, tx:tx + tile] = m
        dense.append(full)
    dense = np.stack(dense)
    areas = [int(d.sum()) for d in dense]
    scores = [p["predicted_iou"] for p in preds]
    for iomin in (False, True):
        for thr in (0.1, 0.5):
            order = np.argsort(-np.asarray(scores), kind="stable")
            alive, ref = np.ones(len(preds), bool), []
            for pos, i in enumerate(order):
                if not alive[i]:
                    continue
                ref.append(int(i))
                for j in order[pos + 1:]:
                    inter = np.float32((dense[i] & dense[j]).sum())
                    den = np.float32(min(areas[i], areas[j])) if iomin else np.float32(areas[i] + areas[j] - inter)
                    if np.float32(inter / den) > np.float32(thr):
                        alive[j] = False
            got = util._tiled_mask_nms_sparse(preds, list(range(len(preds))), scores, thr, iomin, areas).tolist()
            assert got == ref and len(ref) < len(preds), (iomin, thr, got, ref)


def test_precompute_state_loads_a_cached_amg_state_without_a_gpu(tmp_path):
    """precompute_state.cache_amg_state: an existing `amg_state.pickle` / `amg_state/state-<i>.pkl` is loaded instead of recomputed
    (reference file names, micro_sam/precompute_state.py:52-64), and the pickle holds CPU tensors only."""
    import pickle
    import torch
    from micro_sam_b200 import precompute_state as ps

    class _Pred:    # the generator's constructor only stores the predictor
        device = "cpu"
    state = {"crop_list": [{"iou_preds": torch.rand(6), "boxes": torch.zeros(6, 4)}], "crop_boxes": [[0, 0, 8, 8]], "original_size": (8, 8)}
    with open(tmp_path / "amg_state.pickle", "wb") as f:
        pickle.dump(ps._state_to_cpu(state), f)
    amg = ps.cache_amg_state(_Pred(), np.zeros((8, 8)), {"input_size": (8, 8)}, str(tmp_path), verbose=False)
    assert amg.is_initialized and amg.get_state()["original_size"] == (8, 8)
    os.makedirs(tmp_path / "amg_state")
    with open(tmp_path / "amg_state" / "state-3.pkl", "wb") as f:
        pickle.dump(ps._state_to_cpu(state), f)
    amg = ps.cache_amg_state(_Pred(), np.zeros((5, 8, 8)), {"input_size": (8, 8)}, str(tmp_path), verbose=False, i=3)
    assert torch.equal(amg.get_state()["crop_list"][0]["iou_preds"], state["crop_list"][0]["iou_preds"])
    np.save(tmp_path / "img.npy", np.arange(12).reshape(3, 4))
    assert ps.load_image_data(str(tmp_path / "img.npy")).shape == (3, 4)
    with pytest.raises(NotImplementedError):
        ps.cache_is_state()


def test_prompt_table_index_follows_the_prompt_encoder():
    """sam.prompt_table_index (gradient routing of the training decoder) against the oracle PromptEncoder: a sparse token minus its
    positional part equals the embedding row the index names."""
    import torch
    from oracle import sam_ref
    from micro_sam_b200.sam import prompt_table_index
    osam = sam_ref.build_seeded_sam("vit_test", seed=1)
    pe = osam.prompt_encoder
    tables = torch.cat([pe.point_embeddings[i].weight for i in range(4)] + [pe.not_a_point_embed.weight]).detach()
    g = torch.Generator().manual_seed(0)
    P = 3
    coords = torch.rand(P, 2, 2, generator=g) * 1000
    labels = torch.tensor([[1.0, 0.0], [0.0, -1.0], [1.0, 1.0]])
    boxes = torch.rand(P, 4, generator=g) * 500 + torch.tensor([0.0, 0.0, 500.0, 500.0])
    for pts, bx in (((coords, labels), None), (None, boxes), ((coords, labels), boxes)):
        with torch.no_grad():
            sparse, _ = pe(points=pts, boxes=bx, masks=None)
        idx = prompt_table_index(None if pts is None else pts[1], bx is not None, P)
        assert idx.shape == sparse.shape[:2]
        # positional part: what the encoder produces with zeroed tables
        saved = [p.detach().clone() for p in (*[e.weight for e in pe.point_embeddings], pe.not_a_point_embed.weight)]
        with torch.no_grad():
            for p in (*[e.weight for e in pe.point_embeddings], pe.not_a_point_embed.weight):
                p.zero_()
            pos, _ = pe(points=pts, boxes=bx, masks=None)
            for p, s in zip((*[e.weight for e in pe.point_embeddings], pe.not_a_point_embed.weight), saved):
                p.copy_(s)
        assert torch.allclose(sparse - pos, tables[idx], atol=1e-6)


def test_resize_of_replicated_gray_images_is_bit_identical_to_the_rgb_resize():
    """ResizeLongestSide.apply_image: the single-band shortcut for gray images (3 equal channels) against PIL's RGB resize, up- and
    down-scaling (antialiased) and non-square shapes; true RGB inputs take the ordinary path."""
    from PIL import Image
    from micro_sam_b200.sam import ResizeLongestSide, get_preprocess_shape
    rng = np.random.default_rng(0)
    rs = ResizeLongestSide(1024)
    for shape in ((512, 512), (1152, 1152), (300, 500), (2048, 1536), (1024, 1024)):
        g = rng.integers(0, 256, shape, dtype=np.uint8)
        img = np.repeat(g[..., None], 3, axis=2)
        th, tw = get_preprocess_shape(shape[0], shape[1], 1024)
        ref = np.array(Image.fromarray(img).resize((tw, th), Image.BILINEAR))
        out = rs.apply_image(img)
        assert out.shape == (th, tw, 3) and out.flags["C_CONTIGUOUS"] and np.array_equal(out, ref), shape
    rgb = rng.integers(0, 256, (400, 640, 3), dtype=np.uint8)
    th, tw = get_preprocess_shape(400, 640, 1024)
    assert np.array_equal(rs.apply_image(rgb), np.array(Image.fromarray(rgb).resize((tw, th), Image.BILINEAR)))


def test_mask_data_to_segmentation_tiled_records_against_the_oracle():
    """Tiled records (tile-local mask + bbox, global_bbox): the host painter against the oracle restatement of util.py:1799-1829."""
    from micro_sam_b200 import util
    from oracle import amg_ref
    rng = np.random.default_rng(3)
    H, W, tile = 120, 160, 64
    recs = []
    yy, xx = np.mgrid[:tile, :tile]
    for k in range(25):
        ty, tx = int(rng.integers(0, H - tile + 1)), int(rng.integers(0, W - tile + 1))
        cy, cx, r = rng.integers(6, tile - 6), rng.integers(6, tile - 6), rng.integers(3, 14)
        m = (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
        ys, xs = np.where(m)
        bbox = [int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)]
        recs.append({"segmentation": m, "area": int(m.sum()), "bbox": bbox, "global_bbox": [bbox[0] + tx, bbox[1] + ty, bbox[2], bbox[3]]})
    for kw in (dict(), dict(merge_exclusively=False), dict(min_object_size=40, with_background=True)):
        a = util.mask_data_to_segmentation(recs, shape=(H, W), **kw)
        b = amg_ref.mask_data_to_segmentation(recs, shape=(H, W), **kw)
        assert np.array_equal(a == 0, b == 0) and a.max() > 0
        pairs = np.unique(np.stack([a.ravel(), b.ravel()], 1), axis=0)
        assert len(pairs) == len(np.unique(a)) == len(np.unique(b)), kw
