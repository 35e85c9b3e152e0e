"""Pin oracle/amg_ref.py (and, where noted, the product's host code) against (a) golden vectors produced by executing the reference's own code
(tests/golden/make_golden.py) and (b) the reference's own known-answer tests."""
import os

import numpy as np
import torch

from oracle import amg_ref

G = os.path.join(os.path.dirname(__file__), "golden")


def _unpack(packed, shape):
    n, h, w = shape
    return np.unpackbits(packed, axis=-1)[..., :w].astype(bool).reshape(n, h, w)


def test_vendored_golden_boxes_and_rle():
    z = np.load(os.path.join(G, "vendored.npz"))
    for tag in "abc":
        m = torch.from_numpy(_unpack(z[f"masks_{tag}"], z[f"shape_{tag}"]))
        assert np.array_equal(amg_ref.batched_mask_to_box(m).numpy(), z[f"boxes_{tag}"])
        rles = amg_ref.mask_to_rle(m)
        counts = np.concatenate([np.asarray(r["counts"], dtype=np.int64) for r in rles])
        assert np.array_equal(counts, z[f"rle_counts_{tag}"])
        assert np.array_equal(np.array([len(r["counts"]) for r in rles]), z[f"rle_lens_{tag}"])
        for r, mm in zip(rles, m.numpy()):
            assert sum(r["counts"]) == mm.size                       # test/test_vendored.py:64-70
            assert np.array_equal(amg_ref.rle_to_mask(r), mm)        # round trip
            assert amg_ref.area_from_rle(r) == mm.sum()


def test_reference_kat_mask_to_box():
    """test/test_vendored.py:12-25."""
    mask = np.zeros((10, 10), dtype=bool)
    mask[7:9, 3:5] = True
    assert amg_ref.batched_mask_to_box(torch.from_numpy(mask)).tolist() == [3, 7, 4, 8]


def test_to_image_golden():
    z = np.load(os.path.join(G, "util.npz"))
    for k in ("gray_f32", "gray_u16", "one_ch", "two_ch", "rgb_u8", "const"):
        out = amg_ref.to_image(z[f"in_{k}"])
        assert out.dtype == np.uint8 and np.array_equal(out, z[f"out_{k}"]), k


def test_mask_nms_golden():
    z = np.load(os.path.join(G, "util.npz"))
    masks = torch.from_numpy(_unpack(z["nms_masks"], (24, 64, 64)))
    boxes, scores = torch.from_numpy(z["nms_boxes"]), torch.from_numpy(z["nms_scores"])
    np.testing.assert_allclose(amg_ref.ious_between_pred_masks(masks, boxes).numpy(), z["nms_iou_matrix"], rtol=0, atol=1e-7)
    for thr in (0.3, 0.9):
        assert amg_ref.batched_mask_nms(masks, boxes, scores, thr, False).tolist() == z[f"nms_keep_iou_{thr}"].tolist()
        assert amg_ref.batched_mask_nms(masks, boxes, scores, thr, True).tolist() == z[f"nms_keep_iomin_{thr}"].tolist()


def test_tiled_mask_nms_golden():
    """util.py:1687-1770 (tile-local masks + global boxes): oracle == outputs of the reference's own functions."""
    z = np.load(os.path.join(G, "tiled_nms.npz"))
    n = int(z["n"])
    masks = [z[f"mask_{k}"] for k in range(n)]
    boxes, gboxes, scores = z["boxes"], z["global_boxes"], z["scores"]
    preds = [{"segmentation": masks[k], "bbox": boxes[k].tolist(), "global_bbox": gboxes[k].tolist()} for k in range(n)]
    assert list(amg_ref.infer_tiled_shape(preds)) == z["inferred_shape"].tolist()
    for iomin in (False, True):
        mat = amg_ref.tiled_mask_overlap_matrix(masks, boxes, gboxes, iomin).numpy()
        np.testing.assert_allclose(mat, z[f"overlap_{int(iomin)}"], rtol=0, atol=1e-7)
        for thr in (0.3, 0.9):
            keep = amg_ref.batched_tiled_mask_nms(masks, boxes, gboxes, scores, thr, iomin).tolist()
            assert keep == z[f"keep_{int(iomin)}_{thr}"].tolist()


def test_stitch_segmentation_golden():
    """inference._stitch_segmentation / _merge_segmentations: oracle AND product host code == the output of the reference's
    own functions (tests/golden/stitch.npz), incl. tile subsets and a sequence that does not start with tile 0."""
    from micro_sam_b200 import inference
    from micro_sam_b200._amg_utils import Blocking
    z = np.load(os.path.join(G, "stitch.npz"))
    shape, tile_shape, halo = tuple(z["shape"].tolist()), tuple(z["tile_shape"].tolist()), tuple(z["halo"].tolist())
    ot, pt = amg_ref.Blocking([0, 0], shape, tile_shape), Blocking([0, 0], shape, tile_shape)
    for name in ("all", "subset", "no_first"):
        ids = z[f"{name}_ids"].tolist()
        segs = [z[f"{name}_seg{k}"] for k in range(len(ids))]
        out = z[f"{name}_out"]
        assert np.array_equal(amg_ref.stitch_segmentation([s.copy() for s in segs], ids, ot, halo, shape), out), name
        assert np.array_equal(inference._stitch_segmentation([s.copy() for s in segs], ids, pt, halo, shape), out), name


def test_local_otsu_golden():
    """mask_threshold="auto": the oracle's per-mask restatement == the reference's own `_local_otsu_threshold` (executed from
    its source by tests/golden/make_golden.py) on the seeded inputs, bit for bit."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    z = np.load(os.path.join(G, "otsu.npz"))
    x = mg.otsu_inputs()
    assert abs(x.double().sum().item() - float(z["checksum"])) < 1e-6, "seeded inputs drifted"
    thr = amg_ref.local_otsu_threshold(x).reshape(-1).numpy()
    assert np.array_equal(thr, z["thresholds"])


def test_box_nms_matches_torchvision():
    import torchvision
    g = torch.Generator().manual_seed(0)
    for n in (1, 17, 400):
        xy = torch.randint(0, 900, (n, 2), generator=g)
        wh = torch.randint(1, 200, (n, 2), generator=g)
        boxes = torch.cat([xy, xy + wh], 1).float()
        boxes[: n // 2] = boxes[n // 2: 2 * (n // 2)]
        scores = torch.rand(n, generator=g)
        assert amg_ref.nms(boxes, scores, 0.7).tolist() == torchvision.ops.nms(boxes, scores, 0.7).tolist()


def test_point_grid_and_crop_boxes():
    g = amg_ref.build_point_grid(32)
    assert g.shape == (1024, 2) and np.isclose(g[0, 0], 1 / 64) and np.isclose(g[-1, 1], 1 - 1 / 64)
    boxes, layers = amg_ref.generate_crop_boxes((100, 200), 0, 512 / 1500)
    assert boxes == [[0, 0, 200, 100]] and layers == [0]


def test_mask_data_to_segmentation_semantics():
    a = np.zeros((8, 8), bool); a[:6, :6] = True
    b = np.zeros((8, 8), bool); b[2:4, 2:4] = True
    recs = [dict(segmentation=b, area=int(b.sum())), dict(segmentation=a, area=int(a.sum()))]
    seg = amg_ref.mask_data_to_segmentation(recs, merge_exclusively=False)   # AMG: smaller overwrites
    assert seg[3, 3] != seg[0, 0] and seg[3, 3] != 0 and seg.max() == 2
    seg = amg_ref.mask_data_to_segmentation(recs, merge_exclusively=True)    # first (largest) wins
    assert seg[3, 3] == seg[0, 0] and seg.max() == 1
