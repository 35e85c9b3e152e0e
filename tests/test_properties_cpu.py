"""Size-independent properties of the host-side integer stages (product `_amg_utils` and the oracle), CPU only."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from micro_sam_b200 import _amg_utils as au
from oracle import amg_ref


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 4), st.integers(1, 40), st.integers(1, 40), st.integers(0, 2 ** 31 - 1), st.floats(0.0, 1.0))
def test_rle_round_trip_and_area(b, h, w, seed, density):
    """mask -> RLE -> mask is the identity, sum(counts) = H*W, odd entries sum to the area (_vendored.py:104-152;
    test/test_vendored.py:44-78 checks the same invariants against upstream)."""
    rng = np.random.default_rng(seed)
    masks = rng.random((b, h, w)) < density
    rles = au.mask_to_rle(masks)
    orles = amg_ref.mask_to_rle(torch.from_numpy(masks))
    assert len(rles) == b
    for k in range(b):
        assert rles[k]["size"] == [h, w] and sum(rles[k]["counts"]) == h * w
        assert rles[k]["counts"] == orles[k]["counts"]
        assert np.array_equal(au.rle_to_mask(rles[k]), masks[k])
        assert au.area_from_rle(rles[k]) == int(masks[k].sum())


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 300), st.integers(1, 300), st.integers(1, 128), st.integers(1, 128), st.integers(0, 40), st.integers(0, 40))
def test_blocking_tiles_the_roi(h, w, bh, bw, hy, hx):
    """Inner blocks partition the ROI; outer blocks contain their inner block, are clipped to the ROI and the local inner
    coordinates are consistent (nifty / bioimage_cpp Blocking semantics, util.py:766, inference.py:343-466)."""
    tiling = au.Blocking([0, 0], [h, w], [bh, bw])
    otiling = amg_ref.Blocking([0, 0], [h, w], [bh, bw])
    cover = np.zeros((h, w), dtype=np.int32)
    assert tiling.number_of_blocks == otiling.number_of_blocks
    for t in range(tiling.number_of_blocks):
        blk = tiling.get_block_with_halo(t, [hy, hx])
        oblk = otiling.get_block_with_halo(t, [hy, hx])
        ib, ob, lb = blk.inner_block, blk.outer_block, blk.inner_block_local
        assert ib.begin == oblk.inner_block.begin and ib.end == oblk.inner_block.end
        assert ob.begin == oblk.outer_block.begin and ob.end == oblk.outer_block.end
        cover[ib.begin[0]:ib.end[0], ib.begin[1]:ib.end[1]] += 1
        for d in range(2):
            assert 0 <= ob.begin[d] <= ib.begin[d] < ib.end[d] <= ob.end[d] <= (h, w)[d]
            assert ib.begin[d] - ob.begin[d] <= (hy, hx)[d] and ob.end[d] - ib.end[d] <= (hy, hx)[d]
            assert lb.begin[d] == ib.begin[d] - ob.begin[d] and lb.end[d] == ib.end[d] - ob.begin[d]
        centre = [(ib.begin[0] + ib.end[0] - 1) // 2, (ib.begin[1] + ib.end[1] - 1) // 2]
        assert tiling.coordinates_to_block_id(centre) == t
    assert (cover == 1).all()


@settings(max_examples=30, deadline=None)
@given(st.integers(1, 6), st.integers(4, 48), st.integers(4, 48), st.integers(0, 2 ** 31 - 1))
def test_boxes_contain_their_masks(b, h, w, seed):
    """batched_mask_to_box (_vendored.py:33-85): the box is the tight bounding box, [0,0,0,0] for an empty mask."""
    rng = np.random.default_rng(seed)
    masks = rng.random((b, h, w)) < 0.15
    masks[0] = False
    boxes = amg_ref.batched_mask_to_box(torch.from_numpy(masks)).numpy()
    assert boxes[0].tolist() == [0, 0, 0, 0]
    for k in range(1, b):
        if not masks[k].any():
            assert boxes[k].tolist() == [0, 0, 0, 0]
            continue
        ys, xs = np.where(masks[k])
        assert boxes[k].tolist() == [xs.min(), ys.min(), xs.max(), ys.max()]
        xywh = au.box_xyxy_to_xywh(boxes[k].copy())
        assert xywh[2] == xs.max() - xs.min() and xywh[3] == ys.max() - ys.min()


@settings(max_examples=25, deadline=None)
@given(st.integers(2, 40), st.integers(0, 2 ** 31 - 1), st.floats(0.1, 0.9))
def test_box_nms_idempotent_and_sorted(n, seed, thr):
    """Greedy box NMS: the keep list is in descending score order, contains the best box, and applying NMS to the kept
    boxes again keeps all of them (no pair above the threshold survives)."""
    rng = np.random.default_rng(seed)
    xy = rng.integers(0, 60, size=(n, 2))
    wh = rng.integers(1, 40, size=(n, 2))
    boxes = torch.tensor(np.concatenate([xy, xy + wh], 1), dtype=torch.float32)
    scores = torch.tensor(rng.random(n).astype("float32"))
    keep = amg_ref.nms(boxes, scores, thr)
    ks = scores[keep]
    assert int(keep[0]) == int(torch.argmax(scores)) and bool((ks[:-1] >= ks[1:]).all())
    again = amg_ref.nms(boxes[keep], scores[keep], thr)
    assert again.tolist() == list(range(len(keep)))


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 70), st.integers(1, 70), st.integers(0, 2 ** 31 - 1), st.floats(0.0, 1.0))
def test_coco_rle_round_trip(h, w, seed, density):
    """COCO compressed RLE string (cocoapi rleToString / rleFrString): decode(encode(rle)) == rle, printable ASCII only."""
    rng = np.random.default_rng(seed)
    mask = rng.random((1, h, w)) < density
    rle = au.mask_to_rle(mask)[0]
    enc = au.coco_encode_rle(rle)
    assert enc["size"] == [h, w] and all(48 <= ord(c) < 48 + 64 for c in enc["counts"])
    assert au.coco_decode_rle(enc)["counts"] == rle["counts"]


def test_coco_rle_known_values():
    """Hand-derived from the cocoapi algorithm: small counts are single characters ('0' + value), counts from the fourth on
    are stored as differences to counts[i-2], values >= 16 need a continuation group, negative differences set bit 4."""
    assert au.coco_encode_rle({"size": [2, 3], "counts": [3, 2, 1]})["counts"] == "321"
    assert au.coco_encode_rle({"size": [1, 19], "counts": [5, 3, 7, 4]})["counts"] == "5371"       # 4 - 3 = 1
    assert au.coco_encode_rle({"size": [1, 40], "counts": [40]})["counts"] == "X1"                # 40 = 8 + 32*1
    assert au.coco_encode_rle({"size": [1, 30], "counts": [0, 10, 5, 7]})["counts"] == "0:5M"      # 7 - 10 = -3 -> 'M'
