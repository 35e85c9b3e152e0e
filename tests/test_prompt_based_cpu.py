"""CPU tests of the interactive-segmentation host glue (f1; micro_sam/prompt_based_segmentation.py:28-506) with a recording
stand-in predictor: prompt conventions (row/col -> XY, box python -> XYXY), mask -> logits / box / points derivation, routing
of prompts to the tile of tiled embeddings and pasting the tile mask back."""
import numpy as np
import pytest
import torch

from micro_sam_b200 import prompt_based_segmentation as pbs
from micro_sam_b200.sam import ResizeLongestSide


class _Rec:
    """predict() returns a mask that marks the prompt locations so the geometry can be checked."""

    def __init__(self, size=(100, 140)):
        self.original_size = size
        self.calls = []
        self.transform = ResizeLongestSide(1024)
        self.device = torch.device("cpu")
        self.is_image_set = True

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output=True, return_logits=False):
        self.calls.append(dict(point_coords=point_coords, point_labels=point_labels, box=box, mask_input=mask_input,
                               multimask_output=multimask_output, return_logits=return_logits))
        M = 3 if multimask_output else 1
        h, w = self.original_size
        m = np.zeros((M, h, w), bool)
        if box is not None:
            m[:, box[1]:box[3], box[0]:box[2]] = True
        if point_coords is not None:
            for x, y in np.asarray(point_coords).astype(int):
                m[:, y, x] = True
        return m, np.arange(M, dtype="float32")[::-1].copy() if M == 3 else np.array([0.5], "float32"), np.zeros((M, 256, 256), "float32")


def test_points_box_conventions():
    p = _Rec()
    m = pbs.segment_from_points(p, np.array([[10, 20]]), np.array([1]))
    c = p.calls[-1]
    assert bool(c["multimask_output"]) and np.array_equal(c["point_coords"], [[20, 10]])     # single positive point -> best of 3
    assert m.shape == (1, 100, 140) and m[0, 10, 20]
    pbs.segment_from_points(p, np.array([[10, 20], [30, 40]]), np.array([1, 0]))
    assert not p.calls[-1]["multimask_output"]
    m = pbs.segment_from_box(p, np.array([5, 7, 50, 90]))
    assert np.array_equal(p.calls[-1]["box"], [7, 5, 90, 50]) and m[0, 5:50, 7:90].all() and m.sum() == 45 * 83
    pbs.segment_from_box(p, np.array([5, 7, 50, 90]), box_extension=0.1)
    assert np.array_equal(p.calls[-1]["box"], np.round([7 - 8.3, 5 - 4.5, 90 + 8.3, 50 + 4.5]).clip(0).astype(int))
    out = pbs.segment_from_box_and_points(p, np.array([5, 7, 50, 90]), np.array([[10, 20]]), np.array([1]), return_all=True)
    assert len(out) == 3 and np.array_equal(p.calls[-1]["point_coords"], [[20, 10]])


def test_mask_prompt_derivation():
    p = _Rec((64, 96))
    mask = np.zeros((64, 96), "uint8")
    mask[20:40, 30:70] = 1
    pbs.segment_from_mask(p, mask, use_points=True)
    c = p.calls[-1]
    assert np.array_equal(c["box"], [30, 20, 70, 40])
    lg = c["mask_input"]
    assert lg.shape == (1, 256, 256) and sorted(round(float(v), 3) for v in np.unique(lg)) == [-6.907, 6.907]
    # longest side 96 -> 256: the object covers rows 20..40 -> ~53..107, cols 30..70 -> 80..187; below row 171 is zero padding
    assert lg[0, 80, 130] > 0 and lg[0, 10, 10] < 0 and (lg[0, 172:, :] < 0).all()
    pts, lbl = c["point_coords"], c["point_labels"]
    assert (lbl == 1).sum() >= 1 and len(pts) == len(lbl)
    for (x, y), l in zip(pts, lbl):
        assert bool(mask[int(y), int(x)]) == bool(l)      # positives inside, negatives outside the object
    yy, xx = np.mgrid[:64, :96]
    disk = ((yy - 30) ** 2 + (xx - 50) ** 2 < 15 ** 2).astype("uint8")
    pbs.segment_from_mask(p, disk, use_points=True, use_single_point=True, use_box=False, use_mask=False)
    c = p.calls[-1]
    assert c["box"] is None and c["mask_input"] is None and len(c["point_coords"]) == 1
    x, y = c["point_coords"][0]
    assert abs(x - 50) <= 2 and abs(y - 30) <= 2     # centre of the distance transform
    with pytest.raises(ValueError):
        pbs.segment_from_mask(p, mask, points=np.array([[1, 1]]))
    empty = pbs.segment_from_mask(p, np.zeros((64, 96), "uint8"))
    assert p.calls[-1]["box"] is None and empty.shape == (1, 64, 96)


def test_tiled_routing_and_paste(monkeypatch):
    p = _Rec((272, 272))

    class F(dict):
        attrs = {"shape": (512, 512), "tile_shape": (256, 256), "halo": (16, 16)}
    seen = {}

    def fake_set_precomputed(predictor, emb, i=None, tile_id=None):
        seen["tile_id"] = tile_id
    monkeypatch.setattr(pbs.util, "set_precomputed", fake_set_precomputed)
    emb = {"features": F(), "input_size": None, "original_size": None}
    m = pbs.segment_from_points(p, np.array([[300, 40], [310, 60]]), np.array([1, 1]), image_embeddings=emb)
    assert seen["tile_id"] == 2                        # rows 256.. -> second tile row, first column
    assert np.array_equal(p.calls[-1]["point_coords"], [[40, 300 - 240], [60, 310 - 240]])   # tile 2 starts at row 240 (halo 16)
    assert m.shape == (1, 512, 512) and m[0, 300, 40] and m[0, 310, 60] and m.sum() == 2
    m = pbs.segment_from_box(p, np.array([260, 300, 400, 500]), image_embeddings=emb)
    assert seen["tile_id"] == 3 and m[0, 260:400, 300:500].all() and not m[0, :240].any()
    with pytest.warns(UserWarning):
        pbs.segment_from_points(p, np.array([[300, 40], [5, 5]]), np.array([1, 1]), image_embeddings=emb)
    with pytest.raises(RuntimeError):
        pbs.segment_from_box_and_points(p, np.array([260, 300, 400, 500]), np.array([[10, 10]]), np.array([1]), image_embeddings=emb)
