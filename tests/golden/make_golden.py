"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE's own code in this container (run once here; the GPU box has
no /root/reference).  Third-party imports the reference needs but this image lacks are stubbed only where the stubbed
symbol is not on the executed path:
  * micro_sam/_vendored.py is imported whole with a stub `bioimage_cpp.utils.compute_rle` (the numpy RLE implementation
    of the same file, _vendored.py:104, is the one exercised: rle_implementation="numpy");
  * micro_sam/util.py cannot be imported (zarr/elf/pooch/...), so the *source text* of the listed functions is extracted
    with `ast` and exec'ed unchanged: _to_image, _overlap_matrix, _calculate_ious_between_pred_masks,
    _calculate_iomin_between_pred_masks, _batched_mask_nms, _xywh_to_xyxy, _infer_tiled_shape,
    _calculate_tiled_mask_overlap_matrix, _batched_tiled_mask_nms.
Usage:  python tests/golden/make_golden.py
"""
import ast
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch

REF = "/root/reference/micro_sam"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_vendored():
    stub = types.ModuleType("bioimage_cpp")
    stub.utils = types.ModuleType("bioimage_cpp.utils")
    stub.utils.compute_rle = lambda mask: (_ for _ in ()).throw(RuntimeError("stub: bioimage_cpp not available"))
    sys.modules["bioimage_cpp"] = stub
    sys.modules["bioimage_cpp.utils"] = stub.utils
    spec = importlib.util.spec_from_file_location("ref_vendored", os.path.join(REF, "_vendored.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_util_functions(names):
    src = open(os.path.join(REF, "util.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "torch": torch, "warnings": warnings}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), "util.py", "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def blobs(rng, n, h, w):
    yy, xx = np.mgrid[:h, :w]
    out = np.zeros((n, h, w), dtype=bool)
    for i in range(n):
        for _ in range(rng.integers(1, 4)):
            cy, cx, r = rng.integers(0, h), rng.integers(0, w), rng.integers(2, max(3, min(h, w) // 3))
            out[i] |= (yy - cy) ** 2 + (xx - cx) ** 2 < r * r
    return out


def otsu_inputs():
    """(3,1,256,256) fp32 low-res logits: smooth, smooth + noise, shifted negative (seeded; also used by the tests)."""
    g = torch.Generator().manual_seed(7)
    lo = torch.nn.functional.interpolate(torch.randn(3, 1, 12, 12, generator=g), (256, 256), mode="bicubic") * 4
    lo[1] += torch.randn(1, 256, 256, generator=g) * 0.5
    lo[2] = lo[2] - 6.0
    return lo


def main():
    rng = np.random.default_rng(0)
    ven = load_vendored()
    res = {}
    # --- batched_mask_to_box + mask_to_rle_pytorch (numpy implementation) on random blob masks, incl. empty / full
    for tag, (n, h, w) in {"a": (6, 128, 256), "b": (5, 37, 53), "c": (4, 64, 64)}.items():
        m = blobs(rng, n, h, w)
        m[0] = False
        m[1] = True
        t = torch.from_numpy(m)
        res[f"masks_{tag}"] = np.packbits(m, axis=-1)
        res[f"shape_{tag}"] = np.array([n, h, w])
        res[f"boxes_{tag}"] = ven.batched_mask_to_box(t).numpy()
        rles = ven.mask_to_rle_pytorch(t, rle_implementation="numpy")
        res[f"rle_counts_{tag}"] = np.concatenate([np.asarray(r["counts"], dtype=np.int64) for r in rles])
        res[f"rle_lens_{tag}"] = np.array([len(r["counts"]) for r in rles])
    np.savez_compressed(os.path.join(OUT, "vendored.npz"), **res)

    fns = load_util_functions(["_to_image", "_overlap_matrix", "_calculate_ious_between_pred_masks",
                               "_calculate_iomin_between_pred_masks", "_batched_mask_nms"])
    res = {}
    # --- _to_image
    imgs = {
        "gray_f32": rng.normal(100, 30, (40, 50)).astype("float32"),
        "gray_u16": rng.integers(0, 60000, (33, 47)).astype("uint16"),
        "one_ch": rng.random((20, 30, 1)).astype("float64"),
        "two_ch": rng.integers(0, 255, (25, 25, 2)).astype("uint8"),
        "rgb_u8": rng.integers(0, 255, (31, 29, 3)).astype("uint8"),
        "const": np.full((16, 16), 7.0, dtype="float32"),
    }
    for k, v in imgs.items():
        res[f"in_{k}"] = v
        res[f"out_{k}"] = fns["_to_image"](v)
    # --- mask NMS (IoU and IoMin), greedy order incl. ties
    m = blobs(rng, 24, 64, 64)
    m[5] = m[4]                      # exact duplicate
    m[7] = m[6] & (rng.random((64, 64)) > 0.02)  # near duplicate
    masks = torch.from_numpy(m)
    boxes = ven.batched_mask_to_box(masks)
    scores = torch.from_numpy(rng.random(24).astype("float32"))
    scores[5] = scores[4]
    res["nms_masks"] = np.packbits(m, axis=-1)
    res["nms_boxes"] = boxes.numpy()
    res["nms_scores"] = scores.numpy()
    for thr in (0.3, 0.9):
        res[f"nms_keep_iou_{thr}"] = fns["_batched_mask_nms"](masks, boxes, scores, thr, False).numpy()
        res[f"nms_keep_iomin_{thr}"] = fns["_batched_mask_nms"](masks, boxes, scores, thr, True).numpy()
    res["nms_iou_matrix"] = fns["_calculate_ious_between_pred_masks"](masks, boxes).numpy()
    np.savez_compressed(os.path.join(OUT, "util.npz"), **res)

    # --- tiled mask NMS (util.py:1687-1770): tile-local masks with `bbox` (local xywh) and `global_bbox`
    fns = load_util_functions(["_overlap_matrix", "_xywh_to_xyxy", "_infer_tiled_shape", "_calculate_tiled_mask_overlap_matrix",
                               "_batched_tiled_mask_nms"])
    rng = np.random.default_rng(1)
    res = {}
    shape, tile, halo = (96, 128), (64, 64), (16, 16)
    tiles = []
    for ty in range(0, shape[0], tile[0]):
        for tx in range(0, shape[1], tile[1]):
            y0, x0 = max(ty - halo[0], 0), max(tx - halo[1], 0)
            y1, x1 = min(ty + tile[0] + halo[0], shape[0]), min(tx + tile[1] + halo[1], shape[1])
            tiles.append((y0, x0, y1, x1))
    preds = []
    for k in range(18):
        y0, x0, y1, x1 = tiles[k % len(tiles)]
        m = blobs(rng, 1, y1 - y0, x1 - x0)[0]
        if k == 7:   # same object seen from the neighbouring tile: paste the overlap part of mask 6
            py0, px0, py1, px1 = tiles[6 % len(tiles)]
            g = np.zeros(shape, dtype=bool)
            g[py0:py1, px0:px1] = preds[6]["segmentation"]
            m = g[y0:y1, x0:x1].copy()
        if not m.any():
            m[3:9, 4:12] = True
        ys, xs = np.where(m)
        bbox = [int(xs.min()), int(ys.min()), int(xs.max() - xs.min()), int(ys.max() - ys.min())]  # xywh as batched_mask_to_box
        preds.append({"segmentation": m, "bbox": bbox, "global_bbox": [bbox[0] + x0, bbox[1] + y0, bbox[2], bbox[3]]})
    scores = rng.random(len(preds)).astype("float32")
    masks = [torch.from_numpy(p["segmentation"]) for p in preds]
    boxes = torch.tensor([p["bbox"] for p in preds])
    gboxes = torch.tensor([p["global_bbox"] for p in preds])
    res["n"] = np.array(len(preds))
    for k, p_ in enumerate(preds):
        res[f"mask_{k}"] = p_["segmentation"]
    res["boxes"], res["global_boxes"], res["scores"] = boxes.numpy(), gboxes.numpy(), scores
    res["inferred_shape"] = np.array(fns["_infer_tiled_shape"](preds))
    for iomin in (False, True):
        res[f"overlap_{int(iomin)}"] = fns["_calculate_tiled_mask_overlap_matrix"](masks, boxes, gboxes, iomin).numpy()
        for thr in (0.3, 0.9):
            res[f"keep_{int(iomin)}_{thr}"] = fns["_batched_tiled_mask_nms"](masks, boxes, gboxes, torch.from_numpy(scores), thr,
                                                                            iomin).numpy()
    np.savez_compressed(os.path.join(OUT, "tiled_nms.npz"), **res)

    # --- _merge_segmentations / _stitch_segmentation (inference.py:315-356), executed from the reference source.  The absent
    # bioimage_cpp.segmentation_overlap is stubbed with an object that reports a 100 % overlap for every id: the reference
    # collects such ids in `discard_ids` but never uses the list, so the stub cannot influence the result -- which is
    # exactly the behaviour the oracle / product restate ("the previous segmentation is fully preserved").
    class _Ovlp:
        def overlaps_for_label_a(self, seg_id, normalize=True):
            return {"label": np.array([0, 1]), "fraction": np.array([0.0, 1.0])}

    class _Blk:
        def __init__(self, b, e):
            self.begin, self.end = list(b), list(e)

    class _Tiling:  # nifty-style blocking with the camelCase API the reference calls
        def __init__(self, shape, block):
            self.shape, self.block = shape, block
            self.grid = [-(-s // b) for s, b in zip(shape, block)]

        def getBlockWithHalo(self, tile_id, halo):
            pos = np.unravel_index(tile_id, self.grid)
            ib = [p_ * b for p_, b in zip(pos, self.block)]
            ie = [min(x + b, s_) for x, b, s_ in zip(ib, self.block, self.shape)]
            outer = _Blk([max(x - h, 0) for x, h in zip(ib, halo)], [min(x + h, s_) for x, h, s_ in zip(ie, halo, self.shape)])
            return types.SimpleNamespace(outerBlock=outer)

    src = open(os.path.join(REF, "inference.py")).read()
    ns = {"np": np, "segmentation_overlap": lambda a, b: _Ovlp(), "tqdm": lambda it, **kw: it}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("_merge_segmentations", "_stitch_segmentation"):
            exec(compile(ast.Module(body=[node], type_ignores=[]), "inference.py", "exec"), ns)
    rng = np.random.default_rng(2)
    shape, tile, halo = (300, 420), (160, 224), (24, 24)
    tiling = _Tiling(shape, tile)
    res = {"shape": np.array(shape), "tile_shape": np.array(tile), "halo": np.array(halo)}
    for name, ids in (("all", [0, 1, 2, 3]), ("subset", [0, 2, 3]), ("no_first", [1, 3])):
        segs = []
        for t in ids:
            ob = tiling.getBlockWithHalo(t, list(halo)).outerBlock
            segs.append(rng.integers(0, 5, (ob.end[0] - ob.begin[0], ob.end[1] - ob.begin[1])).astype("uint32") * (t + 1))
        res[f"{name}_ids"] = np.array(ids)
        for k, sg in enumerate(segs):
            res[f"{name}_seg{k}"] = sg
        res[f"{name}_out"] = ns["_stitch_segmentation"]([sg.copy() for sg in segs], ids, tiling, halo, shape)
    np.savez_compressed(os.path.join(OUT, "stitch.npz"), **res)

    # --- _local_otsu_threshold (inference.py:70-134, mask_threshold="auto"), executed from the reference source on CPU.
    # Inputs are regenerated from the seed by the tests (otsu_inputs below); their float64 sum is stored to detect drift.
    import torch.nn.functional as F
    ns = {"np": np, "torch": torch, "F": F}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "_local_otsu_threshold":
            exec(compile(ast.Module(body=[node], type_ignores=[]), "inference.py", "exec"), ns)
    x = otsu_inputs()
    thr = ns["_local_otsu_threshold"](x)
    np.savez_compressed(os.path.join(OUT, "otsu.npz"), thresholds=thr.reshape(-1).numpy(), checksum=np.array(x.double().sum().item()))
    print("written", os.listdir(OUT))


if __name__ == "__main__":
    main()
