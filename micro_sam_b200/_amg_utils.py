"""Host-side helpers of the AMG / batched-inference path (the role `segment_anything.utils.amg` and
`micro_sam._vendored` play for the reference; call sites in SURVEY.md A.5).  Pure bookkeeping on small arrays -- all
per-pixel work lives in csrc/postprocess.cu."""
from __future__ import annotations

import math
from copy import deepcopy
from itertools import product
from typing import Any, Dict, Generator, List

import numpy as np
import torch


class MaskData:
    """Dict of per-mask lists / arrays / tensors with batched filter & cat (instance_segmentation.py:233,385,500)."""

    def __init__(self, **kwargs):
        for v in kwargs.values():
            assert isinstance(v, (list, np.ndarray, torch.Tensor))
        self._stats = dict(**kwargs)

    def __setitem__(self, key, item):
        assert isinstance(item, (list, np.ndarray, torch.Tensor))
        self._stats[key] = item

    def __delitem__(self, key):
        del self._stats[key]

    def __getitem__(self, key):
        return self._stats[key]

    def __contains__(self, key):
        return key in self._stats

    def keys(self):
        return self._stats.keys()

    def items(self):
        return self._stats.items()

    def filter(self, keep: torch.Tensor) -> None:
        for k, v in self._stats.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                self._stats[k] = v[torch.as_tensor(keep, device=v.device)]
            elif isinstance(v, np.ndarray):
                self._stats[k] = v[keep.detach().cpu().numpy()]
            elif isinstance(v, list) and keep.dtype == torch.bool:
                self._stats[k] = [a for i, a in enumerate(v) if keep[i]]
            elif isinstance(v, list):
                self._stats[k] = [v[i] for i in keep]
            else:
                raise TypeError(f"MaskData key {k} has an unsupported type {type(v)}.")

    def cat(self, new_stats: "MaskData", copy: bool = True) -> None:
        """copy=False adopts the first batch's tensors instead of cloning them (the AMG path hands over freshly computed
        per-batch tensors; cloning [3072, 256, 256] fp32 low-res logits is an 805 MB device copy per tile)."""
        for k, v in new_stats.items():
            if k not in self._stats or self._stats[k] is None:
                if isinstance(v, torch.Tensor):
                    self._stats[k] = v.clone() if copy else v
                else:
                    self._stats[k] = deepcopy(v)
            elif isinstance(v, torch.Tensor):
                self._stats[k] = torch.cat([self._stats[k], v], dim=0)
            elif isinstance(v, np.ndarray):
                self._stats[k] = np.concatenate([self._stats[k], v], axis=0)
            elif isinstance(v, list):
                self._stats[k] = self._stats[k] + deepcopy(v)
            else:
                raise TypeError(f"MaskData key {k} has an unsupported type {type(v)}.")

    def to_numpy(self) -> None:
        for k, v in self._stats.items():
            if isinstance(v, torch.Tensor):
                self._stats[k] = v.detach().cpu().numpy()


def build_point_grid(n_per_side: int) -> np.ndarray:
    offset = 1 / (2 * n_per_side)
    pts = np.linspace(offset, 1 - offset, n_per_side)
    px = np.tile(pts[None, :], (n_per_side, 1))
    py = np.tile(pts[:, None], (1, n_per_side))
    return np.stack([px, py], axis=-1).reshape(-1, 2)


def build_all_layer_point_grids(n_per_side: int, n_layers: int, scale_per_layer: int) -> List[np.ndarray]:
    return [build_point_grid(int(n_per_side / (scale_per_layer ** i))) for i in range(n_layers + 1)]


def generate_crop_boxes(im_size, n_layers: int, overlap_ratio: float):
    crop_boxes, layer_idxs = [], []
    im_h, im_w = im_size
    short_side = min(im_h, im_w)
    crop_boxes.append([0, 0, im_w, im_h])
    layer_idxs.append(0)

    def crop_len(orig_len, n_crops, overlap):
        return int(math.ceil((overlap * (n_crops - 1) + orig_len) / n_crops))

    for i_layer in range(n_layers):
        n_crops_per_side = 2 ** (i_layer + 1)
        overlap = int(overlap_ratio * short_side * (2 / n_crops_per_side))
        crop_w = crop_len(im_w, n_crops_per_side, overlap)
        crop_h = crop_len(im_h, n_crops_per_side, overlap)
        xs = [int((crop_w - overlap) * i) for i in range(n_crops_per_side)]
        ys = [int((crop_h - overlap) * i) for i in range(n_crops_per_side)]
        for x0, y0 in product(xs, ys):
            crop_boxes.append([x0, y0, min(x0 + crop_w, im_w), min(y0 + crop_h, im_h)])
            layer_idxs.append(i_layer + 1)
    return crop_boxes, layer_idxs


def batch_iterator(batch_size: int, *args) -> Generator[List[Any], None, None]:
    assert len(args) > 0 and all(len(a) == len(args[0]) for a in args)
    n_batches = len(args[0]) // batch_size + int(len(args[0]) % batch_size != 0)
    for b in range(n_batches):
        yield [arg[b * batch_size: (b + 1) * batch_size] for arg in args]


def box_xyxy_to_xywh(box_xyxy):
    box_xywh = deepcopy(box_xyxy)
    box_xywh[2] = box_xywh[2] - box_xywh[0]
    box_xywh[3] = box_xywh[3] - box_xywh[1]
    return box_xywh


def compute_rle(mask_fortran_flat: np.ndarray) -> List[int]:
    """Run lengths of a flat 0/1 vector, leading 0 if it starts with 1 (_vendored.py:104-111)."""
    m = mask_fortran_flat
    diffs = m[1:] != m[:-1]
    indices = np.append(np.where(diffs), len(m) - 1)
    counts = [] if m[0] == 0 else [0]
    counts += np.diff(np.append(-1, indices)).tolist()
    return counts


def mask_to_rle(masks: np.ndarray) -> List[Dict[str, Any]]:
    """(b,h,w) bool -> uncompressed column-major RLE dicts (_vendored.py:114-152)."""
    b, h, w = masks.shape
    if b == 0:
        return []
    flat = np.ascontiguousarray(masks.transpose(0, 2, 1)).reshape(b, -1)
    return [{"size": [h, w], "counts": compute_rle(m)} for m in flat]


def rle_to_mask(rle: Dict[str, Any]) -> np.ndarray:
    h, w = rle["size"]
    mask = np.empty(h * w, dtype=bool)
    idx, parity = 0, False
    for count in rle["counts"]:
        mask[idx: idx + count] = parity
        idx += count
        parity ^= True
    return mask.reshape(w, h).transpose()


def coco_encode_rle(uncompressed_rle: Dict[str, Any]) -> Dict[str, Any]:
    """segment_anything.utils.amg.coco_encode_rle (call site instance_segmentation.py:192): the upstream function hands the
    uncompressed column-major RLE to pycocotools (`frPyObjects`) and decodes the bytes to str.  pycocotools is absent here;
    this is its `rleToString` (cocoapi common/maskApi.c): counts[i] (i > 2: minus counts[i-2]) as little-endian 5-bit groups,
    bit 5 = continuation, + 48 -> ASCII.  Parity unpinned (no pycocotools vector in this image); round trip tested."""
    counts = uncompressed_rle["counts"]
    out = []
    for i, c in enumerate(counts):
        x = int(c)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            ch = x & 0x1F
            x >>= 5
            more = (x != -1) if (ch & 0x10) else (x != 0)
            if more:
                ch |= 0x20
            out.append(chr(ch + 48))
    h, w = uncompressed_rle["size"]
    return {"size": [h, w], "counts": "".join(out)}


def coco_decode_rle(rle: Dict[str, Any]) -> Dict[str, Any]:
    """Inverse of coco_encode_rle (cocoapi rleFrString): compressed string -> uncompressed counts."""
    s = rle["counts"]
    counts: List[int] = []
    p = 0
    while p < len(s):
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1F) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return {"size": list(rle["size"]), "counts": counts}


def area_from_rle(rle: Dict[str, Any]) -> int:
    return sum(rle["counts"][1::2])


class Blocking:
    """nifty/bioimage_cpp-style regular blocking with halo (util.py:766,857; inference.py:343-466)."""

    class _Block:
        def __init__(self, begin, end):
            self.begin, self.end = list(begin), list(end)
            self.shape = [e - b for b, e in zip(begin, end)]

    class _BlockWithHalo:
        def __init__(self, outer, inner, inner_local):
            self.outer_block = self.outerBlock = outer
            self.inner_block = self.innerBlock = inner
            self.inner_block_local = self.innerBlockLocal = inner_local

    def __init__(self, roi_begin, roi_end, block_shape):
        self.roi_begin, self.roi_end, self.block_shape = list(roi_begin), list(roi_end), list(block_shape)
        self.blocks_per_axis = [int(math.ceil((e - b) / s)) for b, e, s in zip(roi_begin, roi_end, block_shape)]
        self.number_of_blocks = int(np.prod(self.blocks_per_axis))
        self.numberOfBlocks = self.number_of_blocks

    def block_grid_position(self, block_id: int):
        return list(np.unravel_index(block_id, self.blocks_per_axis))

    blockGridPosition = block_grid_position

    def get_block(self, block_id: int):
        pos = self.block_grid_position(block_id)
        begin = [rb + p * s for rb, p, s in zip(self.roi_begin, pos, self.block_shape)]
        end = [min(b + s, re) for b, s, re in zip(begin, self.block_shape, self.roi_end)]
        return Blocking._Block(begin, end)

    getBlock = get_block

    def get_block_with_halo(self, block_id: int, halo):
        inner = self.get_block(block_id)
        ob = [max(b - h, rb) for b, h, rb in zip(inner.begin, halo, self.roi_begin)]
        oe = [min(e + h, re) for e, h, re in zip(inner.end, halo, self.roi_end)]
        outer = Blocking._Block(ob, oe)
        local = Blocking._Block([b - o for b, o in zip(inner.begin, ob)], [e - o for e, o in zip(inner.end, ob)])
        return Blocking._BlockWithHalo(outer, inner, local)

    getBlockWithHalo = get_block_with_halo

    def coordinates_to_block_id(self, coords) -> int:
        pos = [int((c - rb) // s) for c, rb, s in zip(coords, self.roi_begin, self.block_shape)]
        return int(np.ravel_multi_index(pos, self.blocks_per_axis))
