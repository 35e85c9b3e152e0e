"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the integer / host side of micro-sam's
automatic-mask-generation and batched-inference path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this.

Follows (reference file:line):
* ``segment_anything.utils.amg`` (third party, not in /root/reference; call sites listed in SURVEY.md A.5):
  MaskData, build_point_grid, generate_crop_boxes, batch_iterator, calculate_stability_score,
  is_box_near_crop_edge, uncrop_*, rle_to_mask, area_from_rle, box_xyxy_to_xywh
* micro_sam/_vendored.py:33-85 (batched_mask_to_box), :104-152 (mask_to_rle_pytorch, numpy implementation)
* micro_sam/util.py:618-651 (_to_image), :654-681 (_compute_embeddings_batched), :902-917 (_compute_2d),
  :1589-1676 (mask NMS), :1773-1848 (mask_data_to_segmentation)
* micro_sam/instance_segmentation.py:99-144 (_postprocess_batch), :188-255 (_postprocess_masks, _to_mask_data),
  :356-530 (AutomaticMaskGenerator)
* micro_sam/inference.py:137-286 (batched_inference)
* torchvision.ops.boxes.batched_nms semantics (greedy, descending score, suppress IoU > thr)

Pinned against the reference's own known-answer tests in tests/test_oracle_golden.py
(test/test_vendored.py:12-25 box [3,7,4,8]; :44-78 RLE properties; test/test_util.py:81-104 tiled-NMS border masks).
"""
from __future__ import annotations

import math
from copy import deepcopy
from itertools import product
from typing import Any, Dict, Generator, List, Optional, Tuple

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ amg utils
class MaskData:
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

    def items(self):
        return self._stats.items()

    def filter(self, keep: torch.Tensor) -> None:
        for k, v in self._stats.items():
            if v is None:
                self._stats[k] = None
            elif isinstance(v, torch.Tensor):
                self._stats[k] = v[torch.as_tensor(keep, device=v.device)]
            elif isinstance(v, np.ndarray):
                self._stats[k] = v[keep.detach().cpu().numpy()]
            elif isinstance(v, list) and keep.dtype == torch.bool:
                self._stats[k] = [a for i, a in enumerate(v) if keep[i]]
            elif isinstance(v, list):
                self._stats[k] = [v[i] for i in keep]
            else:
                raise TypeError(f"MaskData key {k} has an unsupported type {type(v)}.")

    def cat(self, new_stats: "MaskData") -> None:
        for k, v in new_stats.items():
            if k not in self._stats or self._stats[k] is None:
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
                self._stats[k] = v.float().detach().cpu().numpy() if v.dtype == torch.bfloat16 else v.detach().cpu().numpy()


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
        crop_box_x0 = [int((crop_w - overlap) * i) for i in range(n_crops_per_side)]
        crop_box_y0 = [int((crop_h - overlap) * i) for i in range(n_crops_per_side)]
        for x0, y0 in product(crop_box_x0, crop_box_y0):
            crop_boxes.append([x0, y0, min(x0 + crop_w, im_w), min(y0 + crop_h, im_h)])
            layer_idxs.append(i_layer + 1)
    return crop_boxes, layer_idxs


def batch_iterator(batch_size: int, *args) -> Generator[List[Any], None, None]:
    assert len(args) > 0 and all(len(a) == len(args[0]) for a in args)
    n_batches = len(args[0]) // batch_size + int(len(args[0]) % batch_size != 0)
    for b in range(n_batches):
        yield [arg[b * batch_size: (b + 1) * batch_size] for arg in args]


def calculate_stability_score(masks: torch.Tensor, mask_threshold, threshold_offset: float) -> torch.Tensor:
    intersections = (masks > (mask_threshold + threshold_offset)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    unions = (masks > (mask_threshold - threshold_offset)).sum(-1, dtype=torch.int16).sum(-1, dtype=torch.int32)
    return intersections / unions


def uncrop_boxes_xyxy(boxes: torch.Tensor, crop_box) -> torch.Tensor:
    x0, y0, _, _ = crop_box
    offset = torch.tensor([[x0, y0, x0, y0]], device=boxes.device)
    if len(boxes.shape) == 3:
        offset = offset.unsqueeze(1)
    return boxes + offset


def uncrop_points(points: torch.Tensor, crop_box) -> torch.Tensor:
    x0, y0, _, _ = crop_box
    offset = torch.tensor([[x0, y0]], device=points.device)
    if len(points.shape) == 3:
        offset = offset.unsqueeze(1)
    return points + offset


def uncrop_masks(masks: torch.Tensor, crop_box, orig_h: int, orig_w: int) -> torch.Tensor:
    x0, y0, x1, y1 = crop_box
    if x0 == 0 and y0 == 0 and x1 == orig_w and y1 == orig_h:
        return masks
    pad_x, pad_y = orig_w - (x1 - x0), orig_h - (y1 - y0)
    return torch.nn.functional.pad(masks, (x0, pad_x - x0, y0, pad_y - y0), value=0)


def is_box_near_crop_edge(boxes: torch.Tensor, crop_box, orig_box, atol: float = 20.0) -> torch.Tensor:
    crop_box_torch = torch.as_tensor(crop_box, dtype=torch.float, device=boxes.device)
    orig_box_torch = torch.as_tensor(orig_box, dtype=torch.float, device=boxes.device)
    boxes = uncrop_boxes_xyxy(boxes, crop_box).float()
    near_crop_edge = torch.isclose(boxes, crop_box_torch[None, :], atol=atol, rtol=0)
    near_image_edge = torch.isclose(boxes, orig_box_torch[None, :], atol=atol, rtol=0)
    near_crop_edge = torch.logical_and(near_crop_edge, ~near_image_edge)
    return torch.any(near_crop_edge, dim=1)


def remove_small_regions(mask: np.ndarray, area_thresh: float, mode: str):
    """segment_anything.utils.amg.remove_small_regions (third party, unpinned; call site instance_segmentation.py:157-160).
    Upstream uses cv2.connectedComponentsWithStats(working_mask, 8); cv2 is absent here, scipy.ndimage.label with the full
    3x3 structure is the same 8-connected labelling in the same raster order of first pixels (parity unpinned: no cv2)."""
    from scipy import ndimage
    assert mode in ("holes", "islands")
    correct_holes = mode == "holes"
    working_mask = (correct_holes ^ mask).astype(np.uint8)
    regions, n_labels = ndimage.label(working_mask, structure=np.ones((3, 3), dtype=int))
    n_labels += 1  # cv2 counts the background label 0
    sizes = np.bincount(regions.ravel(), minlength=n_labels)[1:]
    small_regions = [i + 1 for i, s_ in enumerate(sizes) if s_ < area_thresh]
    if len(small_regions) == 0:
        return mask, False
    fill_labels = [0] + small_regions
    if not correct_holes:
        fill_labels = [i for i in range(n_labels) if i not in fill_labels]
        if len(fill_labels) == 0:  # every region is below the threshold: keep the largest
            fill_labels = [int(np.argmax(sizes)) + 1]
    mask = np.isin(regions, fill_labels)
    return mask, True


def box_xyxy_to_xywh(box_xyxy):
    box_xywh = deepcopy(box_xyxy)
    box_xywh[2] = box_xywh[2] - box_xywh[0]
    box_xywh[3] = box_xywh[3] - box_xywh[1]
    return box_xywh


def rle_to_mask(rle: Dict[str, Any]) -> np.ndarray:
    h, w = rle["size"]
    mask = np.empty(h * w, dtype=bool)
    idx, parity = 0, False
    for count in rle["counts"]:
        mask[idx: idx + count] = parity
        idx += count
        parity ^= True
    return mask.reshape(w, h).transpose()


def area_from_rle(rle: Dict[str, Any]) -> int:
    return sum(rle["counts"][1::2])


# ------------------------------------------------------------------------------------------------ _vendored.py
def batched_mask_to_box(masks: torch.Tensor) -> torch.Tensor:
    """_vendored.py:33-85."""
    assert masks.dtype == torch.bool, masks.dtype
    if torch.numel(masks) == 0:
        return torch.zeros(*masks.shape[:-2], 4, device=masks.device)
    shape = masks.shape
    h, w = shape[-2:]
    masks = masks.flatten(0, -3) if len(shape) > 2 else masks.unsqueeze(0)
    in_height, _ = torch.max(masks, dim=-1)
    in_height_coords = in_height * torch.arange(h, dtype=torch.int)[None, :]
    bottom_edges, _ = torch.max(in_height_coords, dim=-1)
    in_height_coords = (in_height_coords + h * (~in_height)).type(torch.int)
    top_edges, _ = torch.min(in_height_coords, dim=-1)
    in_width, _ = torch.max(masks, dim=-2)
    in_width_coords = in_width * torch.arange(w, dtype=torch.int)[None, :]
    right_edges, _ = torch.max(in_width_coords, dim=-1)
    in_width_coords = (in_width_coords + w * (~in_width)).type(torch.int)
    left_edges, _ = torch.min(in_width_coords, dim=-1)
    empty_filter = (right_edges < left_edges) | (bottom_edges < top_edges)
    out = torch.stack([left_edges, top_edges, right_edges, bottom_edges], dim=-1)
    out = out * (~empty_filter).unsqueeze(-1)
    return out.reshape(*shape[:-2], 4) if len(shape) > 2 else out[0]


def compute_rle_numpy(mask: np.ndarray) -> List[int]:
    """_vendored.py:104-111."""
    diffs = mask[1:] != mask[:-1]
    indices = np.append(np.where(diffs), len(mask) - 1)
    counts = [] if mask[0] == 0 else [0]
    counts += np.diff(np.append(-1, indices)).tolist()
    return counts


def mask_to_rle(tensor: torch.Tensor) -> List[Dict[str, Any]]:
    """_vendored.py:114-152 (Fortran-order run lengths)."""
    b, h, w = tensor.shape
    flat = tensor.permute(0, 2, 1).flatten(1).detach().cpu().numpy()
    return [{"size": [h, w], "counts": compute_rle_numpy(m)} for m in flat]


# ------------------------------------------------------------------------------------------------ box NMS
def box_area(boxes: torch.Tensor) -> torch.Tensor:
    return (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.nms semantics, fp32 arithmetic: greedy over descending score, suppress IoU > thr."""
    boxes = boxes.to(torch.float32).cpu().numpy()
    order = torch.argsort(scores.float().cpu(), descending=True, stable=True).numpy()
    areas = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    suppressed = np.zeros(len(boxes), dtype=bool)
    keep = []
    for _i, i in enumerate(order):
        if suppressed[i]:
            continue
        keep.append(i)
        rest = order[_i + 1:]
        xx1 = np.maximum(boxes[i, 0], boxes[rest, 0])
        yy1 = np.maximum(boxes[i, 1], boxes[rest, 1])
        xx2 = np.minimum(boxes[i, 2], boxes[rest, 2])
        yy2 = np.minimum(boxes[i, 3], boxes[rest, 3])
        w = np.maximum(np.float32(0), xx2 - xx1)
        h = np.maximum(np.float32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (areas[i] + areas[rest] - inter)
        suppressed[rest[ovr > np.float32(iou_threshold)]] = True
    return torch.as_tensor(np.array(keep, dtype=np.int64))


def batched_nms(boxes, scores, idxs, iou_threshold):
    assert torch.all(idxs == idxs.flatten()[0]) if idxs.numel() else True  # the reference always passes one category
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    return nms(boxes, scores, iou_threshold)


# ------------------------------------------------------------------------------------------------ util.py
def to_image(image: np.ndarray) -> np.ndarray:
    """util.py:618-651."""
    input_ = image
    ndim = input_.ndim
    n_channels = 1 if ndim == 2 else input_.shape[-1]
    if ndim == 2:
        input_ = np.concatenate([input_[..., None]] * 3, axis=-1)
    elif ndim == 3 and n_channels == 1:
        input_ = np.concatenate([input_] * 3, axis=-1)
    elif ndim == 3 and n_channels == 2:
        zero_channel = np.zeros(input_.shape[:2] + (1,), dtype=input_.dtype)
        input_ = np.concatenate([input_, zero_channel], axis=-1)
    elif ndim == 3 and n_channels == 3:
        pass
    elif ndim == 3 and n_channels > 3:
        input_ = input_[..., :3]
    else:
        raise ValueError(f"Invalid input dimensionality {ndim}.")
    input_ = input_.astype("float32")
    input_ -= input_.min(axis=(0, 1))[None, None]
    input_ /= (input_.max(axis=(0, 1))[None, None] + 1e-7)
    input_ = (input_ * 255).astype("uint8")
    return np.array(input_)


@torch.no_grad()
def compute_embeddings_batched(predictor, batched_images):
    """util.py:654-681."""
    predictor.reset_image()
    tensors, original_sizes, input_sizes = [], [], []
    for image in batched_images:
        t = predictor.transform.apply_image(image)
        t = torch.as_tensor(t, device=predictor.device).permute(2, 0, 1).contiguous()[None]
        original_sizes.append(image.shape[:2])
        input_sizes.append(tuple(t.shape[-2:]))
        tensors.append(predictor.model.preprocess(t))
    features = predictor.model.image_encoder(torch.cat(tensors))
    predictor.original_size = original_sizes[-1]
    predictor.input_size = input_sizes[-1]
    predictor.features = features[-1]
    predictor.is_image_set = True
    return features, original_sizes, input_sizes


def precompute_image_embeddings_2d(predictor, image: np.ndarray) -> Dict[str, Any]:
    """util.py:902-917 (_compute_2d, in-memory branch of precompute_image_embeddings :1133)."""
    predictor.reset_image()
    predictor.set_image(to_image(image))
    features = predictor.get_image_embedding().cpu().numpy()
    return {"features": features, "input_size": predictor.input_size, "original_size": image.shape[:2]}


def set_precomputed(predictor, image_embeddings: Dict[str, Any]):
    """util.py:1215-1258, 2-D non-tiled branch."""
    features = image_embeddings["features"]
    predictor.features = torch.as_tensor(np.asarray(features), device=predictor.device)
    predictor.original_size = tuple(image_embeddings["original_size"])
    predictor.input_size = tuple(image_embeddings["input_size"])
    predictor.is_image_set = True
    return predictor


def label_connected(seg: np.ndarray) -> np.ndarray:
    """elf.parallel.label on a label image: connected components per label value, 4-connectivity (elf default
    connectivity=1), background 0 kept.  Component ids are assigned in raster order of first occurrence.
    (Per label value one ndimage.label inside its bounding box with a running id offset, then ONE renumbering pass by first
    raster position -- linear in pixels x values; a per-component Python loop is quadratic on noisy label images.)"""
    from scipy import ndimage
    tmp = np.zeros(seg.shape, dtype=np.int64)
    base = 0
    for val, sl in enumerate(ndimage.find_objects(seg.astype(np.int64)), start=1):
        if sl is None:
            continue
        lab, n = ndimage.label(seg[sl] == val)
        view = tmp[sl]
        view[lab > 0] = lab[lab > 0] + base
        base += n
    flat = tmp.ravel()
    ids, first = np.unique(flat, return_index=True)
    ids, first = ids[ids != 0], first[ids != 0]
    lut = np.zeros(base + 1, dtype=seg.dtype)
    lut[ids[np.argsort(first, kind="stable")]] = np.arange(1, len(ids) + 1, dtype=seg.dtype)
    return lut[tmp]


def mask_data_to_segmentation(masks: List[Dict[str, Any]], shape=None, min_object_size: int = 0,
                              max_object_size: Optional[int] = None, label_masks: bool = True,
                              with_background: bool = False, merge_exclusively: bool = True) -> np.ndarray:
    """util.py:1773-1848."""
    masks = sorted(masks, key=(lambda x: x["area"]), reverse=True)
    if shape is None:
        shape = next(iter(masks))["segmentation"].shape
    segmentation = np.zeros(shape, dtype="uint32")
    seg_id = 1
    for mask_data in masks:
        area = mask_data["area"]
        if (area < min_object_size) or (max_object_size is not None and area > max_object_size):
            continue
        this_mask = mask_data["segmentation"]
        this_mask = this_mask.cpu().numpy() if torch.is_tensor(this_mask) else this_mask
        this_seg_id = mask_data.get("seg_id", seg_id)
        if "global_bbox" in mask_data:  # tiled records: paint the local box window at its global position (:1814-1823)
            bb = mask_data["bbox"]
            bb = np.s_[bb[1]:bb[1] + bb[3], bb[0]:bb[0] + bb[2]]
            gbb = mask_data["global_bbox"]
            gbb = np.s_[gbb[1]:gbb[1] + gbb[3], gbb[0]:gbb[0] + gbb[2]]
            this_mask = np.logical_and(this_mask[bb], segmentation[gbb] == 0) if merge_exclusively else this_mask[bb]
            segmentation[gbb][this_mask] = this_seg_id
        else:
            if merge_exclusively:
                this_mask = np.logical_and(this_mask, segmentation == 0)
            segmentation[this_mask] = this_seg_id
        seg_id = this_seg_id + 1
    if label_masks:
        segmentation = label_connected(segmentation)
    seg_ids, sizes = np.unique(segmentation, return_counts=True)
    filter_ids = seg_ids[sizes < min_object_size]
    if with_background:
        bg_id = seg_ids[np.argmax(sizes)]
        filter_ids = np.concatenate([filter_ids, [bg_id]])
    segmentation[np.isin(segmentation, filter_ids)] = 0
    # relabel consecutive, keeping 0
    ids = np.unique(segmentation)
    ids = ids[ids != 0]
    lut = np.zeros(int(segmentation.max()) + 1, dtype=segmentation.dtype)
    lut[ids] = np.arange(1, len(ids) + 1, dtype=segmentation.dtype)
    return lut[segmentation]


# ------------------------------------------------------------------------------------------------ AMG
class AutomaticMaskGenerator:
    """instance_segmentation.py:65-530 (AMGBase + AutomaticMaskGenerator)."""

    def __init__(self, predictor, points_per_side: int = 32, points_per_batch: int = 64, stability_score_offset=1.0,
                 crop_n_layers: int = 0, crop_overlap_ratio: float = 512 / 1500, crop_n_points_downscale_factor: int = 1):
        self.point_grids = build_all_layer_point_grids(points_per_side, crop_n_layers, crop_n_points_downscale_factor)
        self._crop_n_layers, self._crop_overlap_ratio = crop_n_layers, crop_overlap_ratio
        self._predictor = predictor
        self._points_per_batch = points_per_batch
        self._stability_score_offset = stability_score_offset
        self._is_initialized = False

    def _to_mask_data(self, masks, iou_preds, crop_box, original_size, points=None):
        orig_h, orig_w = original_size
        data = MaskData(masks=masks.flatten(0, 1), iou_preds=iou_preds.flatten(0, 1))
        if points is not None:
            data["points"] = torch.as_tensor(points.repeat(masks.shape[1], axis=0), dtype=torch.float)
        thr = self._predictor.model.mask_threshold
        data["stability_score"] = calculate_stability_score(data["masks"], thr, self._stability_score_offset)
        data["masks"] = (data["masks"] > thr).type(torch.bool)
        data["boxes"] = batched_mask_to_box(data["masks"])
        data["masks"] = uncrop_masks(data["masks"], crop_box, orig_h, orig_w)
        data["rles"] = mask_to_rle(data["masks"])
        del data["masks"]
        return data

    def _process_batch(self, points, im_size, crop_box, original_size):
        transformed_points = self._predictor.transform.apply_coords(points, im_size)
        in_points = torch.as_tensor(transformed_points, device=self._predictor.device, dtype=torch.float)
        in_labels = torch.ones(in_points.shape[0], dtype=torch.int, device=in_points.device)
        masks, iou_preds, _ = self._predictor.predict_torch(
            point_coords=in_points[:, None, :], point_labels=in_labels[:, None], multimask_output=True,
            return_logits=True)
        return self._to_mask_data(masks, iou_preds, crop_box, original_size, points=points)

    @torch.no_grad()
    def initialize(self, image: np.ndarray, image_embeddings=None):
        original_size = image.shape[:2]
        self._original_size = original_size
        crop_boxes, layer_idxs = generate_crop_boxes(original_size, self._crop_n_layers, self._crop_overlap_ratio)
        precomputed = len(crop_boxes) == 1   # :433-441: with several crops every crop is embedded on its own
        if precomputed:
            if image_embeddings is None:
                image_embeddings = precompute_image_embeddings_2d(self._predictor, image)
            set_precomputed(self._predictor, image_embeddings)
        image = to_image(image)
        crop_list = []
        for crop_box, layer_idx in zip(crop_boxes, layer_idxs):
            x0, y0, x1, y1 = crop_box
            cropped_im = image[y0:y1, x0:x1, :]
            cropped_im_size = cropped_im.shape[:2]
            if not precomputed:
                self._predictor.set_image(cropped_im)
            points_scale = np.array(cropped_im_size)[None, ::-1]
            points_for_image = self.point_grids[layer_idx] * points_scale
            data = MaskData()
            for (points,) in batch_iterator(self._points_per_batch, points_for_image):
                data.cat(self._process_batch(points, cropped_im_size, crop_box, original_size))
            crop_list.append(data)
        self._is_initialized = True
        self._crop_list = crop_list
        self._crop_boxes = crop_boxes

    def _postprocess_batch(self, data, crop_box, original_size, pred_iou_thresh, stability_score_thresh, box_nms_thresh):
        orig_h, orig_w = original_size
        if pred_iou_thresh > 0.0:
            data.filter(data["iou_preds"] > pred_iou_thresh)
        if stability_score_thresh > 0.0:
            data.filter(data["stability_score"] >= stability_score_thresh)
        keep_mask = ~is_box_near_crop_edge(data["boxes"], crop_box, [0, 0, orig_w, orig_h])
        if not torch.all(keep_mask):
            data.filter(keep_mask)
        keep_by_nms = batched_nms(data["boxes"].float(), data["iou_preds"], torch.zeros_like(data["boxes"][:, 0]),
                                  iou_threshold=box_nms_thresh)
        data.filter(keep_by_nms)
        data["boxes"] = uncrop_boxes_xyxy(data["boxes"], crop_box)
        data["crop_boxes"] = torch.tensor([crop_box for _ in range(len(data["rles"]))])
        data["points"] = uncrop_points(data["points"], crop_box)
        return data

    @torch.no_grad()
    def generate(self, pred_iou_thresh=0.88, stability_score_thresh=0.95, box_nms_thresh=0.7, crop_nms_thresh=0.7,
                 min_mask_region_area=0, output_mode="instance_segmentation", with_background=True):
        if not self._is_initialized:
            raise RuntimeError("AutomaticMaskGenerator has not been initialized. Call initialize first.")
        data = MaskData()
        for data_, crop_box in zip(self._crop_list, self._crop_boxes):
            data.cat(self._postprocess_batch(deepcopy(data_), crop_box, self._original_size, pred_iou_thresh,
                                             stability_score_thresh, box_nms_thresh))
        if len(self._crop_boxes) > 1 and len(data["crop_boxes"]) > 0:  # prefer masks from smaller crops (:511-521)
            scores = 1 / box_area(data["crop_boxes"])
            keep_by_nms = batched_nms(data["boxes"].float(), scores, torch.zeros_like(data["boxes"][:, 0]),
                                      iou_threshold=crop_nms_thresh)
            data.filter(keep_by_nms)
        data.to_numpy()
        if min_mask_region_area > 0 and len(data["rles"]) > 0:   # _postprocess_small_regions, :146-186
            nms_thresh = max(box_nms_thresh, crop_nms_thresh)
            new_masks, scores = [], []
            for rle in data["rles"]:
                mask = rle_to_mask(rle)
                mask, changed = remove_small_regions(mask, min_mask_region_area, mode="holes")
                unchanged = not changed
                mask, changed = remove_small_regions(mask, min_mask_region_area, mode="islands")
                unchanged = unchanged and not changed
                new_masks.append(torch.as_tensor(mask, dtype=torch.int).unsqueeze(0))
                scores.append(float(unchanged))   # NMS prefers masks that did not need post-processing
            masks = torch.cat(new_masks, dim=0)
            boxes = batched_mask_to_box(masks.to(torch.bool))
            keep_by_nms = batched_nms(boxes.float(), torch.as_tensor(scores, dtype=torch.float), torch.zeros_like(boxes[:, 0]),
                                      iou_threshold=nms_thresh)
            for i_mask in keep_by_nms:
                if scores[i_mask] == 0.0:
                    data["rles"][i_mask] = mask_to_rle(masks[i_mask].unsqueeze(0).to(torch.bool))[0]
                    data["boxes"][i_mask] = boxes[i_mask]
            data.filter(keep_by_nms)
        if output_mode in ("binary_mask", "instance_segmentation"):
            segs = [rle_to_mask(rle) for rle in data["rles"]]
        elif output_mode == "rle":
            segs = data["rles"]
        else:
            raise ValueError(f"Invalid output mode {output_mode}.")
        anns = []
        for idx in range(len(segs)):
            anns.append({
                "segmentation": segs[idx],
                "area": area_from_rle(data["rles"][idx]),
                "bbox": box_xyxy_to_xywh(data["boxes"][idx]).tolist(),
                "predicted_iou": data["iou_preds"][idx].item(),
                "stability_score": data["stability_score"][idx].item(),
                "crop_box": box_xyxy_to_xywh(data["crop_boxes"][idx]).tolist(),
                "point_coords": [data["points"][idx].tolist()],
            })
        if output_mode == "instance_segmentation":
            shape = next(iter(anns))["segmentation"].shape if len(anns) > 0 else self._original_size
            return mask_data_to_segmentation(anns, shape=shape, with_background=with_background,
                                             merge_exclusively=False)
        return anns


# ------------------------------------------------------------------------------------------------ inference.py
def local_otsu_threshold(images: torch.Tensor, window_size: int = 31, num_bins: int = 64, eps: float = 1e-6) -> torch.Tensor:
    """inference.py:70-134, evaluated mask by mask (the reference unfolds the whole batch at once: same values, but
    (B, 961, 65536) fp32 temporaries).  images (B,1,H,W) -> thresholds (B,1,1)."""
    import torch.nn.functional as F
    out = []
    for b in range(images.shape[0]):
        x = images[b:b + 1].to(torch.float32)
        _, _, H, W = x.shape
        x_min, x_max = x.min().view(1, 1, 1, 1), x.max().view(1, 1, 1, 1)
        x_range = (x_max - x_min).clamp_min(eps)
        x_norm = (x - x_min) / x_range
        patches = F.unfold(x_norm, kernel_size=window_size, padding=window_size // 2)      # (1, P, L)
        bin_idx = (patches * (num_bins - 1)).long().clamp(0, num_bins - 1)
        L = bin_idx.shape[2]
        one_hot = torch.zeros(1, L, num_bins, dtype=torch.float32)
        idx = bin_idx.transpose(1, 2)
        one_hot.scatter_add_(2, idx, torch.ones_like(idx, dtype=one_hot.dtype))
        hist = one_hot.permute(0, 2, 1)
        p = hist / hist.sum(dim=1, keepdim=True).clamp_min(eps)
        bins = torch.arange(num_bins, dtype=torch.float32).view(1, num_bins, 1)
        omega1 = torch.cumsum(p, dim=1)
        mu = torch.cumsum(p * bins, dim=1)
        mu_T = mu[:, -1:, :]
        omega2 = 1.0 - omega1
        mu1 = mu / omega1.clamp_min(eps)
        mu2 = (mu_T - mu) / omega2.clamp_min(eps)
        sigma_b2 = omega1 * omega2 * (mu1 - mu2) ** 2
        t_bin = torch.argmax(sigma_b2, dim=1)
        t_norm = t_bin.to(torch.float32) / (num_bins - 1)
        thr_vals = (x_min.view(1, 1) + t_norm * x_range.view(1, 1)).clamp_min(0.0)
        out.append(torch.amax(thr_vals.view(1, H, W), dim=(1, 2), keepdim=True))
    return torch.cat(out)


@torch.no_grad()
def batched_inference(predictor, image, batch_size: int, boxes=None, points=None, point_labels=None,
                      multimasking: bool = False, embedding_path=None, return_instance_segmentation: bool = True,
                      image_embeddings=None, mask_threshold: float = 0.0, logits_masks=None):
    """inference.py:155-286 (default threshold path; `logits_masks` (N,1,256,256) are passed on as mask prompts :240-247)."""
    if boxes is None and points is None:
        raise ValueError("batched_inference needs boxes and/or points")
    if image_embeddings is None:
        image_embeddings = precompute_image_embeddings_2d(predictor, image)
    set_precomputed(predictor, image_embeddings)
    image_shape = tuple(image_embeddings["original_size"]) if image is None else image.shape[:2]
    have_points, have_boxes = points is not None, boxes is not None
    n_prompts = boxes.shape[0] if have_boxes else points.shape[0]
    if have_boxes:
        bx = predictor.transform.apply_boxes(boxes, image_shape)
        bx = torch.tensor(bx, dtype=torch.float32).to(predictor.device)
    if have_points:
        pt = predictor.transform.apply_coords(points, image_shape)
        pt = torch.tensor(pt, dtype=torch.float32).to(predictor.device)
        pl = torch.tensor(point_labels, dtype=torch.float32).to(predictor.device)
    masks = MaskData()
    n_batches = int(np.ceil(float(n_prompts) / batch_size))
    for b in range(n_batches):
        s, e = b * batch_size, (b + 1) * batch_size
        bm, bi, bl = predictor.predict_torch(
            point_coords=pt[s:e] if have_points else None, point_labels=pl[s:e] if have_points else None,
            boxes=bx[s:e] if have_boxes else None, mask_input=None if logits_masks is None else logits_masks[s:e],
            multimask_output=multimasking, return_logits=True)
        if multimasking:
            best = torch.argmax(bi, dim=1)
            sel = torch.arange(bm.shape[0])
            bm, bi, bl = bm[sel, best][:, None], bi[sel, best][:, None], bl[sel, best][:, None]
        data = MaskData(masks=bm.flatten(0, 1), iou_preds=bi.flatten(0, 1))
        thr_b = local_otsu_threshold(bl) if isinstance(mask_threshold, str) else mask_threshold   # inference.py:137-151
        data["stability_scores"] = calculate_stability_score(data["masks"], thr_b, 1.0)
        data["masks"] = (data["masks"] > thr_b).type(torch.bool)
        data["boxes"] = batched_mask_to_box(data["masks"])
        masks.cat(data)
    recs = [{
        "segmentation": mask, "area": mask.sum(), "bbox": box_xyxy_to_xywh(box).tolist(),
        "predicted_iou": iou.item(), "stability_score": sc.item(), "seg_id": idx,
    } for idx, (mask, box, iou, sc) in enumerate(zip(masks["masks"], masks["boxes"], masks["iou_preds"],
                                                      masks["stability_scores"]), 1)]
    if return_instance_segmentation:
        return mask_data_to_segmentation(recs, min_object_size=0, shape=image_shape)
    return recs


def stitch_segmentation(masks, tile_ids, tiling, halo, output_shape):
    """inference.py:337-356 (+ _merge_segmentations :315-332: `discard_ids` is computed there but never used, so the merge
    reduces to "the previous segmentation is fully preserved")."""
    segmentation = np.zeros(output_shape, dtype="uint32")
    for tile_id, this_seg in zip(tile_ids, masks):
        t = tiling.get_block_with_halo(tile_id, list(halo)).outer_block
        bb = tuple(slice(b, e) for b, e in zip(t.begin, t.end))
        if tile_id == 0:
            segmentation[bb] = this_seg
        else:
            prev = segmentation[bb]
            this_seg = this_seg.copy()
            captured = prev != 0
            this_seg[captured] = prev[captured]
            segmentation[bb] = this_seg
    return segmentation


def coordinates_to_block_id(tiling, coords) -> int:
    pos = [int((c - b) // s) for c, b, s in zip(coords, tiling.rb, tiling.bs)]
    return int(np.ravel_multi_index(pos, tiling.blocks_per_axis))


@torch.no_grad()
def batched_tiled_inference(predictor, image, batch_size: int, image_embeddings=None, boxes=None, points=None,
                            point_labels=None, multimasking: bool = False, return_instance_segmentation: bool = True,
                            tile_shape=None, halo=None, mask_threshold: float = 0.0):
    """inference.py:359-538, default (optimize_memory=False) path: prompts are routed to the tile that contains the box
    centre / the point, run per tile with tile-local coordinates, and the records get a `global_bbox`."""
    if image_embeddings is None:
        image_embeddings = precompute_tiled_embeddings_2d(predictor, image, tile_shape, halo)
    tile_shape, halo = image_embeddings["tile_shape"], image_embeddings["halo"]
    shape = image.shape[:2]
    tiling = Blocking([0, 0], shape, tile_shape)
    have_boxes, have_points = boxes is not None, points is not None
    n_prompts = boxes.shape[0] if have_boxes else points.shape[0]
    box_to_tile, point_to_tile, label_to_tile, tile_ids = {}, {}, {}, []
    for k in range(n_prompts):
        tid = None
        if have_boxes:
            box = boxes[k]
            center = np.array([(box[1] + box[3]) / 2, (box[0] + box[2]) / 2]).round().astype("int").tolist()
            tid = coordinates_to_block_id(tiling, center)
            t = tiling.get_block_with_halo(tid, list(halo)).outer_block
            off, ts = t.begin, t.shape
            b = np.array([max(box[1] - off[0], 0), max(box[0] - off[1], 0), min(box[3] - off[0], ts[0]),
                          min(box[2] - off[1], ts[1])])[None]
            box_to_tile[tid] = np.concatenate([box_to_tile[tid], b]) if tid in box_to_tile else b
        if have_points:
            pt = points[k, 0][::-1].round().astype("int").tolist()
            if tid is None:
                tid = coordinates_to_block_id(tiling, pt)
            t = tiling.get_block_with_halo(tid, list(halo)).outer_block
            pin = (points[k, 0] - np.array(t.begin)[::-1])[None, None]
            lin = point_labels[k][None]
            point_to_tile[tid] = np.concatenate([point_to_tile[tid], pin]) if tid in point_to_tile else pin
            label_to_tile[tid] = np.concatenate([label_to_tile[tid], lin]) if tid in label_to_tile else lin
        tile_ids.append(tid)
    tile_ids = sorted(set(tile_ids))
    masks = []
    for tid in tile_ids:
        recs = batched_inference(predictor, None, batch_size, boxes=box_to_tile.get(tid), points=point_to_tile.get(tid),
                                 point_labels=label_to_tile.get(tid), multimasking=multimasking,
                                 return_instance_segmentation=False, image_embeddings=image_embeddings["features"][str(tid)],
                                 mask_threshold=mask_threshold)
        t = tiling.get_block_with_halo(tid, list(halo)).outer_block
        offset = np.array(t.begin[::-1] + [0, 0])
        masks.extend({**m, "global_bbox": (np.array(m["bbox"]) + offset).tolist()} for m in recs)
    if return_instance_segmentation:
        return mask_data_to_segmentation(masks, shape=shape, min_object_size=0)
    return masks


# ------------------------------------------------------------------------------------------------ mask NMS (util.py)
def overlap_matrix(boxes: torch.Tensor) -> torch.Tensor:
    """util.py:1589-1598."""
    x1 = torch.max(boxes[:, None, 0], boxes[:, 0])
    y1 = torch.max(boxes[:, None, 1], boxes[:, 1])
    x2 = torch.min(boxes[:, None, 2], boxes[:, 2])
    y2 = torch.min(boxes[:, None, 3], boxes[:, 3])
    return (torch.clamp(x2 - x1, min=0) * torch.clamp(y2 - y1, min=0)) > 0


def ious_between_pred_masks(masks: torch.Tensor, boxes: torch.Tensor, diagonal_value=1) -> torch.Tensor:
    """util.py:1601-1619 (integer popcounts -> fp32 ratio; pairs without box overlap stay 0)."""
    n = masks.shape[0]
    flat = masks.reshape(n, -1).to(torch.float64)
    inter = flat @ flat.t()
    area = flat.sum(1)
    union = area[:, None] + area[None, :] - inter
    iou = (inter.to(torch.int64).to(torch.float32) / union.to(torch.int64).to(torch.float32))
    m = torch.where(overlap_matrix(boxes), iou, torch.zeros_like(iou))
    m = torch.triu(m, diagonal=1)
    m = m + m.T
    m.fill_diagonal_(diagonal_value)
    return m


def iomin_between_pred_masks(masks: torch.Tensor, boxes: torch.Tensor, eps=1e-6) -> torch.Tensor:
    """util.py:1622-1644."""
    n = masks.shape[0]
    flat = masks.reshape(n, -1).float()
    areas = flat.sum(dim=1)
    inter = flat @ flat.t()
    iomin = inter / (torch.minimum(areas[:, None], areas[None, :]) + eps)
    iomin[~overlap_matrix(boxes)] = 0
    return iomin


def batched_mask_nms(masks, boxes, scores, nms_thresh: float, intersection_over_min: bool) -> torch.Tensor:
    """util.py:1647-1676: greedy, descending score, keep `iou <= thresh` (non-strict)."""
    masks, boxes, scores = torch.as_tensor(masks).cpu(), torch.as_tensor(boxes).cpu(), torch.as_tensor(scores).cpu()
    mat = iomin_between_pred_masks(masks, boxes) if intersection_over_min else ious_between_pred_masks(masks, boxes)
    sorted_indices = torch.argsort(scores, descending=True)
    keep = []
    while len(sorted_indices) > 0:
        i = sorted_indices[0]
        keep.append(i)
        if len(sorted_indices) == 1:
            break
        sorted_indices = sorted_indices[1:][mat[i, sorted_indices[1:]] <= nms_thresh]
    return torch.tensor(keep)


def xywh_to_xyxy(boxes):
    """util.py:1679-1684."""
    boxes = boxes.clone() if isinstance(boxes, torch.Tensor) else torch.tensor(boxes)
    boxes[:, 2] += boxes[:, 0]
    boxes[:, 3] += boxes[:, 1]
    return boxes


def infer_tiled_shape(predictions):
    """util.py:1687-1695."""
    shape = [0, 0]
    for pred in predictions:
        bbox, gbb = pred["bbox"], pred["global_bbox"]
        offset = (gbb[0] - bbox[0], gbb[1] - bbox[1])
        ms = pred["segmentation"].shape
        shape[0] = max(shape[0], offset[1] + ms[0])
        shape[1] = max(shape[1], offset[0] + ms[1])
    return tuple(int(v) for v in shape)


def tiled_mask_overlap_matrix(masks, boxes, global_boxes, intersection_over_min: bool) -> torch.Tensor:
    """util.py:1698-1747: pairwise IoU / IoMin of tile-local masks, evaluated on the overlap window of their global boxes."""
    n = len(masks)
    boxes = torch.as_tensor(np.asarray(boxes)).to(torch.long)
    global_boxes = torch.as_tensor(np.asarray(global_boxes)).to(torch.long)
    gxyxy = xywh_to_xyxy(global_boxes).to(torch.long)
    ovl = overlap_matrix(gxyxy)
    masks = [torch.as_tensor(np.asarray(m)) for m in masks]
    areas = torch.tensor([float(m.sum()) for m in masks], dtype=torch.float32)
    out = torch.zeros((n, n))
    for i in range(n):
        js = torch.where(ovl[i])[0]
        off_i = global_boxes[i, :2] - boxes[i, :2]
        for j in js[js > i]:
            off_j = global_boxes[j, :2] - boxes[j, :2]
            o = [max(gxyxy[i, 0], gxyxy[j, 0]), max(gxyxy[i, 1], gxyxy[j, 1]), min(gxyxy[i, 2], gxyxy[j, 2]),
                 min(gxyxy[i, 3], gxyxy[j, 3])]
            mi = masks[i][o[1] - off_i[1]:o[3] - off_i[1], o[0] - off_i[0]:o[2] - off_i[0]]
            mj = masks[j][o[1] - off_j[1]:o[3] - off_j[1], o[0] - off_j[0]:o[2] - off_j[0]]
            inter = torch.logical_and(mi, mj).sum()
            den = torch.minimum(areas[i], areas[j]) if intersection_over_min else areas[i] + areas[j] - inter
            out[i, j] = inter / den
    out = out + out.T
    out.fill_diagonal_(1)
    return out


def batched_tiled_mask_nms(masks, boxes, global_boxes, scores, nms_thresh: float, intersection_over_min: bool) -> torch.Tensor:
    """util.py:1750-1770."""
    scores = torch.as_tensor(np.asarray(scores))
    mat = tiled_mask_overlap_matrix(masks, boxes, global_boxes, intersection_over_min)
    order = torch.argsort(scores, descending=True)
    keep = []
    while len(order) > 0:
        i = order[0]
        keep.append(int(i))
        if len(order) == 1:
            break
        order = order[1:][mat[i, order[1:]] <= nms_thresh]
    return torch.tensor(keep)


def apply_nms(predictions, min_size: int, shape=None, perform_box_nms: bool = False, nms_thresh: float = 0.9,
              max_size=None, intersection_over_min: bool = False) -> np.ndarray:
    """util.py:1851-1957."""
    is_tiled = "global_bbox" in predictions[0]
    if is_tiled and shape is None:
        shape = infer_tiled_shape(predictions)
    masks = [torch.as_tensor(np.asarray(p["segmentation"].cpu() if torch.is_tensor(p["segmentation"]) else p["segmentation"]))
             for p in predictions]
    idx = list(range(len(predictions)))
    area = [int(m.sum()) for m in masks]
    if min_size > 0:
        idx = [k for k in idx if area[k] > min_size]
    if max_size is not None:
        idx = [k for k in idx if area[k] < max_size]
    if shape is None:
        shape = tuple(predictions[0]["segmentation"].shape)
    if len(idx) == 0:
        return np.zeros(shape, dtype="uint32")
    scores = torch.tensor([predictions[k]["predicted_iou"] * predictions[k]["stability_score"] for k in idx], dtype=torch.float32)
    boxes = torch.tensor(np.array([predictions[k]["bbox"] for k in idx]))
    gboxes = torch.tensor(np.array([predictions[k]["global_bbox"] for k in idx])) if is_tiled else None
    bxyxy = xywh_to_xyxy(gboxes if is_tiled else boxes)
    if perform_box_nms:
        keep = batched_nms(bxyxy.float(), scores, torch.zeros(len(idx)), nms_thresh)
    elif is_tiled:
        keep = batched_tiled_mask_nms([masks[k] for k in idx], boxes, gboxes, scores, nms_thresh, intersection_over_min)
    else:
        keep = batched_mask_nms(torch.stack([masks[k] for k in idx]), bxyxy, scores, nms_thresh, intersection_over_min)
    recs = []
    for q in keep.tolist():
        k = idx[q]
        rec = {"segmentation": masks[k], "area": area[k], "bbox": boxes[q].tolist()}
        if is_tiled:
            rec["global_bbox"] = gboxes[q].tolist()
        recs.append(rec)
    return mask_data_to_segmentation(recs, shape=shape, min_object_size=min_size)


# ------------------------------------------------------------------------------------------------ tiling
class Blocking:
    """bioimage_cpp.utils.Blocking (nifty semantics): row-major block ids, halo clipped to the ROI."""

    class _B:
        def __init__(self, begin, end):
            self.begin, self.end = list(begin), list(end)
            self.shape = [e - b for b, e in zip(begin, end)]

    def __init__(self, roi_begin, roi_end, block_shape):
        self.rb, self.re, self.bs = list(roi_begin), list(roi_end), list(block_shape)
        self.blocks_per_axis = [-(-(e - b) // s) for b, e, s in zip(self.rb, self.re, self.bs)]
        self.number_of_blocks = int(np.prod(self.blocks_per_axis))

    def get_block_with_halo(self, block_id, halo):
        pos = np.unravel_index(block_id, self.blocks_per_axis)
        ib = [b + p * s for b, p, s in zip(self.rb, pos, self.bs)]
        ie = [min(x + s, e) for x, s, e in zip(ib, self.bs, self.re)]
        ob = [max(x - h, b) for x, h, b in zip(ib, halo, self.rb)]
        oe = [min(x + h, e) for x, h, e in zip(ie, halo, self.re)]

        class R:
            pass
        r = R()
        r.inner_block, r.outer_block = Blocking._B(ib, ie), Blocking._B(ob, oe)
        r.inner_block_local = Blocking._B([a - b for a, b in zip(ib, ob)], [a - b for a, b in zip(ie, ob)])
        return r


def precompute_tiled_embeddings_2d(predictor, image, tile_shape, halo):
    """util.py:765-803 (_compute_tiled_features_2d): every outer tile is normalised on its own (_to_image)."""
    tiling = Blocking([0, 0], image.shape[:2], tile_shape)
    feats = {}
    for tile_id in range(tiling.number_of_blocks):
        t = tiling.get_block_with_halo(tile_id, list(halo)).outer_block
        tile = image[t.begin[0]:t.end[0], t.begin[1]:t.end[1]]
        feats[str(tile_id)] = precompute_image_embeddings_2d(predictor, tile)
    return {"features": feats, "tile_shape": tuple(tile_shape), "halo": tuple(halo)}


class TiledAutomaticMaskGenerator(AutomaticMaskGenerator):
    """instance_segmentation.py:564-680."""

    @torch.no_grad()
    def initialize(self, image, image_embeddings=None, tile_shape=None, halo=None):
        original_size = image.shape[:2]
        self._original_size = original_size
        if image_embeddings is None:
            image_embeddings = precompute_tiled_embeddings_2d(self._predictor, image, tile_shape, halo)
        tile_shape, halo = image_embeddings["tile_shape"], image_embeddings["halo"]
        tiling = Blocking([0, 0], original_size, tile_shape)
        tiles = [tiling.get_block_with_halo(t, list(halo)).outer_block for t in range(tiling.number_of_blocks)]
        crop_boxes = [[t.begin[1], t.begin[0], t.end[1], t.end[0]] for t in tiles]
        image = to_image(image)
        mask_data = []
        for tile_id, crop_box in enumerate(crop_boxes):
            set_precomputed(self._predictor, image_embeddings["features"][str(tile_id)])
            x0, y0, x1, y1 = crop_box
            cropped_im_size = image[y0:y1, x0:x1, :].shape[:2]
            points_for_image = self.point_grids[0] * np.array(cropped_im_size)[None, ::-1]
            data = MaskData()
            for (points,) in batch_iterator(self._points_per_batch, points_for_image):
                data.cat(self._process_batch(points, cropped_im_size, crop_box, original_size))
            mask_data.append(data)
        self._is_initialized = True
        self._crop_list = mask_data
        self._crop_boxes = crop_boxes
