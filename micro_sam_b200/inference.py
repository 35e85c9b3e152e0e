"""micro_sam.inference.batched_inference (inference.py:155-286) on the B200 core.

Same signature and return values; the per-batch `predict_torch` + `_process_masks_for_batch` (:137-151, three passes
over (P,1,H,W) fp32 logits) + per-mask `.item()` record building + CPU painting are replaced by: decode -> fused
`msam_mask_stats` -> `msam_paint` (exclusive, descending area) on the device; only the uint32 label image is copied back.
"""
from __future__ import annotations

import ctypes
from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch

from . import _amg_utils as amg_utils
from . import _lib, util
from .sam import ResizeLongestSide, mask_stats


def _validate_inputs(boxes, points, point_labels, multimasking, return_instance_segmentation, segmentation_ids,
                     logits_masks):
    """inference.py:22-67."""
    if multimasking and (segmentation_ids is not None) and (not return_instance_segmentation):
        raise NotImplementedError
    if (points is None) != (point_labels is None):
        raise ValueError("If you have point prompts both `points` and `point_labels` have to be passed, "
                         "but you passed only one of them.")
    have_points, have_boxes, have_logits = points is not None, boxes is not None, logits_masks is not None
    if (not have_points) and (not have_boxes):
        raise ValueError("Point and/or box prompts have to be passed, you passed neither.")
    if have_points and (len(point_labels) != len(points)):
        raise ValueError(f"The number of point coordinates and labels does not match: {len(point_labels)} != {len(points)}")
    if (have_points and have_boxes) and (len(points) != len(boxes)):
        raise ValueError(f"The number of point and box prompts does not match: {len(points)} != {len(boxes)}")
    n_prompts = boxes.shape[0] if have_boxes else points.shape[0]
    if (segmentation_ids is not None) and (len(segmentation_ids) != n_prompts):
        raise ValueError(f"The number of segmentation ids and prompts does not match: {len(segmentation_ids)} != {n_prompts}")
    return n_prompts, have_boxes, have_points, have_logits


@torch.no_grad()
def batched_inference(predictor, image: Optional[np.ndarray], batch_size: int, boxes: Optional[np.ndarray] = None,
                      points: Optional[np.ndarray] = None, point_labels: Optional[np.ndarray] = None,
                      multimasking: bool = False, embedding_path=None, return_instance_segmentation: bool = True,
                      segmentation_ids: Optional[list] = None, reduce_multimasking: bool = True,
                      logits_masks: Optional[torch.Tensor] = None, verbose_embeddings: bool = False,
                      mask_threshold: Optional[Union[float, str]] = None, return_highres_logits: bool = False,
                      i: Optional[int] = None):
    n_prompts, have_boxes, have_points, have_logits = _validate_inputs(
        boxes, points, point_labels, multimasking, return_instance_segmentation, segmentation_ids, logits_masks)
    if have_logits:
        raise NotImplementedError("logits_masks (mask prompts) are not supported by the B200 decoder yet")
    if mask_threshold == "auto":
        raise NotImplementedError("mask_threshold='auto' (local Otsu, inference.py:70-134) is not on the B200 path")
    if multimasking and not reduce_multimasking:
        raise NotImplementedError("multimasking without reduce_multimasking")
    if image is None:
        predictor.get_image_embedding()
    else:
        input_ = image if i is None else image[i]
        emb = util.precompute_image_embeddings(predictor, input_, embedding_path, verbose=verbose_embeddings, to_numpy=False)
        util.set_precomputed(predictor, emb)
    device = predictor.device
    tf = ResizeLongestSide(1024)
    image_shape = predictor.original_size
    if have_boxes:
        boxes_t = torch.tensor(tf.apply_boxes(boxes, image_shape), dtype=torch.float32).to(device)
    if have_points:
        points_t = torch.tensor(tf.apply_coords(points, image_shape), dtype=torch.float32).to(device)
        labels_t = torch.tensor(point_labels, dtype=torch.float32).to(device)
    thr = predictor.model.mask_threshold if mask_threshold is None else float(mask_threshold)

    lows, ious = [], []
    for s in range(0, n_prompts, batch_size):
        e = min(s + batch_size, n_prompts)
        low, iou = predictor.decode_low_res(points_t[s:e] if have_points else None, labels_t[s:e] if have_points else None,
                                            boxes_t[s:e] if have_boxes else None, multimask_output=multimasking)
        if multimasking:  # keep the mask with the highest predicted IoU (inference.py:259-263)
            best = iou.argmax(dim=1)
            sel = torch.arange(low.shape[0], device=device)
            low, iou = low[sel, best][:, None], iou[sel, best][:, None]
        lows.append(low[:, 0])
        ious.append(iou[:, 0])
    low = torch.cat(lows).contiguous()
    iou = torch.cat(ious)
    bxs, stab, area = mask_stats(low, predictor.input_size, image_shape, thr, 1.0)
    H, W = image_shape
    inp = predictor.input_size
    seg_ids = np.arange(1, n_prompts + 1) if segmentation_ids is None else np.asarray(segmentation_ids, dtype=np.int64)

    if return_instance_segmentation:
        # mask_data_to_segmentation(masks, min_object_size=0): descending area (stable), first painter wins, CC, relabel
        order = torch.argsort(area, descending=True, stable=True)
        sel = order.to(torch.int32).contiguous()
        ids = torch.as_tensor(seg_ids, device=device)[order].to(torch.int32).contiguous()
        label = torch.zeros(H, W, dtype=torch.int32, device=device)
        _lib.check(_lib.lib().msam_paint(_lib.ptr(low), _lib.ptr(sel), _lib.ptr(bxs), _lib.ptr(ids), n_prompts, int(inp[0]),
                                         int(inp[1]), H, W, float(thr), 1, _lib.ptr(label), W, _lib.cur_stream()))
        return util._finish_segmentation(label.cpu().numpy().astype(np.uint32), min_object_size=0, label_masks=True,
                                         with_background=False)

    binm = torch.empty(n_prompts, H, W, dtype=torch.uint8, device=device)
    logits = torch.empty(n_prompts, H, W, dtype=torch.float32, device=device) if return_highres_logits else None
    _lib.check(_lib.lib().msam_upsample_masks(_lib.ptr(low), None, n_prompts, int(inp[0]), int(inp[1]), H, W, float(thr),
                                              _lib.ptr(logits), _lib.ptr(binm), _lib.cur_stream()))
    binm = binm.bool()
    bx, io, st, ar = bxs.cpu().numpy(), iou.cpu().numpy(), stab.cpu().numpy(), area.cpu().numpy()
    return [{
        "segmentation": binm[k], "area": int(ar[k]), "bbox": amg_utils.box_xyxy_to_xywh(bx[k].astype(np.int64)).tolist(),
        "predicted_iou": float(io[k]), "stability_score": float(st[k]), "seg_id": int(seg_ids[k]),
        "logits": (logits[k][None] if return_highres_logits else low[k][None]),
    } for k in range(n_prompts)]
