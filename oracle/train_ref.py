"""ORACLE (test infrastructure, NOT product code) -- fp32 PyTorch restatement of the forward half of micro-sam's fine-tuning
step: `TrainableSAM` (micro_sam/training/trainable_sam.py:12-114) and `SamTrainer._compute_iou / _compute_loss`
(micro_sam/training/sam_trainer.py:122-172) on the oracle `Sam` of oracle/sam_ref.py.

The dice term is torch_em's `DiceLoss(reduce_channel=None)` (third party, not under /root/reference, not installed):
channel-wise `1 - 2 sum(x y) / max(sum x^2 + sum y^2, eps)` with eps = 1e-7 over the flattened (N, H, W) samples of each
channel -- restated from torch_em.loss.dice; PARITY UNPINNED for that formula (no fixture in the reference).  Everything is
differentiable here (autograd), which is what the wgrad check in tests/ uses as its reference.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .sam_ref import Sam, get_preprocess_shape


class TrainableSAM(torch.nn.Module):
    def __init__(self, sam: Sam):
        super().__init__()
        self.sam = sam
        self.img_size = sam.image_encoder.img_size

    def preprocess(self, x):
        th, tw = get_preprocess_shape(x.shape[2], x.shape[3], self.img_size)
        x = F.interpolate(x, (th, tw), mode="bilinear", align_corners=False, antialias=True)   # apply_image_torch
        input_size = x.shape[-2:]
        x = (x - self.sam.pixel_mean.unsqueeze(0)) / self.sam.pixel_std.unsqueeze(0)
        return F.pad(x, (0, self.img_size - x.shape[-1], 0, self.img_size - x.shape[-2])), input_size

    def image_embeddings_oft(self, batched_inputs):
        images, input_size = self.preprocess(torch.stack([x["image"] for x in batched_inputs], dim=0))
        for rec in batched_inputs:
            rec["input_size"] = input_size
        return self.sam.image_encoder(images), batched_inputs

    def forward(self, batched_inputs, image_embeddings, multimask_output=False):
        outputs = []
        for rec, emb in zip(batched_inputs, image_embeddings):
            points = (rec["point_coords"], rec["point_labels"]) if "point_coords" in rec else None
            sparse, dense = self.sam.prompt_encoder(points=points, boxes=rec.get("boxes"), masks=rec.get("mask_inputs"))
            low, iou = self.sam.mask_decoder(image_embeddings=emb.unsqueeze(0), image_pe=self.sam.prompt_encoder.get_dense_pe(),
                                             sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense,
                                             multimask_output=multimask_output)
            masks = self.sam.postprocess_masks(low, input_size=rec["input_size"], original_size=rec["original_size"])
            outputs.append({"low_res_masks": low, "masks": masks, "iou_predictions": iou})
        return outputs


def dice_loss_per_channel(x, y, eps=1e-7):
    """torch_em.loss.DiceLoss(channelwise=True, reduce_channel=None): x, y (N, C, H, W) -> (C,)."""
    xf = x.transpose(0, 1).flatten(1)
    yf = y.transpose(0, 1).flatten(1).to(xf.dtype)
    num = (xf * yf).sum(-1)
    den = (xf * xf).sum(-1) + (yf * yf).sum(-1)
    return 1.0 - 2.0 * (num / den.clamp(min=eps))


def compute_iou(pred, true, eps=1e-7):
    """sam_trainer.py:122-129."""
    pm = pred > 0.5
    overlap = pm.logical_and(true).sum(dim=(1, 2, 3))
    union = pm.logical_or(true).sum(dim=(1, 2, 3))
    return overlap / (union + eps)


def compute_loss(batched_outputs, y_one_hot):
    """sam_trainer.py:131-172."""
    mask_loss, iou_loss = 0.0, 0.0
    for out, targets in zip(batched_outputs, y_one_hot):
        pred = torch.sigmoid(out["masks"])
        tb = targets.bool()
        dice = torch.stack([dice_loss_per_channel(pred[:, i:i + 1].swapaxes(0, 1), targets.swapaxes(0, 1))
                            for i in range(pred.shape[1])])
        dice, _ = torch.min(dice, dim=0)
        with torch.no_grad():
            true_iou = torch.stack([compute_iou(pred[:, i:i + 1], tb) for i in range(pred.shape[1])])
        iou_loss = iou_loss + F.mse_loss(true_iou.swapaxes(0, 1), out["iou_predictions"])
        mask_loss = mask_loss + torch.mean(dice)
    n = len(batched_outputs)
    mask_loss, iou_loss = mask_loss / n, iou_loss / n
    return mask_loss + iou_loss, mask_loss, iou_loss
