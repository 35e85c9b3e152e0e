"""ORACLE (test infrastructure, NOT product code) -- fp32 CPU restatement of the SAM arithmetic.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product (``micro_sam_b200``) never imports it.

What it restates
----------------
micro-sam delegates every FLOP of its hot path to the third-party package ``segment_anything``
(unpinned, ``/root/reference/setup.cfg:58``; imported at ``micro_sam/util.py:39-43``), which is NOT
under ``/root/reference`` and not installed in this image.  This file restates that published
algorithm (``segment_anything.modeling.{ImageEncoderViT,PromptEncoder,MaskDecoder,TwoWayTransformer,
Sam}``, ``segment_anything.predictor.SamPredictor``, ``segment_anything.utils.transforms.
ResizeLongestSide``) with upstream state-dict key names, anchored on the reference's own statements:

* architecture constants:            micro_sam/models/build_sam.py:40-142
* encoder block order / window pad:  micro_sam/models/sam_3d_wrapper.py:161-172,203-250
* qkv layout ([q|k|v] on last dim):  micro_sam/models/peft_sam.py:96-110
* preprocess / decoder protocol:     micro_sam/training/trainable_sam.py:24-114
* predictor protocol:                micro_sam/util.py:654-681, micro_sam/inference.py:248-255

PARITY PINNING: the reference's tests hold no numeric fixture for embeddings / logits / IoU
predictions -> for those values this oracle is pinned against a second independent
implementation of the same arithmetic (HuggingFace ``transformers.models.sam``, see
``tests/test_oracle_vs_hf.py``) and is otherwise "parity unpinned" w.r.t. real checkpoints.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# Architecture table (micro_sam/models/build_sam.py:40-76).  "vit_test*" are tiny shapes used only
# by the tests so the CPU oracle finishes in seconds.
# ----------------------------------------------------------------------------------------------
ARCH = {
    "vit_b": dict(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11)),
    "vit_l": dict(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23)),
    "vit_h": dict(embed_dim=1280, depth=32, num_heads=16, global_attn_indexes=(7, 15, 23, 31)),
    "vit_test": dict(embed_dim=128, depth=2, num_heads=2, global_attn_indexes=(1,)),
    "vit_test80": dict(embed_dim=160, depth=2, num_heads=2, global_attn_indexes=(1,)),
}


class LayerNorm2d(nn.Module):
    def __init__(self, num_channels: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps

    def forward(self, x):
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class MLPBlock(nn.Module):
    def __init__(self, dim, mlp_dim, act=nn.GELU):
        super().__init__()
        self.lin1 = nn.Linear(dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, dim)
        self.act = act()

    def forward(self, x):
        return self.lin2(self.act(self.lin1(x)))


# ------------------------------------ image encoder ------------------------------------------
def window_partition(x, ws):
    B, H, W, C = x.shape
    pad_h = (ws - H % ws) % ws
    pad_w = (ws - W % ws) % ws
    if pad_h > 0 or pad_w > 0:
        x = F.pad(x, (0, 0, 0, pad_w, 0, pad_h))
    Hp, Wp = H + pad_h, W + pad_w
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C)
    windows = x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)
    return windows, (Hp, Wp)


def window_unpartition(windows, ws, pad_hw, hw):
    Hp, Wp = pad_hw
    H, W = hw
    B = windows.shape[0] // (Hp * Wp // ws // ws)
    x = windows.view(B, Hp // ws, Wp // ws, ws, ws, -1)
    x = x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, Hp, Wp, -1)
    if Hp > H or Wp > W:
        x = x[:, :H, :W, :].contiguous()
    return x


def get_rel_pos(q_size, k_size, rel_pos):
    max_rel_dist = int(2 * max(q_size, k_size) - 1)
    if rel_pos.shape[0] != max_rel_dist:
        rel_pos_resized = F.interpolate(
            rel_pos.reshape(1, rel_pos.shape[0], -1).permute(0, 2, 1), size=max_rel_dist, mode="linear"
        )
        rel_pos_resized = rel_pos_resized.reshape(-1, max_rel_dist).permute(1, 0)
    else:
        rel_pos_resized = rel_pos
    q_coords = torch.arange(q_size)[:, None] * max(k_size / q_size, 1.0)
    k_coords = torch.arange(k_size)[None, :] * max(q_size / k_size, 1.0)
    relative_coords = (q_coords - k_coords) + (k_size - 1) * max(q_size / k_size, 1.0)
    return rel_pos_resized[relative_coords.long()]


def add_decomposed_rel_pos(attn, q, rel_pos_h, rel_pos_w, q_size, k_size):
    q_h, q_w = q_size
    k_h, k_w = k_size
    Rh = get_rel_pos(q_h, k_h, rel_pos_h)
    Rw = get_rel_pos(q_w, k_w, rel_pos_w)
    B, _, dim = q.shape
    r_q = q.reshape(B, q_h, q_w, dim)
    rel_h = torch.einsum("bhwc,hkc->bhwk", r_q, Rh)
    rel_w = torch.einsum("bhwc,wkc->bhwk", r_q, Rw)
    attn = (attn.view(B, q_h, q_w, k_h, k_w) + rel_h[:, :, :, :, None] + rel_w[:, :, :, None, :]).view(
        B, q_h * q_w, k_h * k_w
    )
    return attn


class Attention(nn.Module):
    def __init__(self, dim, num_heads, input_size):
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
        self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))

    def forward(self, x):
        B, H, W, _ = x.shape
        qkv = self.qkv(x).reshape(B, H * W, 3, self.num_heads, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.reshape(3, B * self.num_heads, H * W, -1).unbind(0)
        attn = (q * self.scale) @ k.transpose(-2, -1)
        # NB: the bias uses the UNSCALED q
        attn = add_decomposed_rel_pos(attn, q, self.rel_pos_h, self.rel_pos_w, (H, W), (H, W))
        attn = attn.softmax(dim=-1)
        x = (attn @ v).view(B, self.num_heads, H, W, -1).permute(0, 2, 3, 1, 4).reshape(B, H, W, -1)
        return self.proj(x)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio, window_size, input_size):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads, input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = MLPBlock(dim, int(dim * mlp_ratio))
        self.window_size = window_size

    def forward(self, x):
        shortcut = x
        x = self.norm1(x)
        if self.window_size > 0:
            H, W = x.shape[1], x.shape[2]
            # zero pad AFTER norm1: pad tokens carry q=k=v=qkv.bias (SURVEY §7 hard part 1)
            x, pad_hw = window_partition(x, self.window_size)
        x = self.attn(x)
        if self.window_size > 0:
            x = window_unpartition(x, self.window_size, pad_hw, (H, W))
        x = shortcut + x
        x = x + self.mlp(self.norm2(x))
        return x


class PatchEmbed(nn.Module):
    def __init__(self, patch, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).permute(0, 2, 3, 1)


class ImageEncoderViT(nn.Module):
    def __init__(self, img_size=1024, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4.0,
                 out_chans=256, window_size=14, global_attn_indexes=()):
        super().__init__()
        self.img_size = img_size
        self.patch_embed = PatchEmbed(patch_size, 3, embed_dim)
        g = img_size // patch_size
        self.pos_embed = nn.Parameter(torch.zeros(1, g, g, embed_dim))
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, mlp_ratio, window_size if i not in global_attn_indexes else 0, (g, g))
            for i in range(depth)
        ])
        self.neck = nn.Sequential(
            nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False),
            LayerNorm2d(out_chans),
            nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
            LayerNorm2d(out_chans),
        )

    def forward(self, x):
        x = self.patch_embed(x)
        x = x + self.pos_embed
        for blk in self.blocks:
            x = blk(x)
        return self.neck(x.permute(0, 3, 1, 2))


# ------------------------------------ prompt encoder -----------------------------------------
class PositionEmbeddingRandom(nn.Module):
    def __init__(self, num_pos_feats=64, scale=None):
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))

    def _pe_encoding(self, coords):
        coords = 2 * coords - 1
        coords = coords @ self.positional_encoding_gaussian_matrix
        coords = 2 * np.pi * coords
        return torch.cat([torch.sin(coords), torch.cos(coords)], dim=-1)

    def forward(self, size):
        h, w = size
        grid = torch.ones((h, w), dtype=torch.float32, device=self.positional_encoding_gaussian_matrix.device)
        y_embed = (grid.cumsum(dim=0) - 0.5) / h
        x_embed = (grid.cumsum(dim=1) - 0.5) / w
        pe = self._pe_encoding(torch.stack([x_embed, y_embed], dim=-1))
        return pe.permute(2, 0, 1)

    def forward_with_coords(self, coords_input, image_size):
        coords = coords_input.clone()
        coords[:, :, 0] = coords[:, :, 0] / image_size[1]
        coords[:, :, 1] = coords[:, :, 1] / image_size[0]
        return self._pe_encoding(coords.to(torch.float))


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim, image_embedding_size, input_image_size, mask_in_chans):
        super().__init__()
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings = 4
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(4)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans // 4),
            nn.GELU(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2),
            LayerNorm2d(mask_in_chans),
            nn.GELU(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1),
        )
        self.no_mask_embed = nn.Embedding(1, embed_dim)

    def get_dense_pe(self):
        return self.pe_layer(self.image_embedding_size).unsqueeze(0)

    def _embed_points(self, points, labels, pad):
        points = points + 0.5
        if pad:
            padding_point = torch.zeros((points.shape[0], 1, 2), device=points.device)
            padding_label = -torch.ones((labels.shape[0], 1), device=labels.device)
            points = torch.cat([points, padding_point], dim=1)
            labels = torch.cat([labels, padding_label], dim=1)
        point_embedding = self.pe_layer.forward_with_coords(points, self.input_image_size)
        point_embedding[labels == -1] = 0.0
        point_embedding[labels == -1] += self.not_a_point_embed.weight
        point_embedding[labels == 0] += self.point_embeddings[0].weight
        point_embedding[labels == 1] += self.point_embeddings[1].weight
        return point_embedding

    def _embed_boxes(self, boxes):
        boxes = boxes + 0.5
        coords = boxes.reshape(-1, 2, 2)
        corner_embedding = self.pe_layer.forward_with_coords(coords, self.input_image_size)
        corner_embedding[:, 0, :] += self.point_embeddings[2].weight
        corner_embedding[:, 1, :] += self.point_embeddings[3].weight
        return corner_embedding

    def forward(self, points, boxes, masks):
        if points is not None:
            bs = points[0].shape[0]
        elif boxes is not None:
            bs = boxes.shape[0]
        elif masks is not None:
            bs = masks.shape[0]
        else:
            bs = 1
        dev = self.point_embeddings[0].weight.device
        sparse = torch.empty((bs, 0, self.embed_dim), device=dev)
        if points is not None:
            coords, labels = points
            sparse = torch.cat([sparse, self._embed_points(coords, labels, pad=(boxes is None))], dim=1)
        if boxes is not None:
            sparse = torch.cat([sparse, self._embed_boxes(boxes)], dim=1)
        if masks is not None:
            dense = self.mask_downscaling(masks)
        else:
            dense = self.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(
                bs, -1, self.image_embedding_size[0], self.image_embedding_size[1]
            )
        return sparse, dense


# ------------------------------------ mask decoder -------------------------------------------
class DecAttention(nn.Module):
    def __init__(self, embedding_dim, num_heads, downsample_rate=1):
        super().__init__()
        self.internal_dim = embedding_dim // downsample_rate
        self.num_heads = num_heads
        self.q_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.k_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.v_proj = nn.Linear(embedding_dim, self.internal_dim)
        self.out_proj = nn.Linear(self.internal_dim, embedding_dim)

    def _separate_heads(self, x):
        b, n, c = x.shape
        return x.reshape(b, n, self.num_heads, c // self.num_heads).transpose(1, 2)

    def forward(self, q, k, v):
        q, k, v = self.q_proj(q), self.k_proj(k), self.v_proj(v)
        q, k, v = self._separate_heads(q), self._separate_heads(k), self._separate_heads(v)
        c_per_head = q.shape[-1]
        attn = q @ k.permute(0, 1, 3, 2)
        attn = attn / math.sqrt(c_per_head)
        attn = torch.softmax(attn, dim=-1)
        out = attn @ v
        b, h, n, c = out.shape
        out = out.transpose(1, 2).reshape(b, n, h * c)
        return self.out_proj(out)


class TwoWayAttentionBlock(nn.Module):
    def __init__(self, embedding_dim, num_heads, mlp_dim, attention_downsample_rate=2, skip_first_layer_pe=False):
        super().__init__()
        self.self_attn = DecAttention(embedding_dim, num_heads)
        self.norm1 = nn.LayerNorm(embedding_dim)  # eps 1e-5
        self.cross_attn_token_to_image = DecAttention(embedding_dim, num_heads, attention_downsample_rate)
        self.norm2 = nn.LayerNorm(embedding_dim)
        self.mlp = MLPBlock(embedding_dim, mlp_dim, nn.ReLU)
        self.norm3 = nn.LayerNorm(embedding_dim)
        self.norm4 = nn.LayerNorm(embedding_dim)
        self.cross_attn_image_to_token = DecAttention(embedding_dim, num_heads, attention_downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe

    def forward(self, queries, keys, query_pe, key_pe):
        if self.skip_first_layer_pe:
            queries = self.self_attn(q=queries, k=queries, v=queries)
        else:
            q = queries + query_pe
            queries = queries + self.self_attn(q=q, k=q, v=queries)
        queries = self.norm1(queries)
        q = queries + query_pe
        k = keys + key_pe
        queries = queries + self.cross_attn_token_to_image(q=q, k=k, v=keys)
        queries = self.norm2(queries)
        queries = self.norm3(queries + self.mlp(queries))
        q = queries + query_pe
        k = keys + key_pe
        keys = keys + self.cross_attn_image_to_token(q=k, k=q, v=queries)
        keys = self.norm4(keys)
        return queries, keys


class TwoWayTransformer(nn.Module):
    def __init__(self, depth, embedding_dim, num_heads, mlp_dim, attention_downsample_rate=2):
        super().__init__()
        self.layers = nn.ModuleList([
            TwoWayAttentionBlock(embedding_dim, num_heads, mlp_dim, attention_downsample_rate, (i == 0))
            for i in range(depth)
        ])
        self.final_attn_token_to_image = DecAttention(embedding_dim, num_heads, attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)

    def forward(self, image_embedding, image_pe, point_embedding):
        image_embedding = image_embedding.flatten(2).permute(0, 2, 1)
        image_pe = image_pe.flatten(2).permute(0, 2, 1)
        queries, keys = point_embedding, image_embedding
        for layer in self.layers:
            queries, keys = layer(queries, keys, point_embedding, image_pe)
        q = queries + point_embedding
        k = keys + image_pe
        queries = queries + self.final_attn_token_to_image(q=q, k=k, v=keys)
        queries = self.norm_final_attn(queries)
        return queries, keys


class MLP(nn.Module):
    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = F.relu(layer(x)) if i < self.num_layers - 1 else layer(x)
        return x


class MaskDecoder(nn.Module):
    def __init__(self, transformer_dim, transformer, num_multimask_outputs=3, iou_head_depth=3,
                 iou_head_hidden_dim=256):
        super().__init__()
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4),
            nn.GELU(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2),
            nn.GELU(),
        )
        self.output_hypernetworks_mlps = nn.ModuleList(
            [MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens)]
        )
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)

    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings,
                multimask_output):
        masks, iou_pred = self.predict_masks(image_embeddings, image_pe, sparse_prompt_embeddings,
                                             dense_prompt_embeddings)
        mask_slice = slice(1, None) if multimask_output else slice(0, 1)
        return masks[:, mask_slice, :, :], iou_pred[:, mask_slice]

    def predict_masks(self, image_embeddings, image_pe, sparse, dense):
        output_tokens = torch.cat([self.iou_token.weight, self.mask_tokens.weight], dim=0)
        output_tokens = output_tokens.unsqueeze(0).expand(sparse.size(0), -1, -1)
        tokens = torch.cat((output_tokens, sparse), dim=1)
        src = torch.repeat_interleave(image_embeddings, tokens.shape[0], dim=0)
        src = src + dense
        pos_src = torch.repeat_interleave(image_pe, tokens.shape[0], dim=0)
        b, c, h, w = src.shape
        hs, src = self.transformer(src, pos_src, tokens)
        iou_token_out = hs[:, 0, :]
        mask_tokens_out = hs[:, 1:(1 + self.num_mask_tokens), :]
        src = src.transpose(1, 2).view(b, c, h, w)
        upscaled = self.output_upscaling(src)
        hyper_in = torch.stack(
            [self.output_hypernetworks_mlps[i](mask_tokens_out[:, i, :]) for i in range(self.num_mask_tokens)], dim=1
        )
        b, c, h, w = upscaled.shape
        masks = (hyper_in @ upscaled.view(b, c, h * w)).view(b, -1, h, w)
        iou_pred = self.iou_prediction_head(iou_token_out)
        return masks, iou_pred


# ------------------------------------ Sam / predictor ----------------------------------------
class Sam(nn.Module):
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, image_encoder, prompt_encoder, mask_decoder,
                 pixel_mean=(123.675, 116.28, 103.53), pixel_std=(58.395, 57.12, 57.375)):
        super().__init__()
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.register_buffer("pixel_mean", torch.Tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.Tensor(pixel_std).view(-1, 1, 1), False)

    @property
    def device(self):
        return self.pixel_mean.device

    def preprocess(self, x):
        x = (x - self.pixel_mean) / self.pixel_std
        h, w = x.shape[-2:]
        return F.pad(x, (0, self.image_encoder.img_size - w, 0, self.image_encoder.img_size - h))

    def postprocess_masks(self, masks, input_size, original_size):
        s = self.image_encoder.img_size
        masks = F.interpolate(masks, (s, s), mode="bilinear", align_corners=False)
        masks = masks[..., : input_size[0], : input_size[1]]
        return F.interpolate(masks, original_size, mode="bilinear", align_corners=False)


def get_preprocess_shape(oldh, oldw, long_side):
    scale = long_side * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


class ResizeLongestSide:
    def __init__(self, target_length):
        self.target_length = target_length

    def apply_image(self, image: np.ndarray) -> np.ndarray:
        """uint8 HWC -> uint8 HWC.  Upstream: torchvision ``resize(to_pil_image(img), target)`` = PIL bilinear
        with antialias.  PIL is present in this image, so the oracle uses exactly that call."""
        from PIL import Image
        th, tw = get_preprocess_shape(image.shape[0], image.shape[1], self.target_length)
        if (th, tw) == image.shape[:2]:
            return np.array(image)
        return np.array(Image.fromarray(image).resize((tw, th), Image.BILINEAR))

    def apply_coords(self, coords, original_size):
        old_h, old_w = original_size
        new_h, new_w = get_preprocess_shape(old_h, old_w, self.target_length)
        coords = np.array(coords, dtype=float, copy=True)
        coords[..., 0] = coords[..., 0] * (new_w / old_w)
        coords[..., 1] = coords[..., 1] * (new_h / old_h)
        return coords

    def apply_boxes(self, boxes, original_size):
        return self.apply_coords(np.asarray(boxes).reshape(-1, 2, 2), original_size).reshape(-1, 4)


class SamPredictor:
    def __init__(self, sam_model: Sam):
        self.model = sam_model
        self.transform = ResizeLongestSide(sam_model.image_encoder.img_size)
        self.reset_image()

    @property
    def device(self):
        return self.model.device

    def reset_image(self):
        self.is_image_set = False
        self.features = None
        self.orig_h = self.orig_w = self.input_h = self.input_w = None

    @torch.no_grad()
    def set_image(self, image: np.ndarray, image_format: str = "RGB"):
        if image_format != self.model.image_format:
            image = image[..., ::-1]
        x = self.transform.apply_image(image)
        x = torch.as_tensor(x, device=self.device).permute(2, 0, 1).contiguous()[None]
        self.set_torch_image(x, image.shape[:2])

    @torch.no_grad()
    def set_torch_image(self, transformed_image, original_image_size):
        self.reset_image()
        self.original_size = tuple(original_image_size)
        self.input_size = tuple(transformed_image.shape[-2:])
        self.features = self.model.image_encoder(self.model.preprocess(transformed_image))
        self.is_image_set = True

    def get_image_embedding(self):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        return self.features

    def predict(self, point_coords=None, point_labels=None, box=None, mask_input=None, multimask_output=True,
                return_logits=False):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        coords_t = labels_t = box_t = mask_t = None
        if point_coords is not None:
            pc = self.transform.apply_coords(point_coords, self.original_size)
            coords_t = torch.as_tensor(pc, dtype=torch.float, device=self.device)[None]
            labels_t = torch.as_tensor(point_labels, dtype=torch.int, device=self.device)[None]
        if box is not None:
            box_t = torch.as_tensor(self.transform.apply_boxes(box, self.original_size), dtype=torch.float,
                                    device=self.device)[None]
            box_t = box_t.reshape(1, 4)
        if mask_input is not None:
            mask_t = torch.as_tensor(mask_input, dtype=torch.float, device=self.device)[None]
        m, s, l = self.predict_torch(coords_t, labels_t, box_t, mask_t, multimask_output, return_logits)
        return m[0].cpu().numpy(), s[0].cpu().numpy(), l[0].cpu().numpy()

    @torch.no_grad()
    def predict_torch(self, point_coords, point_labels, boxes=None, mask_input=None, multimask_output=True,
                      return_logits=False):
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        points = (point_coords, point_labels) if point_coords is not None else None
        sparse, dense = self.model.prompt_encoder(points=points, boxes=boxes, masks=mask_input)
        low_res, iou = self.model.mask_decoder(
            image_embeddings=self.features, image_pe=self.model.prompt_encoder.get_dense_pe(),
            sparse_prompt_embeddings=sparse, dense_prompt_embeddings=dense, multimask_output=multimask_output,
        )
        masks = self.model.postprocess_masks(low_res, self.input_size, self.original_size)
        if not return_logits:
            masks = masks > self.model.mask_threshold
        return masks, iou, low_res


# ------------------------------------ builders -----------------------------------------------
def build_sam(model_type: str = "vit_b", image_size: int = 1024, num_multimask_outputs: int = 3) -> Sam:
    prompt_embed_dim, patch = 256, 16
    g = image_size // patch
    if model_type == "vit_t":   # MobileSAM: TinyViT encoder, segment_anything's prompt encoder / mask decoder (util.py:436-441)
        from .tinyvit_ref import TinyViT
        encoder = TinyViT(img_size=image_size)
    else:
        a = ARCH[model_type]
        encoder = ImageEncoderViT(img_size=image_size, patch_size=patch, embed_dim=a["embed_dim"],
                                  depth=a["depth"], num_heads=a["num_heads"], mlp_ratio=4,
                                  out_chans=prompt_embed_dim, window_size=14,
                                  global_attn_indexes=a["global_attn_indexes"])
    sam = Sam(
        image_encoder=encoder,
        prompt_encoder=PromptEncoder(prompt_embed_dim, (g, g), (image_size, image_size), 16),
        mask_decoder=MaskDecoder(prompt_embed_dim,
                                 TwoWayTransformer(depth=2, embedding_dim=prompt_embed_dim, mlp_dim=2048,
                                                   num_heads=8),
                                 num_multimask_outputs=num_multimask_outputs, iou_head_depth=3,
                                 iou_head_hidden_dim=256),
    )
    sam.eval()
    return sam


def seeded_state_dict(model_type: str = "vit_b", seed: int = 0, image_size: int = 1024):
    """Seeded random weights with upstream key names: Linear/Conv ~ N(0, s), LN gamma~1, beta~0,
    non-zero pos_embed / rel_pos so every path is exercised.  ``s`` is chosen per-tensor as
    1/sqrt(fan_in) so activations stay O(1) through 12-32 blocks (N(0,0.02) collapses the decoder
    logits to ~0 which would make mask parity vacuous)."""
    sam = build_sam(model_type, image_size)
    g = torch.Generator().manual_seed(seed)
    ln_mods = {n for n, m in sam.named_modules()
               if isinstance(m, (nn.LayerNorm, nn.BatchNorm2d)) or type(m).__name__ == "LayerNorm2d"}
    sd = {}
    for k, v in sam.state_dict().items():
        mod, leaf = k.rsplit(".", 1) if "." in k else ("", k)
        if k.endswith("positional_encoding_gaussian_matrix"):
            sd[k] = torch.randn(v.shape, generator=g)
        elif leaf == "num_batches_tracked":
            sd[k] = torch.zeros((), dtype=torch.long)
        elif leaf == "running_mean":      # BatchNorm statistics (vit_t): non-trivial so the folding is exercised
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif leaf == "running_var":
            sd[k] = 0.5 + torch.rand(v.shape, generator=g)
        elif leaf == "attention_biases":
            sd[k] = 0.5 * torch.randn(v.shape, generator=g)
        elif mod in ln_mods:
            sd[k] = (1.0 if leaf == "weight" else 0.0) + 0.1 * torch.randn(v.shape, generator=g)
        elif v.ndim == 1:  # biases
            sd[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif "pos_embed" in k or "rel_pos" in k:
            sd[k] = 0.3 * torch.randn(v.shape, generator=g)
        elif k.endswith("embed.weight") or "point_embeddings" in k or k.endswith("_token.weight") \
                or k.endswith("mask_tokens.weight"):
            sd[k] = 0.5 * torch.randn(v.shape, generator=g)
        else:
            if "output_upscaling" in k and v.ndim == 4:   # ConvTranspose2d weight (in, out, k, k)
                fan_in = v.shape[0]
            else:
                fan_in = int(np.prod(v.shape[1:]))
            sd[k] = torch.randn(v.shape, generator=g) / math.sqrt(fan_in)
    return sd


def build_seeded_sam(model_type="vit_b", seed=0, image_size=1024) -> Sam:
    sam = build_sam(model_type, image_size)
    sam.load_state_dict(seeded_state_dict(model_type, seed, image_size))
    return sam
