"""ORACLE (test infrastructure, NOT product code) -- fp32 CPU restatement of MobileSAM's TinyViT image encoder (``vit_t``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this.

micro-sam reaches this encoder only through ``mobile_sam.sam_model_registry["vit_t"]()`` (``micro_sam/util.py:35-43,436-441``);
``mobile_sam`` (git+https://github.com/ChaoningZhang/MobileSAM.git @ HEAD, ``environment.yaml:34``) is not under
``/root/reference`` and not installed in this image, and the reference holds no architectural statement or numeric fixture
for it.  This file restates the published ``mobile_sam/modeling/tiny_vit_sam.py`` (TinyViT-5M configured as in
``mobile_sam/build_sam.py:build_sam_vit_t``) with upstream state-dict key names:

    TinyViT(img_size=1024, embed_dims=[64,128,160,320], depths=[2,2,6,2], num_heads=[2,4,5,10],
            window_sizes=[7,7,14,7], mlp_ratio=4, mbconv_expand_ratio=4, local_conv_size=3)

PARITY PINNING: **parity unpinned** -- there is no second implementation of TinyViT in this container (no timm, no HF
port) and no golden vector in the reference; the restatement is checked only for internal consistency (shapes, BN folding,
key names of the published checkpoint layout).  The prompt encoder / mask decoder of ``vit_t`` are the ones in
``oracle/sam_ref.py`` (MobileSAM reuses segment_anything's)."""
from __future__ import annotations

import itertools

import torch
import torch.nn as nn
import torch.nn.functional as F

EMBED_DIMS = (64, 128, 160, 320)
DEPTHS = (2, 2, 6, 2)
NUM_HEADS = (2, 4, 5, 10)
WINDOW_SIZES = (7, 7, 14, 7)


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


class Conv2d_BN(nn.Sequential):
    """conv (no bias) + BatchNorm2d; keys ``c.weight`` / ``bn.{weight,bias,running_mean,running_var,num_batches_tracked}``."""

    def __init__(self, a, b, ks=1, stride=1, pad=0, groups=1):
        super().__init__()
        self.add_module("c", nn.Conv2d(a, b, ks, stride, pad, 1, groups, bias=False))
        self.add_module("bn", nn.BatchNorm2d(b))


class PatchEmbed(nn.Module):
    def __init__(self, in_chans, embed_dim):
        super().__init__()
        n = embed_dim
        self.seq = nn.Sequential(Conv2d_BN(in_chans, n // 2, 3, 2, 1), nn.GELU(), Conv2d_BN(n // 2, n, 3, 2, 1))

    def forward(self, x):
        return self.seq(x)


class MBConv(nn.Module):
    def __init__(self, in_chans, out_chans, expand_ratio):
        super().__init__()
        hidden = int(in_chans * expand_ratio)
        self.conv1 = Conv2d_BN(in_chans, hidden, ks=1)
        self.act1 = nn.GELU()
        self.conv2 = Conv2d_BN(hidden, hidden, ks=3, stride=1, pad=1, groups=hidden)
        self.act2 = nn.GELU()
        self.conv3 = Conv2d_BN(hidden, out_chans, ks=1)
        self.act3 = nn.GELU()

    def forward(self, x):
        shortcut = x
        x = self.act1(self.conv1(x))
        x = self.act2(self.conv2(x))
        x = self.conv3(x)
        x = x + shortcut
        return self.act3(x)


class PatchMerging(nn.Module):
    def __init__(self, input_resolution, dim, out_dim):
        super().__init__()
        self.input_resolution = input_resolution
        self.act = nn.GELU()
        self.conv1 = Conv2d_BN(dim, out_dim, 1, 1, 0)
        stride_c = 1 if out_dim in (320, 448, 576) else 2
        self.conv2 = Conv2d_BN(out_dim, out_dim, 3, stride_c, 1, groups=out_dim)
        self.conv3 = Conv2d_BN(out_dim, out_dim, 1, 1, 0)

    def forward(self, x):
        if x.ndim == 3:
            H, W = self.input_resolution
            B = len(x)
            x = x.view(B, H, W, -1).permute(0, 3, 1, 2)
        x = self.act(self.conv1(x))
        x = self.act(self.conv2(x))
        x = self.conv3(x)
        return x.flatten(2).transpose(1, 2)


class ConvLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, out_dim, conv_expand_ratio):
        super().__init__()
        self.blocks = nn.ModuleList([MBConv(dim, dim, conv_expand_ratio) for _ in range(depth)])
        self.downsample = PatchMerging(input_resolution, dim=dim, out_dim=out_dim)

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return self.downsample(x)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.norm = nn.LayerNorm(in_features)
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)
        self.act = nn.GELU()

    def forward(self, x):
        return self.fc2(self.act(self.fc1(self.norm(x))))


def attention_bias_idxs(resolution):
    """Index table of TinyViT's Attention: offsets (|dy|, |dx|) numbered in order of first appearance."""
    points = list(itertools.product(range(resolution[0]), range(resolution[1])))
    offsets, idxs = {}, []
    for p1 in points:
        for p2 in points:
            off = (abs(p1[0] - p2[0]), abs(p1[1] - p2[1]))
            if off not in offsets:
                offsets[off] = len(offsets)
            idxs.append(offsets[off])
    n = len(points)
    return torch.LongTensor(idxs).view(n, n), len(offsets)


class Attention(nn.Module):
    def __init__(self, dim, key_dim, num_heads, attn_ratio, resolution):
        super().__init__()
        self.num_heads = num_heads
        self.scale = key_dim ** -0.5
        self.key_dim = key_dim
        self.d = int(attn_ratio * key_dim)
        self.dh = self.d * num_heads
        h = self.dh + key_dim * num_heads * 2
        self.norm = nn.LayerNorm(dim)
        self.qkv = nn.Linear(dim, h)
        self.proj = nn.Linear(self.dh, dim)
        idxs, n_off = attention_bias_idxs(resolution)
        self.attention_biases = nn.Parameter(torch.zeros(num_heads, n_off))
        self.register_buffer("attention_bias_idxs", idxs, persistent=False)

    def forward(self, x):
        B, N, _ = x.shape
        x = self.norm(x)
        qkv = self.qkv(x)
        # per head the channels are [q | k | v]
        q, k, v = qkv.view(B, N, self.num_heads, -1).split([self.key_dim, self.key_dim, self.d], dim=3)
        q, k, v = q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)
        attn = (q @ k.transpose(-2, -1)) * self.scale + self.attention_biases[:, self.attention_bias_idxs]
        attn = attn.softmax(dim=-1)
        x = (attn @ v).transpose(1, 2).reshape(B, N, self.dh)
        return self.proj(x)


class TinyViTBlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, window_size, mlp_ratio, local_conv_size):
        super().__init__()
        self.input_resolution = input_resolution
        self.window_size = window_size
        head_dim = dim // num_heads
        self.attn = Attention(dim, head_dim, num_heads, attn_ratio=1, resolution=(window_size, window_size))
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.local_conv = Conv2d_BN(dim, dim, ks=local_conv_size, stride=1, pad=local_conv_size // 2, groups=dim)

    def forward(self, x):
        H, W = self.input_resolution
        B, L, C = x.shape
        ws = self.window_size
        res_x = x
        if H == ws and W == ws:
            x = self.attn(x)
        else:
            x = x.view(B, H, W, C)
            pad_b = (ws - H % ws) % ws
            pad_r = (ws - W % ws) % ws
            if pad_b > 0 or pad_r > 0:
                x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b))   # zero tokens BEFORE attn.norm: they become LN(0) = beta
            pH, pW = H + pad_b, W + pad_r
            nH, nW = pH // ws, pW // ws
            x = x.view(B, nH, ws, nW, ws, C).transpose(2, 3).reshape(B * nH * nW, ws * ws, C)
            x = self.attn(x)
            x = x.view(B, nH, nW, ws, ws, C).transpose(2, 3).reshape(B, pH, pW, C)
            x = x[:, :H, :W].contiguous().view(B, L, C)
        x = res_x + x
        x = x.transpose(1, 2).reshape(B, C, H, W)
        x = self.local_conv(x)
        x = x.view(B, C, L).transpose(1, 2)
        return x + self.mlp(x)


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, window_size, mlp_ratio, local_conv_size, out_dim, downsample):
        super().__init__()
        self.blocks = nn.ModuleList([TinyViTBlock(dim, input_resolution, num_heads, window_size, mlp_ratio, local_conv_size)
                                     for _ in range(depth)])
        self.downsample = PatchMerging(input_resolution, dim=dim, out_dim=out_dim) if downsample else None

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return x if self.downsample is None else self.downsample(x)


class TinyViT(nn.Module):
    def __init__(self, img_size=1024, in_chans=3, num_classes=1000, embed_dims=EMBED_DIMS, depths=DEPTHS,
                 num_heads=NUM_HEADS, window_sizes=WINDOW_SIZES, mlp_ratio=4.0, mbconv_expand_ratio=4.0, local_conv_size=3):
        super().__init__()
        self.img_size = img_size
        self.patch_embed = PatchEmbed(in_chans, embed_dims[0])
        pr = img_size // 4
        self.layers = nn.ModuleList()
        n = len(depths)
        for i in range(n):
            res = pr // (2 ** (i - 1 if i == 3 else i))
            out_dim = embed_dims[min(i + 1, n - 1)]
            if i == 0:
                layer = ConvLayer(embed_dims[0], (res, res), depths[0], out_dim, mbconv_expand_ratio)
            else:
                layer = BasicLayer(embed_dims[i], (res, res), depths[i], num_heads[i], window_sizes[i], mlp_ratio,
                                   local_conv_size, out_dim, downsample=i < n - 1)
            self.layers.append(layer)
        # classifier head: unused by SAM but present in the published checkpoints
        self.norm_head = nn.LayerNorm(embed_dims[-1])
        self.head = nn.Linear(embed_dims[-1], num_classes)
        self.neck = nn.Sequential(
            nn.Conv2d(embed_dims[-1], 256, kernel_size=1, bias=False), LayerNorm2d(256),
            nn.Conv2d(256, 256, kernel_size=3, padding=1, bias=False), LayerNorm2d(256))

    def forward(self, x):
        x = self.patch_embed(x)
        for layer in self.layers:
            x = layer(x)
        B, _, C = x.size()
        g = self.img_size // 16
        x = x.view(B, g, g, C).permute(0, 3, 1, 2)
        return self.neck(x)
