"""CPU checks of the TinyViT (vit_t / MobileSAM) oracle restatement (oracle/tinyvit_ref.py): published checkpoint layout (key
names and shapes), output geometry, the padding rule of the window attention and eval-mode BatchNorm folding (the algebra the
CUDA path applies at load time, csrc/tinyvit.cu:fold_conv_bn).  Parity of the values themselves is unpinned (no second
TinyViT implementation in this image) -- stated in the oracle's header."""
import numpy as np
import torch

from oracle import sam_ref, tinyvit_ref


def test_state_dict_layout_of_the_published_checkpoint():
    sd = sam_ref.seeded_state_dict("vit_t", seed=0)
    expect = {
        "image_encoder.patch_embed.seq.0.c.weight": (32, 3, 3, 3),
        "image_encoder.patch_embed.seq.2.bn.running_var": (64,),
        "image_encoder.layers.0.blocks.1.conv2.c.weight": (256, 1, 3, 3),
        "image_encoder.layers.0.downsample.conv1.c.weight": (128, 64, 1, 1),
        "image_encoder.layers.1.blocks.0.attn.attention_biases": (4, 49),
        "image_encoder.layers.1.blocks.0.attn.qkv.weight": (384, 128),
        "image_encoder.layers.1.downsample.conv2.c.weight": (160, 1, 3, 3),
        "image_encoder.layers.2.blocks.5.attn.attention_biases": (5, 196),
        "image_encoder.layers.2.blocks.5.mlp.fc1.weight": (640, 160),
        "image_encoder.layers.2.downsample.conv3.c.weight": (320, 320, 1, 1),
        "image_encoder.layers.3.blocks.1.attn.attention_biases": (10, 49),
        "image_encoder.layers.3.blocks.1.local_conv.c.weight": (320, 1, 3, 3),
        "image_encoder.norm_head.weight": (320,),
        "image_encoder.head.weight": (1000, 320),
        "image_encoder.neck.0.weight": (256, 320, 1, 1),
        "image_encoder.neck.2.weight": (256, 256, 3, 3),
        "mask_decoder.iou_token.weight": (1, 256),
    }
    for k, shape in expect.items():
        assert k in sd and tuple(sd[k].shape) == shape, (k, tuple(sd[k].shape) if k in sd else None)
    assert not any("attention_bias_idxs" in k for k in sd)     # non-persistent buffer upstream
    assert "image_encoder.layers.3.downsample.conv1.c.weight" not in sd
    from micro_sam_b200.sam import validate_model_type
    assert validate_model_type(sd) == "vit_t"
    n_enc = sum(v.numel() for k, v in sd.items() if k.startswith("image_encoder.") and v.dtype.is_floating_point)
    assert 5.5e6 < n_enc < 6.5e6    # TinyViT-5M + classifier head


def test_forward_geometry_and_determinism():
    sam = sam_ref.build_seeded_sam("vit_t", seed=0)
    x = torch.randn(1, 3, 1024, 1024, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        y = sam.image_encoder(x)
        y2 = sam.image_encoder(x)
    assert y.shape == (1, 256, 64, 64) and torch.equal(y, y2) and 0.3 < float(y.std()) < 3.0


def test_window_padding_tokens_pass_through_the_norm():
    """Pad tokens are zeros BEFORE attn.norm (they become LayerNorm(0) = bias and act as keys)."""
    blk = tinyvit_ref.TinyViTBlock(32, (9, 9), num_heads=1, window_size=7, mlp_ratio=4.0, local_conv_size=3).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.2)
        x = torch.randn(1, 81, 32, generator=g)
        ref = blk(x)
        # restate the attention part by hand: pad to 14 x 14, 4 windows, keys include the pad tokens
        xp = torch.zeros(1, 14, 14, 32)
        xp[:, :9, :9] = x.view(1, 9, 9, 32)
        w = xp.view(1, 2, 7, 2, 7, 32).transpose(2, 3).reshape(4, 49, 32)
        a = blk.attn(w).view(1, 2, 2, 7, 7, 32).transpose(2, 3).reshape(1, 14, 14, 32)[:, :9, :9].reshape(1, 81, 32)
        y = x + a
        y = blk.local_conv(y.transpose(1, 2).reshape(1, 32, 9, 9)).view(1, 32, 81).transpose(1, 2)
        y = y + blk.mlp(y)
    assert torch.allclose(ref, y, atol=1e-6)
    # a pad row after the norm equals the norm's bias
    with torch.no_grad():
        assert torch.allclose(blk.attn.norm(torch.zeros(1, 32)), blk.attn.norm.bias[None], atol=1e-7)


def test_batchnorm_folding_algebra():
    """conv -> BN(eval) == conv with weight * gamma / sqrt(var + eps) and bias beta - mean * gamma / sqrt(var + eps)."""
    g = torch.Generator().manual_seed(2)
    m = tinyvit_ref.Conv2d_BN(8, 16, 3, 2, 1).eval()
    with torch.no_grad():
        m.c.weight.copy_(torch.randn(m.c.weight.shape, generator=g))
        m.bn.weight.copy_(1 + 0.1 * torch.randn(16, generator=g))
        m.bn.bias.copy_(0.1 * torch.randn(16, generator=g))
        m.bn.running_mean.copy_(0.1 * torch.randn(16, generator=g))
        m.bn.running_var.copy_(0.5 + torch.rand(16, generator=g))
        x = torch.randn(2, 8, 10, 10, generator=g)
        s = m.bn.weight / torch.sqrt(m.bn.running_var + 1e-5)
        y = torch.nn.functional.conv2d(x, m.c.weight * s[:, None, None, None], m.bn.bias - m.bn.running_mean * s, 2, 1)
        assert torch.allclose(m(x), y, atol=1e-5)


def test_attention_bias_index_is_the_offset_table():
    for ws in (7, 14):
        idx, n = tinyvit_ref.attention_bias_idxs((ws, ws))
        p = torch.arange(ws * ws)
        y, x = p // ws, p % ws
        assert n == ws * ws and torch.equal(idx, (y[:, None] - y[None]).abs() * ws + (x[:, None] - x[None]).abs())
