"""`-m gpu` parity of the encoder backward pass (BASELINE.json configs[4], csrc/encoder_train.cu) against torch autograd over the
fp32 oracle encoder (oracle/sam_ref.py): op level (batched attention-backward GEMM in its four operand layouts, LayerNorm
backward) and model level (every parameter gradient of the image encoder for a random upstream gradient dL/d embedding).
Tolerance: rel-L2 <= 3e-2 per gradient tensor (bf16 operands, fp32 accumulation; the forward tolerance is 2e-2)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("hd,Tq", [(64, 196), (80, 196), (64, 512)])
def test_bgemm_layouts_against_torch(hd, Tq):
    """C[w,h] = op(A[w,h]) op(B[w,h]) on head / window slices of packed buffers, incl. the zero-filled tails (head_dim 80: the
    second 64-column box; 196-row windows in 64-row boxes)."""
    from micro_sam_b200 import _lib
    L = _lib.lib()
    H, W = 3, 2
    D = H * hd
    g = torch.Generator().manual_seed(0)
    qkv = torch.randn(W * Tq, 3 * D, generator=g).to(DEV).bfloat16()
    q = qkv[:, :D].float().view(W, Tq, H, hd).permute(0, 2, 1, 3)            # [W,H,Tq,hd]
    k = qkv[:, D:2 * D].float().view(W, Tq, H, hd).permute(0, 2, 1, 3)
    pitch = (Tq + 7) // 8 * 8
    # NT: S = Q K^T
    S = torch.full((W, H, Tq, pitch), 7.0, device=DEV)
    _lib.check(L.msam_op_bgemm(_lib.ptr(qkv), ctypes_off(qkv, D), 0, 0, Tq, Tq, hd, 3 * D, 3 * D, hd, Tq * 3 * D, hd, Tq * 3 * D, H, W,
                               _lib.ptr(S), pitch, Tq * pitch, H * Tq * pitch, 0.5, 0, _lib.cur_stream()))
    torch.cuda.synchronize()
    assert _rel(S[..., :Tq], 0.5 * q @ k.transpose(-1, -2)) < 1e-5
    assert bool((S[..., Tq:] == 7.0).all())                                     # columns beyond N untouched
    # TN: dV = P^T dO  (P [Tq, pitch] per batch entry, dO as a head slice of a [rows, D] buffer), then accumulate
    P = torch.randn(W, H, Tq, pitch, generator=g).to(DEV).bfloat16()
    dO = torch.randn(W * Tq, D, generator=g).to(DEV).bfloat16()
    do = dO.float().view(W, Tq, H, hd).permute(0, 2, 1, 3)
    dV = torch.zeros(W, H, Tq, hd, device=DEV)
    for acc in (0, 1):
        _lib.check(L.msam_op_bgemm(_lib.ptr(P), _lib.ptr(dO), 1, 1, Tq, hd, Tq, pitch, D, Tq * pitch, H * Tq * pitch, hd, Tq * D, H, W,
                                   _lib.ptr(dV), hd, Tq * hd, H * Tq * hd, 1.0, acc, _lib.cur_stream()))
    torch.cuda.synchronize()
    assert _rel(dV, 2 * P.float()[..., :Tq].transpose(-1, -2) @ do) < 1e-5
    # NN: dQ = dS K  (A K-major [Tq, Tk], B = the K slice consumed MN-major)
    dQ = torch.empty(W, H, Tq, hd, device=DEV)
    _lib.check(L.msam_op_bgemm(_lib.ptr(P), ctypes_off(qkv, D), 0, 1, Tq, hd, Tq, pitch, 3 * D, Tq * pitch, H * Tq * pitch, hd,
                               Tq * 3 * D, H, W, _lib.ptr(dQ), hd, Tq * hd, H * Tq * hd, 1.0, 0, _lib.cur_stream()))
    torch.cuda.synchronize()
    assert _rel(dQ, P.float()[..., :Tq] @ k) < 1e-5
    # NT against a table shared by every batch entry: T = Q R^T
    R = torch.zeros(64, (hd + 63) // 64 * 64, device=DEV, dtype=torch.bfloat16)
    R[:27, :hd] = torch.randn(27, hd, generator=g).to(DEV).bfloat16()
    T = torch.empty(W, H, Tq, 64, device=DEV)
    _lib.check(L.msam_op_bgemm(_lib.ptr(qkv), _lib.ptr(R), 0, 0, Tq, 64, hd, 3 * D, R.shape[1], hd, Tq * 3 * D, 0, 0, H, W,
                               _lib.ptr(T), 64, Tq * 64, H * Tq * 64, 1.0, 0, _lib.cur_stream()))
    torch.cuda.synchronize()
    assert _rel(T, q @ R.float()[:, :hd].t()) < 1e-5


def ctypes_off(t, elems):
    import ctypes
    return ctypes.c_void_p(t.data_ptr() + elems * t.element_size())


@pytest.mark.parametrize("D,window", [(768, 0), (1280, 0), (128, 1)])
def test_layernorm_backward_against_autograd(D, window):
    from micro_sam_b200 import _lib
    L = _lib.lib()
    B, g = 2, 64
    rows = B * g * g
    gen = torch.Generator().manual_seed(1)
    x = (torch.randn(rows, D, generator=gen) * 2 + 0.5).to(DEV).requires_grad_(True)
    gamma = (1 + 0.1 * torch.randn(D, generator=gen)).to(DEV).requires_grad_(True)
    beta = (0.1 * torch.randn(D, generator=gen)).to(DEV).requires_grad_(True)
    dy = torch.randn(rows, D, generator=gen).to(DEV)
    y = torch.nn.functional.layer_norm(x, (D,), gamma, beta, 1e-6)
    y.backward(dy)
    dy_in = dy
    if window:   # dy handed over in the window-partitioned row order (pad rows hold garbage that must be ignored)
        pad = torch.nn.functional.pad(dy.view(B, g, g, D), (0, 0, 0, 6, 0, 6), value=123.0)
        dy_in = pad.view(B, 5, 14, 5, 14, D).permute(0, 1, 3, 2, 4, 5).reshape(-1, D).contiguous()
    dx = torch.ones(rows, D, device=DEV)
    dg, db = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
    _lib.check(L.msam_op_layernorm_bwd(_lib.ptr(x.detach()), rows, D, _lib.ptr(gamma.detach()), 1e-6, _lib.ptr(dy_in), window, 1,
                                       _lib.ptr(dx), _lib.ptr(dg), _lib.ptr(db), _lib.cur_stream()))
    torch.cuda.synchronize()
    assert _rel(dx - 1.0, x.grad) < 1e-4 and _rel(dg, gamma.grad) < 1e-4 and _rel(db, beta.grad) < 1e-4


def _encoder_grad_case(model_type, B, names=None, tol=3e-2):
    from oracle import sam_ref
    from micro_sam_b200.sam import B200Sam
    from micro_sam_b200.sample_data import lm_tile
    from micro_sam_b200 import util
    torch.set_num_threads(min(64, os.cpu_count() or 8))
    sd = sam_ref.seeded_state_dict(model_type, seed=0)
    osam = sam_ref.build_sam(model_type)
    osam.load_state_dict(sd)
    enc_sd = {k: v for k, v in sd.items() if k.startswith("image_encoder.")}
    sam = B200Sam(model_type, enc_sd, max_batch=B, max_prompts=1).train()
    imgs = np.stack([util._to_image(lm_tile((1024, 1024), 150, seed=s)) for s in range(B)])
    x_pre = torch.stack([osam.preprocess(torch.from_numpy(i).permute(2, 0, 1)[None].float())[0] for i in imgs])
    gen = torch.Generator().manual_seed(5)
    d_out = torch.randn(B, 256, 64, 64, generator=gen)
    enc = osam.image_encoder
    for p in enc.parameters():
        p.requires_grad_(True)
    ref_out = enc(x_pre)
    ref_out.backward(d_out)
    got_out = sam.image_encoder(x_pre.to(DEV))
    assert got_out.requires_grad
    rel_f = _rel(got_out.detach(), ref_out.detach())
    got_out.backward(d_out.to(DEV))
    grads = sam.encoder_grads(names)
    rels = {}
    for k, gten in grads.items():
        ref = dict(enc.named_parameters())[k[len("image_encoder."):]].grad
        assert ref is not None and tuple(ref.shape) == tuple(gten.shape), k
        rels[k] = _rel(gten, ref)
    worst = sorted(rels.items(), key=lambda kv: -kv[1])[:6]
    print(f"{model_type} B={B}: forward rel-L2 {rel_f:.2e}; {len(rels)} gradients, median rel-L2 {np.median(list(rels.values())):.2e}, worst: "
          + ", ".join(f"{k.replace('image_encoder.', '')} {v:.2e}" for k, v in worst))
    assert rel_f < 2e-2
    bad = {k: v for k, v in rels.items() if not v < tol}
    assert not bad, bad
    # the inference path of the same engine is untouched by training mode
    with torch.no_grad():
        assert _rel(sam.eval().image_encoder(x_pre.to(DEV)), ref_out.detach()) < 2e-2
    del sam
    torch.cuda.empty_cache()


@pytest.mark.parametrize("model_type", ["vit_test", "vit_test80"])
def test_encoder_backward_small_archs(model_type):
    """One windowed + one global block, head_dim 64 / 80, batch 2: every parameter gradient."""
    _encoder_grad_case(model_type, B=2)


def test_encoder_backward_vit_b():
    """The architecture cfg 5 names, one image (the fp32 autograd reference needs ~30 GB of host memory for the global blocks)."""
    _encoder_grad_case("vit_b", B=1)
