"""Noise floor of bf16 training for the mask decoder (tool, not a test; run on CPU: PYTHONPATH=. python tests/bf16_noise_floor.py).
The oracle decoder under torch.autocast(bfloat16) against itself in fp32, same seeded weights / prompts / functional as
tests/test_gpu_backward.py::test_decoder_train_against_autograd[boxes-True].  Measured here: forward rel-L2 1.5e-2, d embedding 1.8e-2,
parameter gradients median 8.2e-2 (worst 1.3e-1) -- the tolerances of the GPU test (1e-1, cosine > 0.995) sit at this floor; the CUDA
path itself measures 1.1e-2 / 1.3e-2 / 6e-2 against the same fp32 reference."""
import torch, numpy as np
from oracle import sam_ref
sd = sam_ref.seeded_state_dict("vit_test", seed=1)
def run(bf16):
    osam = sam_ref.build_sam("vit_test"); osam.load_state_dict(sd)
    for p in osam.parameters(): p.requires_grad_(True)
    gen = torch.Generator().manual_seed(3)
    P = 6
    emb = torch.randn(1, 256, 64, 64, generator=gen)
    xy = torch.rand(P, 2, generator=gen) * 600 + 50
    boxes = torch.cat([xy, xy + torch.rand(P, 2, generator=gen) * 300 + 20], 1)
    w_low = torch.randn(P, 3, 256, 256, generator=gen) / 256
    w_iou = torch.randn(P, 3, generator=gen)
    oemb = emb.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=bf16):
        sp, de = osam.prompt_encoder(points=None, boxes=boxes, masks=None)
        olow, oiou = osam.mask_decoder(image_embeddings=oemb, image_pe=osam.prompt_encoder.get_dense_pe(), sparse_prompt_embeddings=sp,
                                       dense_prompt_embeddings=de, multimask_output=True)
    ((olow.float() * w_low).sum() + (oiou.float() * w_iou).sum()).backward()
    g = {k: p.grad.clone() for k, p in osam.named_parameters() if p.grad is not None}
    g["emb"] = oemb.grad.clone()
    return g, olow.detach().float()
g32, l32 = run(False)
g16, l16 = run(True)
print("forward rel", float((l16-l32).norm()/l32.norm()))
rels = {k: float((g16[k].double()-g32[k].double()).norm()/(g32[k].double().norm()+1e-30)) for k in g32 if g32[k].norm() > 1e-6}
print("emb", rels["emb"], "median", np.median(list(rels.values())))
for k, v in sorted(rels.items(), key=lambda kv: -kv[1])[:12]: print(f"{v:.3e} {k}")
