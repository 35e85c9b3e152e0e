"""torchrun worker (world_size >= 2, one rank per GPU, NCCL): data-parallel fine-tuning step (cfg 5).  Every rank runs the training
step on ITS batch (seed = rank) and the flat gradient buffer is all-reduced (averaged) -- torch DDP's semantics, which is what
micro_sam/training/training.py:train_sam uses.  Rank 0 then recomputes every rank's batch locally and checks that the average of
those gradients equals the all-reduced buffer.  Run by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def batch(seed, H=128, W=128, n_obj=3):
    from micro_sam_b200.sample_data import lm_tile
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    rng = np.random.default_rng(seed)
    recs, targets = [], []
    for b in range(2):
        img = torch.from_numpy(np.repeat(lm_tile((H, W), 12, seed=100 * seed + b, dtype="uint8")[None], 3, 0).astype("float32"))
        cen = [(int(rng.integers(20, H - 20)), int(rng.integers(20, W - 20)), int(rng.integers(6, 14))) for _ in range(n_obj)]
        targets.append(torch.stack([(((yy - cy) ** 2 + (xx - cx) ** 2) < r * r).float()[None] for cy, cx, r in cen]))
        boxes = torch.tensor([[cx - r, cy - r, cx + r, cy + r] for cy, cx, r in cen], dtype=torch.float32) * (1024.0 / W)
        recs.append({"image": img, "original_size": (H, W), "boxes": boxes})
    return recs, targets


def step(sam, m, seed):
    from micro_sam_b200 import training
    recs, targets = batch(seed)
    sam.zero_decoder_grads()
    emb, rr = m.image_embeddings_oft([dict(r) for r in recs])
    loss = training.compute_loss(m(rr, emb, multimask_output=True, return_masks=False), targets)
    loss[0].backward()
    return torch.cat([v.reshape(-1).clone() for _, v in sam.grad_views()]), float(loss[0])


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import sam_ref
    from micro_sam_b200 import training, util
    sd = sam_ref.seeded_state_dict("vit_test", seed=1)
    pred = util.get_sam_model("vit_test", device=f"cuda:{local}", state_dict=sd, max_batch=2, max_prompts=64)
    sam = pred.model.train()
    m = training.TrainableSAM(sam)
    _, loss = step(sam, m, rank)
    sam.allreduce_grads(world)
    flat = torch.cat([v.reshape(-1).clone() for _, v in sam.grad_views()])
    ok, rel = 1, 0.0
    if rank == 0:
        ref = sum(step(sam, m, r)[0] for r in range(world)) / world
        rel = float((flat - ref).norm() / ref.norm())
        ok = int(rel < 1e-3 and float(ref.norm()) > 0)       # atomics / reduction order only
    t = torch.tensor([ok], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.savez(sys.argv[1], ok=int(t.item()), rel=rel, world=world, n=flat.numel(), loss=loss)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
