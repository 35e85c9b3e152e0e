"""torchrun worker (world_size >= 2, one rank per GPU, NCCL): TiledAutomaticMaskGenerator with rank-sharded tiles + the
all-gather of the instance tables must reproduce the single-process result bit for bit.  Run by tests/test_gpu_multi.py:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tests/dist_tiled_amg.py OUT.npz
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import sam_ref
    from micro_sam_b200 import instance_segmentation as iseg, util
    from micro_sam_b200.sample_data import lm_tile
    sd = sam_ref.seeded_state_dict("vit_test", seed=1)
    pred = util.get_sam_model("vit_test", device=f"cuda:{local}", state_dict=sd, max_batch=4, max_prompts=64)
    img = lm_tile((500, 700), 40, seed=13)
    # 2 x 2 tiles with halos that reach the image border: the near-crop-edge filter would otherwise remove every near-full-crop mask of
    # the random-init model at the interior tile edges (an empty result compares equal vacuously)
    tile_shape, halo = (250, 350), (250, 350)
    kw = dict(pred_iou_thresh=0.0, stability_score_thresh=0.0, crop_nms_thresh=0.3, with_background=False)   # keeps masks for any seeded noise model
    amg = iseg.TiledAutomaticMaskGenerator(pred, points_per_side=4)
    amg.initialize(img, tile_shape=tile_shape, halo=halo, batch_size=2, rank=rank, world_size=world)
    assert len(amg.crop_list) == (4 * (rank + 1)) // world - (4 * rank) // world
    seg = amg.generate(**kw)
    # every rank holds the full result; compare with the single-process path computed on this rank
    ref_amg = iseg.TiledAutomaticMaskGenerator(pred, points_per_side=4)
    ref_amg.initialize(img, tile_shape=tile_shape, halo=halo, batch_size=2)
    ref = ref_amg.generate(**kw)
    ok = torch.tensor([int(np.array_equal(seg, ref) and seg.max() > 0)], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.savez(sys.argv[1], ok=int(ok.item()), n_instances=int(seg.max()), world=world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
