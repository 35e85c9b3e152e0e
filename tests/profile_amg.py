"""Small driver for ncu: embed + AMG on `n` tiles (vit_b, 32x32 grid), after one untimed warm-up tile."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_sam_b200 import instance_segmentation as iseg, util  # noqa: E402
from micro_sam_b200.sample_data import lm_tile  # noqa: E402
from oracle import sam_ref  # noqa: E402  (seeded weights only)

model = sys.argv[1] if len(sys.argv) > 1 else "vit_b"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sd = sam_ref.seeded_state_dict(model, seed=0)
pred = util.get_sam_model(model, state_dict=sd, max_batch=4, max_prompts=1024)
amg = iseg.AutomaticMaskGenerator(pred, points_per_side=32)
tiles = np.stack([lm_tile((1024, 1024), 150, seed=i) for i in range(n + 1)])
for t in range(n + 1):
    if t == 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    amg.initialize(tiles[t])
    amg.generate_device(pred_iou_thresh=0.5, stability_score_thresh=0.125, box_nms_thresh=1.0)   # bench.py BENCH_THRESH
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
