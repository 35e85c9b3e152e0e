"""Driver for ncu: one fine-tuning step (cfg 5: TrainableSAM forward, loss, backward through decoder + encoder, AdamW) on B images."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_configs  # noqa: E402
from micro_sam_b200 import training, util  # noqa: E402
from oracle import sam_ref  # noqa: E402  (seeded weights only)

model = sys.argv[1] if len(sys.argv) > 1 else "vit_b"
pred = util.get_sam_model(model, state_dict=sam_ref.seeded_state_dict(model, seed=0), max_batch=2, max_prompts=64)
sam = pred.model.train()
m = training.TrainableSAM(sam)
recs, targets = bench_configs._cfg5_batch(100)


def step():
    sam.zero_decoder_grads()
    emb, rr = m.image_embeddings_oft([dict(r) for r in recs])
    loss = training.compute_loss(m(rr, emb, multimask_output=True, return_masks=False), targets)
    loss[0].backward()
    sam.optimizer_step(lr=1e-5)
    return float(loss[0])


step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
l = step()
e1.record()
torch.cuda.synchronize()
print(f"{model}: training step {e0.elapsed_time(e1):.2f} ms, loss {l:.4f}")
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
