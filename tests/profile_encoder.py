"""Driver for ncu: ViT encoder forward on a batch of tiles (random weights)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_sam_b200.sam import B200Sam  # noqa: E402
from oracle import sam_ref  # noqa: E402  (seeded weights only)

model = sys.argv[1] if len(sys.argv) > 1 else "vit_h"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sd = {k: v for k, v in sam_ref.seeded_state_dict(model, seed=0).items() if k.startswith("image_encoder.")}
sam = B200Sam(model, sd, max_batch=B, max_prompts=1)
x = torch.randint(0, 255, (B, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
sam.encode_u8(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
sam.encode_u8(x)
e1.record()
torch.cuda.synchronize()
print(f"{model} B={B}: {e0.elapsed_time(e1) / B:.3f} ms/tile")
torch.cuda.profiler.start()
sam.encode_u8(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
