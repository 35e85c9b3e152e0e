"""Probe: does replaying the encoder as a CUDA graph remove launch gaps?  (eager vs graph, ms per tile)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_sam_b200 import _lib  # noqa: E402
from micro_sam_b200.sam import B200Sam  # noqa: E402
from oracle import sam_ref  # noqa: E402  (seeded weights only)

for model, B in (("vit_h", 8), ("vit_b", 16)):
    sd = {k: v for k, v in sam_ref.seeded_state_dict(model, seed=0).items() if k.startswith("image_encoder.")}
    sam = B200Sam(model, sd, max_batch=B, max_prompts=1)
    x = torch.randint(0, 255, (B, 1024, 1024, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty(B, 256, 64, 64, device="cuda")
    L = _lib.lib()

    def enc():
        _lib.check(L.msam_encode_u8(sam._h, _lib.ptr(x), B, 1024, 1024, _lib.ptr(out), _lib.cur_stream()))

    def timeit(fn, n=4):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n / B

    t_eager = timeit(enc)
    ref = out.clone()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        enc()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            enc()
    t_graph = timeit(g.replay)
    print(f"{model} B={B}: eager {t_eager:.3f} ms/tile, graph {t_graph:.3f} ms/tile, same output {torch.equal(ref, out)}", flush=True)
    del sam
    torch.cuda.empty_cache()
