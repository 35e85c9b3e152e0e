#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 1024x1024 tiles/s for embedding precompute + AMG (vit_b, 32x32 point grid, batch of 16
synthetic LM tiles = BASELINE.json configs[1]) on N B200s, plus the ViT-H encoder forward as a fraction of the bf16
tensor-core roofline.

  python bench.py --gpus N --steps K --warmup W            # ours (torchrun for N > 1, one rank per GPU, weak scaling)
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm's CPU path (oracle port) on host cores

One "step" = one pass of the hot path over one batch of 16 tiles per GPU.
  value : tiles/s, inputs resident in HBM (uint8 tiles on the device), device-side AMG result (painted label image)
  e2e   : tiles/s through the reference-facing API (precompute_image_embeddings + AutomaticMaskGenerator.initialize /
          generate) from HOST uint16 tiles to HOST uint32 label images; H2D / D2H inside the timed region.
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.  Every step streams
multi-GB decoder activations (>> 126 MB L2), so no extra L2 flush is needed ("l2": "working_set_exceeds_l2").
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TILES = 16
TILE = 1024
GRID = 32
# algorithmic FLOPs
ENC_FLOPS = {"vit_b": 0.9376e12, "vit_l": 2.8370e12, "vit_h": 5.6418e12}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(dev), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 9]
        mx = [float(r[2]) for r in rows if len(r) >= 9]
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_tiles(seed0=0):
    from micro_sam_b200.sample_data import lm_tile
    return np.stack([lm_tile((TILE, TILE), 150, seed=seed0 + i) for i in range(N_TILES)])


def best_cpu_threads():
    """All host threads are allowed, but PyTorch's CPU GEMMs peak below the logical-core count on SMT hosts: pick the
    thread count that runs an encoder-sized matmul fastest (candidates: all, 1/2, 1/4 of the logical cores)."""
    n = os.cpu_count() or 1
    a, b = torch.randn(4096, 768), torch.randn(768, 3072)
    best, best_t = n, None
    for c in sorted({n, max(1, n // 2), max(1, n // 4)}, reverse=True):
        torch.set_num_threads(c)
        a @ b
        t0 = time.perf_counter()
        for _ in range(3):
            a @ b
        t = time.perf_counter() - t0
        if best_t is None or t < best_t:
            best, best_t = c, t
    return best


# ---------------------------------------------------------------------------------------------------- reference arm
def cpu_baseline(model_type="vit_b", n_point_batches=2, threads=None):
    """The reference algorithm's CPU path (oracle port of segment_anything + micro-sam's AMG) on the host cores, on a
    BOUNDED sample of the same workload: 1 tile embedding + `n_point_batches` x 64 grid points through predict_torch /
    _to_mask_data, + generate; scaled to the 16-batch (1024 point) grid."""
    from oracle import amg_ref, sam_ref
    threads = threads or best_cpu_threads()
    torch.set_num_threads(threads)
    sam = sam_ref.build_seeded_sam(model_type, seed=0)
    pred = sam_ref.SamPredictor(sam)
    img = make_tiles(0)[0]
    t0 = time.perf_counter()
    emb = amg_ref.precompute_image_embeddings_2d(pred, img)
    t_embed = time.perf_counter() - t0
    amg = amg_ref.AutomaticMaskGenerator(pred, points_per_side=GRID, points_per_batch=64)
    amg.point_grids = [amg.point_grids[0][: 64 * n_point_batches]]
    t0 = time.perf_counter()
    amg.initialize(img, image_embeddings=emb)
    t_init = time.perf_counter() - t0
    t0 = time.perf_counter()
    amg.generate(pred_iou_thresh=0.0, stability_score_thresh=0.0)
    t_gen = time.perf_counter() - t0
    scale = (GRID * GRID) / (64 * n_point_batches)
    per_tile = t_embed + t_init * scale + t_gen * scale
    return {
        "value": 1.0 / per_tile, "unit": "tiles/s", "cores": threads, "kind": "port",
        "sample": f"1 tile: embed {t_embed:.2f}s + {n_point_batches}x64 of 1024 grid points {t_init:.2f}s + generate {t_gen:.2f}s, "
                  f"AMG part scaled x{scale:.0f}",
        "seconds_per_tile": per_tile,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(max(1, args.warmup > 0)):
        cpu_baseline(args.model, 1)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        cb = cpu_baseline(args.model, 1)
        vals.append(cb)
    v = float(np.mean([c["value"] for c in vals]))
    cb = vals[-1]
    cb["value"] = v
    out = {
        "impl": "reference", "metric": "1024x1024 tiles/s, embed + AMG", "value": v, "unit": "tiles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (time.perf_counter() - t_all) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.model} AMG 32x32 grid, 1024^2 LM tiles (bounded sample per step, scaled)"},
        "cpu_baseline": cb, "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------- our arm
def vit_h_roofline(device, steps=3):
    """ViT-H encoder forward: ms/tile and fraction of the bf16 tensor roofline.  Measured at batch 4 and 8 tiles (distinct
    uint8 tiles each step, so nothing is cached): at batch 4 the fp32 residual stream (84 MB) stays in the 126 MB L2,
    at batch 8 it does not -- the better of the two is reported together with its batch."""
    from oracle import sam_ref  # weights only (seeded generator); nothing of the oracle is timed here
    from micro_sam_b200.sam import B200Sam
    sd = {k: v for k, v in sam_ref.seeded_state_dict("vit_h", seed=0).items() if k.startswith("image_encoder.")}
    sam = B200Sam("vit_h", sd, device=device, max_batch=8, max_prompts=1)
    best = None
    for batch in (8, 4):
        xs = [torch.randint(0, 255, (batch, TILE, TILE, 3), dtype=torch.uint8, device=device) for _ in range(steps)]
        for _ in range(2):
            sam.encode_u8(xs[0])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            sam.encode_u8(xs[k])
        e1.record()
        torch.cuda.synchronize()
        ms_tile = e0.elapsed_time(e1) / steps / batch
        if best is None or ms_tile < best[0]:
            best = (ms_tile, batch)
    del sam
    torch.cuda.empty_cache()
    return best


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from oracle import sam_ref  # seeded weight generator only
    from micro_sam_b200 import _lib, instance_segmentation as iseg, util
    pk, pk_src = peaks()

    sd = sam_ref.seeded_state_dict(args.model, seed=0)
    pred = util.get_sam_model(args.model, device=device, state_dict=sd, max_batch=args.enc_batch, max_prompts=args.max_prompts)
    sam = pred.model
    amg = iseg.AutomaticMaskGenerator(pred, points_per_side=GRID)
    tiles = make_tiles(seed0=rank * N_TILES)                                   # host uint16 (16,1024,1024)
    tiles_u8 = torch.from_numpy(np.stack([util._to_image(t) for t in tiles])).to(device)   # device-resident inputs
    gen_kw = dict(pred_iou_thresh=args.pred_iou_thresh, stability_score_thresh=args.stability_score_thresh)

    def step_device():
        feats = sam.encode_u8(tiles_u8)
        out = None
        for t in range(N_TILES):
            emb = {"features": feats[t:t + 1], "input_size": (TILE, TILE), "original_size": (TILE, TILE)}
            amg.initialize(tiles[t], image_embeddings=emb)
            out = amg.generate_device(**gen_kw)
        return out

    def step_e2e():
        emb = util.precompute_image_embeddings(pred, tiles, ndim=3, batch_size=N_TILES, to_numpy=False)
        segs = []
        for z in range(N_TILES):
            amg.initialize(tiles[z], image_embeddings=emb, i=z)
            segs.append(amg.generate(**gen_kw))
        return segs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1e3], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t[0]), float(t[1])

    L = _lib.lib()
    # ---- device-resident throughput (value) with per-kernel event timing for the roofline
    for _ in range(args.warmup):
        step_device()
    L.msam_profile(1)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = _lib.launch_count()
    dev_ms, _ = timed(step_device, args.steps, 0)
    launches = (_lib.launch_count() - l0) / args.steps
    clocks = sampler.stop() if sampler else None
    import ctypes
    prof = (ctypes.c_double * 9)()
    _lib.check(L.msam_profile_summary(prof))
    L.msam_profile(0)
    value = world * N_TILES * args.steps / (dev_ms / 1e3)
    # ---- end to end through the reference-facing API (host in, host out)
    e2e_ms, e2e_wall = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    e2e_time = max(e2e_ms, e2e_wall)  # host work after the last kernel is part of the step
    e2e_val = world * N_TILES * args.steps / (e2e_time / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    gemm_ms, gemm_flops, gemm_n = prof[0], prof[1], prof[2]
    att_ms, att_flops, att_n = prof[3], prof[4], prof[5]
    hbm_ms, hbm_bytes, hbm_n = prof[6], prof[7], prof[8]
    # dominant kernels of the step: the decoder's image-side kernels (up-scaling GEMMs with fused LN / GELU / hyper-product
    # epilogues and the two fused cross-attention blocks) -> algorithmic bytes / CUDA-event time vs the measured HBM peak
    hbm_gbs = hbm_bytes / (hbm_ms * 1e-3) / 1e9 if hbm_ms > 0 else 0.0
    enc_tf = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roof = {"bound": "hbm", "kernel": "decoder image-side kernels: gemm_bf16_kernel K<512 (conv-transpose GEMMs with fused LN2d+GELU / "
                                     "GELU+hyper-product epilogues), i2t_fused_kernel, t2i_fused_kernel",
            "achieved": hbm_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_gbs / pk["hbm_gbs"], "traffic": None,
            "peak_source": f"{pk_src} hbm_gbs", "launches_per_step": hbm_n / args.steps, "share_of_step": hbm_ms / dev_ms,
            "algorithmic_bytes_per_step": hbm_bytes / args.steps,
            "encoder_gemm": {"bound": "tensor", "achieved": enc_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                             "frac": enc_tf / pk["bf16_tflops_sustained"], "launches_per_step": gemm_n / args.steps,
                             "share_of_step": gemm_ms / dev_ms},
            "attention": {"ms_per_step": att_ms / args.steps, "tflops": att_flops / max(att_ms, 1e-9) / 1e9,
                          "launches_per_step": att_n / args.steps, "share_of_step": att_ms / dev_ms}}
    out = {
        "metric": "1024x1024 tiles/s, embed + AMG", "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.model} AutomaticMaskGenerator, 32x32 point grid, batch of 16 synthetic 1024x1024 LM tiles "
                               "per GPU (BASELINE.json configs[1]); random-init weights", "tiles_per_step_per_gpu": N_TILES,
                   "pred_iou_thresh": args.pred_iou_thresh, "stability_score_thresh": args.stability_score_thresh,
                   "l2": "working_set_exceeds_l2", "parallelism": f"tile-sharded x{world}, no collective"},
        "e2e": {"value": e2e_val, "unit": "tiles/s", "ms_per_step": e2e_time / args.steps,
                "h2d_bytes_per_step": int(N_TILES * TILE * TILE * 3 + N_TILES * GRID * GRID * 12),
                "d2h_bytes_per_step": int(N_TILES * TILE * TILE * 4)},
        "gpu_launches": launches, "clocks": clocks, "roofline": roof,
        "survivors_last_tile": int(amg._n_keep_dev.item()) if hasattr(amg, "_n_keep_dev") else None,
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.model, 1)
    if world == 1 and not args.no_vith:
        del pred, sam, amg
        torch.cuda.empty_cache()
        ms_tile, vith_batch = vit_h_roofline(device)
        tf = ENC_FLOPS["vit_h"] / (ms_tile * 1e-3) / 1e12
        out["vit_h_encoder"] = {"ms_per_tile": ms_tile, "tflops": tf, "frac_of_peak": tf / pk["bf16_tflops_sustained"],
                                "peak": pk["bf16_tflops_sustained"], "batch": vith_batch, "algorithmic_tflop_per_tile": 5.6418}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="vit_b")
    ap.add_argument("--max-prompts", type=int, default=1024)
    ap.add_argument("--enc-batch", type=int, default=16, help="tiles per encoder pass (the engine chunks the 16-tile batch)")
    ap.add_argument("--pred-iou-thresh", type=float, default=0.88)
    ap.add_argument("--stability-score-thresh", type=float, default=0.95)
    ap.add_argument("--no-vith", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the B200 path has no CPU fallback; use --impl reference)")
        run_ours(args)


if __name__ == "__main__":
    main()
