"""Window-attention probe (run on a B200): parity of the shipping kernel against a torch fp32 restatement at the ViT-H shape
(16 heads x head_dim 80, all 25 windows of an image), timing at batch 8, and the per-CTA phase timeline
(%globaltimer stamps written by softmax thread 0 of the first 64 windows of head 0, q-tile 0).
MSAM_WIN_V1=1 selects the first-generation kernel (attention.cu) for the same measurements.
Usage: python profiles/scripts/win_attn_probe.py [hd] [heads] [out.txt]"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from micro_sam_b200 import _lib  # noqa: E402
from tests.gpu_diag import attn_ref  # noqa: E402

hd = int(sys.argv[1]) if len(sys.argv) > 1 else 80
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out_path = sys.argv[3] if len(sys.argv) > 3 else None
DEV, S, G = "cuda", 14, 196
D = heads * hd
L = _lib.lib()
L.msam_debug_attn_trace.argtypes = [ctypes.c_void_p]
lines = []


def say(s):
    print(s, flush=True)
    lines.append(s)


def make(B, seed, qscale=1.0):
    g = torch.Generator().manual_seed(seed)
    qkv = (torch.randn(B * 25 * G, 3 * D, generator=g) * qscale).to(DEV).bfloat16()
    rel_h = (torch.randn(2 * S - 1, hd, generator=g) * 0.3).to(DEV).bfloat16()
    rel_w = (torch.randn(2 * S - 1, hd, generator=g) * 0.3).to(DEV).bfloat16()
    cols = ((hd + 63) // 64) * 64
    tab = torch.zeros(64, cols, device=DEV, dtype=torch.bfloat16)
    tab[: 2 * S - 1, :hd] = rel_h
    tab[32: 32 + 2 * S - 1, :hd] = rel_w
    return qkv, rel_h, rel_w, tab


def run(qkv, tab, B):
    out = torch.zeros(B * 4096, D, device=DEV, dtype=torch.bfloat16)
    _lib.check(L.msam_op_attention(_lib.ptr(qkv), _lib.ptr(tab), _lib.ptr(out), B, heads, hd, 14, hd ** -0.5, _lib.cur_stream()))
    return out


say(f"kernel: {'v1 (attention.cu)' if os.environ.get('MSAM_WIN_V1') else 'shipping'}  hd={hd} heads={heads}")
# ---- parity (one image, all windows incl. the partially padded ones), two logit magnitudes
for qscale in (1.0, 3.0):
    qkv, rel_h, rel_w, tab = make(1, 1, qscale)
    out = run(qkv, tab, 1)
    torch.cuda.synchronize()
    ref = attn_ref(qkv.float(), rel_h.float(), rel_w.float(), 1, heads, hd, S, None)
    ref = ref.view(1, 5, 5, 14, 14, D).permute(0, 1, 3, 2, 4, 5).reshape(1, 70, 70, D)[:, :64, :64].reshape(-1, D)
    rel = ((out.float() - ref).norm() / ref.norm()).item()
    worst = max(((out[:, h * hd:(h + 1) * hd].float() - ref[:, h * hd:(h + 1) * hd]).norm() / ref[:, h * hd:(h + 1) * hd].norm()).item()
                for h in range(heads))
    say(f"parity qscale={qscale}: rel_l2={rel:.3e} worst head={worst:.3e} max_abs={(out.float() - ref).abs().max().item():.3e} "
        f"finite={bool(torch.isfinite(out.float()).all())}")

# ---- timing at batch 8 (one launch = one windowed block of the encoder)
B = 8
qkv, _, _, tab = make(B, 2)
for _ in range(3):
    run(qkv, tab, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    run(qkv, tab, B)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / n * 1e3
fl = B * 25 * heads * (4.0 * G * G * hd + 4.0 * G * S * hd)
say(f"timing B={B}: {us:.1f} us per launch = {us / B:.2f} us per tile-block, {fl / us * 1e-6:.1f} TFLOP/s algorithmic")

# ---- timeline
buf = torch.zeros(64 * 16, dtype=torch.int64, device=DEV)
L.msam_debug_attn_trace(_lib.ptr(buf))
run(qkv, tab, B)
torch.cuda.synchronize()
L.msam_debug_attn_trace(None)
t = buf.view(64, 16).cpu()
names = ["start", "setup", "t_full", "t_done", "s_full", "max", "exp", "p_arr", "o_full", "end"]
say("timeline (ns since CTA start; softmax thread 0 of window w, head 0, q-tile 0): " + " ".join(names[1:]))
acc = torch.zeros(10)
cnt = 0
for w in range(64):
    if t[w, 0] == 0:
        continue
    d = (t[w, :10] - t[w, 0]).float()
    acc += d
    cnt += 1
    if w < 8:
        say(f"  w{w:02d} start@{int(t[w, 0] - t[:, 0][t[:, 0] > 0].min())}: " + " ".join(f"{int(x):6d}" for x in d[1:]))
if cnt:
    m = acc / cnt
    say("  mean: " + " ".join(f"{int(x):6d}" for x in m[1:]))
    say("  mean phase lengths: " + " ".join(f"{names[i]}={int(m[i] - m[i - 1])}" for i in range(1, 10)))
if out_path:
    with open(out_path, "w") as f:
        f.write("\n".join(lines) + "\n")
