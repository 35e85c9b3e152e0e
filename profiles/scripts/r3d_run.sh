mkdir -p gpurun_out
(for i in 1 2 3; do python tests/profile_encoder.py vit_h 8 | head -1; done; for i in 1 2 3; do MSAM_NO_PDL=1 python tests/profile_encoder.py vit_h 8 | head -1; done) > gpurun_out/r3d_pdl.log 2>&1
(time timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ops or amg_against or batched_inference") > gpurun_out/r3d_parity.log 2>&1
(time timeout 900 python -m pytest tests/test_gpu_real_arch.py -q -s -m gpu -k "vit_h or vit_b") > gpurun_out/r3d_parity2.log 2>&1
(time python bench.py --steps 3 --warmup 3) > gpurun_out/r3d_bench.log 2>&1
(time MSAM_NO_PDL=1 python bench.py --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/r3d_bench_nopdl.log 2>&1
(time python bench.py --config cfg1 --steps 5 --warmup 3) > gpurun_out/r3d_cfg1.log 2>&1
cat gpurun_out/r3d_pdl.log; tail -n 4 gpurun_out/r3d_parity.log; grep "rel-L2\|passed\|failed" gpurun_out/r3d_parity2.log
python - <<'PY'
import json
for f in ("gpurun_out/r3d_bench.log", "gpurun_out/r3d_bench_nopdl.log"):
    l=[x for x in open(f) if x.startswith('{')]
    if not l: print(f, "no json"); print(open(f).read()[-1500:]); continue
    d=json.loads(l[-1])
    print(f, d['value'], d['e2e']['value'], d.get('vit_h_encoder'))
    for r in d['roofline']['kernels'][:8]: print("   ", round(r['ms_per_step']/16,3), r['kernel'])
l=[x for x in open("gpurun_out/r3d_cfg1.log") if x.startswith('{')]
print(l[-1][:1500] if l else open("gpurun_out/r3d_cfg1.log").read()[-1500:])
PY
