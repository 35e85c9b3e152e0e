mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s -m gpu -k "optimizer or end_to_end or small_archs") > gpurun_out/r3j_bwd.log 2>&1
(time timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ops") > gpurun_out/r3j_ops.log 2>&1
(time python bench.py --config cfg5 --steps 5 --warmup 3) > gpurun_out/r3j_cfg5.log 2>&1
grep -n "AdamW\|training step\|vit_test\|passed\|failed\|Error" gpurun_out/r3j_bwd.log | cut -c1-700 | head; tail -n 3 gpurun_out/r3j_ops.log
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r3j_cfg5.log") if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'])
else: print(open("gpurun_out/r3j_cfg5.log").read()[-1500:])
PY
