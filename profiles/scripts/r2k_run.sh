mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "decoder or amg or model_level or batched or trainable") > gpurun_out/r2m_parity.log 2>&1
MSAM_I2T_TRACE=gpurun_out/i2t_trace_v5.txt python tests/profile_amg.py vit_b 1 > gpurun_out/r2m_t.log 2>&1
(time python bench.py --steps 3 --warmup 3 --no-vith --no-cpu-baseline) > gpurun_out/r2m_bench.log 2>&1
tail -5 gpurun_out/r2m_parity.log; sed -n 10,16p gpurun_out/i2t_trace_v5.txt; python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2m_bench.log') if x.startswith('{')][-1]
d=json.loads(l)
print(d['value'], d['e2e']['value'])
for r in d['roofline']['kernels'][:6]: print(round(r['ms_per_step']/16,3), r['kernel'])
PY
