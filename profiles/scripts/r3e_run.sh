mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s -m gpu -k "bgemm or layernorm or small_archs") > gpurun_out/r3e_bwd.log 2>&1
grep -v "^$" gpurun_out/r3e_bwd.log | grep -v "Warning\|warnings" | tail -60
