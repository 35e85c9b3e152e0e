mkdir -p gpurun_out
(timeout 100 python -m pytest tests/test_gpu_parity.py -q -s -m gpu -k "rank_sharded") > gpurun_out/r3o.log 2>&1
grep -n "world \|passed\|failed\|Error\|assert" gpurun_out/r3o.log | cut -c1-600 | head -12
