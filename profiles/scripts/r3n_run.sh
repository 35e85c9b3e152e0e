mkdir -p gpurun_out
(timeout 240 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py -q -s -m gpu -k "small_archs or precompute_state_round or rank_sharded or end_to_end") > gpurun_out/r3n.log 2>&1
grep -n "world \|vit_test\|training step\|passed\|failed\|Error\|assert" gpurun_out/r3n.log | cut -c1-900 | head -30
