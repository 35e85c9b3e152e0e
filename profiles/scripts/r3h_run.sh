mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s -m gpu -k "decoder_train or end_to_end") > gpurun_out/r3h_bwd.log 2>&1
grep -n "decoder_train \|training step\|passed\|failed\|AssertionError: " gpurun_out/r3h_bwd.log | cut -c1-1800 | head -20
head -40 gpurun_out/decoder_grads_boxes.txt
