mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s -m gpu -k "loss_backward or decoder_train or end_to_end") > gpurun_out/r3g_bwd.log 2>&1
grep -v "^$" gpurun_out/r3g_bwd.log | grep -v "Warning\|warnings" | tail -70 | cut -c1-1200
