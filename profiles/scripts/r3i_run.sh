mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s -m gpu -k "decoder_train or end_to_end") > gpurun_out/r3i_bwd.log 2>&1
(time python bench.py --config cfg5 --steps 5 --warmup 3) > gpurun_out/r3i_cfg5.log 2>&1
(time timeout 1200 python bench.py --impl reference --config cfg5 --steps 1 --warmup 0) > gpurun_out/r3i_cfg5_ref.log 2>&1
grep -n "decoder_train \|training step\|passed\|failed" gpurun_out/r3i_bwd.log | cut -c1-600
grep '^{' gpurun_out/r3i_cfg5.log | tail -1 | cut -c1-3500; tail -3 gpurun_out/r3i_cfg5.log | cut -c1-600
grep -o '"cpu_baseline.*' gpurun_out/r3i_cfg5_ref.log | cut -c1-400
