mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_backward.py -q -s -m gpu -k "vit_b") > gpurun_out/r3f_bwd.log 2>&1
(time python bench.py --config cfg5 --steps 5 --warmup 3) > gpurun_out/r3f_cfg5.log 2>&1
(time timeout 1200 python bench.py --impl reference --config cfg5 --steps 1 --warmup 0) > gpurun_out/r3f_cfg5_ref.log 2>&1
grep -v "^$" gpurun_out/r3f_bwd.log | grep -v "Warning\|warnings" | tail -12
tail -n 4 gpurun_out/r3f_cfg5.log | cut -c1-3000; tail -n 4 gpurun_out/r3f_cfg5_ref.log | cut -c1-1500
