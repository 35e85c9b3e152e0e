mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "decoder or amg or model_level or batched or trainable or interactive") > gpurun_out/r2f_parity.log 2>&1
(time python bench.py --steps 3 --warmup 3 --no-vith --no-cpu-baseline) > gpurun_out/r2f_bench.log 2>&1
(time python -m pytest tests/test_gpu_real_arch.py -q -s -m gpu -k vit_b) > gpurun_out/r2f_real.log 2>&1
tail -8 gpurun_out/r2f_parity.log; tail -c 900 gpurun_out/r2f_bench.log; echo; tail -8 gpurun_out/r2f_real.log
