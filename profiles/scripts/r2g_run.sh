mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py -q -m gpu -x) > gpurun_out/r2g_parity.log 2>&1
(time python bench.py --steps 3 --warmup 3 --no-vith --no-cpu-baseline) > gpurun_out/r2g_bench.log 2>&1
(time python -m pytest tests/test_gpu_real_arch.py -q -s -m gpu -k vit_b) > gpurun_out/r2g_real.log 2>&1
(time python bench.py --config cfg3 --steps 1 --warmup 1) > gpurun_out/r2g_cfg3.log 2>&1
tail -8 gpurun_out/r2g_parity.log; tail -c 900 gpurun_out/r2g_bench.log; echo; tail -8 gpurun_out/r2g_real.log; tail -c 600 gpurun_out/r2g_cfg3.log
