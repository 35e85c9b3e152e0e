mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py -q -m gpu) > gpurun_out/r2d_parity.log 2>&1
(time python bench.py --steps 3 --warmup 3 --no-vith --no-cpu-baseline) > gpurun_out/r2d_bench.log 2>&1
(time python bench.py --config cfg4 --steps 2 --warmup 1) > gpurun_out/r2d_cfg4.log 2>&1
(time python bench.py --config cfg3 --steps 1 --warmup 1) > gpurun_out/r2d_cfg3.log 2>&1
tail -25 gpurun_out/r2d_parity.log; tail -c 1500 gpurun_out/r2d_bench.log; echo; tail -c 2500 gpurun_out/r2d_cfg4.log; echo; tail -c 2500 gpurun_out/r2d_cfg3.log
