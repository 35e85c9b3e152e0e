mkdir -p gpurun_out
(time python -m pytest tests/test_gpu_parity.py -q -m gpu -x) > gpurun_out/r2e_parity.log 2>&1
(time python bench.py --steps 3 --warmup 3 --no-vith --no-cpu-baseline) > gpurun_out/r2e_bench.log 2>&1
(time python bench.py --config cfg3 --steps 1 --warmup 1) > gpurun_out/r2e_cfg3.log 2>&1
(time python tests/gpu_diag.py wgrad) > gpurun_out/r2e_wgrad.log 2>&1
tail -25 gpurun_out/r2e_parity.log; tail -c 1200 gpurun_out/r2e_bench.log; echo; tail -c 2500 gpurun_out/r2e_cfg3.log; tail -8 gpurun_out/r2e_wgrad.log
