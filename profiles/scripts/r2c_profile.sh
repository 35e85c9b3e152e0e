mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2c_launches_amg.csv python tests/profile_amg.py vit_b 1 > gpurun_out/r2c_l.log 2>&1
$NCU --set full --import-source on -k regex:upscale_fused -c 1 -o gpurun_out/r2c_upscale -f python tests/profile_amg.py vit_b 1 > gpurun_out/r2c_u.log 2>&1
$NCU --set full --import-source on -k regex:mask_stats_x4 -c 1 -o gpurun_out/r2c_stats -f python tests/profile_amg.py vit_b 1 > gpurun_out/r2c_s.log 2>&1
$NCU --set full --import-source on -k regex:i2t_fused -s 1 -c 1 -o gpurun_out/r2c_i2t -f python tests/profile_amg.py vit_b 1 > gpurun_out/r2c_i.log 2>&1
(time python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "amg or segmentation or batched") > gpurun_out/r2c_parity.log 2>&1
(time python bench.py --steps 3 --warmup 3 --no-vith) > gpurun_out/r2c_bench.log 2>&1
tail -5 gpurun_out/r2c_parity.log; tail -c 600 gpurun_out/r2c_bench.log; ls -la gpurun_out/r2c*
