mkdir -p gpurun_out
(time python -m pytest tests/ -x -q -m gpu) > gpurun_out/r2i_pytest_all.log 2>&1
(time python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2i_smoke.log 2>&1
(time python bench.py --steps 3 --warmup 3) > gpurun_out/r2i_bench.log 2>&1
(time python bench.py --impl reference --steps 1 --warmup 0) > gpurun_out/r2i_bench_ref.log 2>&1
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2i_launches_amg.csv python tests/profile_amg.py vit_b 1 > gpurun_out/r2i_l.log 2>&1
$NCU --set full --import-source on -k regex:paint_min_area_x4 -c 1 -o gpurun_out/r2i_paint -f python tests/profile_amg.py vit_b 1 > gpurun_out/r2i_p.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2i_launches_vith.csv python tests/profile_encoder.py vit_h 4 > gpurun_out/r2i_lh.log 2>&1
tail -6 gpurun_out/r2i_pytest_all.log; tail -2 gpurun_out/r2i_smoke.log; tail -c 700 gpurun_out/r2i_bench.log; echo; tail -c 500 gpurun_out/r2i_bench_ref.log
