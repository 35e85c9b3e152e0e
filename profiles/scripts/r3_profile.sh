mkdir -p gpurun_out
NCU="ncu --profile-from-start off --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r3_launches_vit_h_4tiles.csv python tests/profile_encoder.py vit_h 4 > gpurun_out/r3p_l1.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/r3_launches_train_vit_b.csv python tests/profile_train.py vit_b > gpurun_out/r3p_l2.log 2>&1
$NCU --set full --import-source on -k regex:attn_window2 -c 1 -o gpurun_out/r3_win2_d80 -f python tests/profile_encoder.py vit_h 4 > gpurun_out/r3p_w.log 2>&1
$NCU --set full --import-source on -k regex:bgemm_kernel -s 8 -c 1 -o gpurun_out/r3_bgemm -f python tests/profile_train.py vit_b > gpurun_out/r3p_b.log 2>&1
$NCU --set full --import-source on -k regex:attn_probs -s 3 -c 1 -o gpurun_out/r3_attn_probs -f python tests/profile_train.py vit_b > gpurun_out/r3p_p.log 2>&1
$NCU --set full --import-source on -k regex:gemm_tn_kernel -s 4 -c 1 -o gpurun_out/r3_gemm_tn -f python tests/profile_train.py vit_b > gpurun_out/r3p_t.log 2>&1
tail -2 gpurun_out/r3p_l1.log gpurun_out/r3p_l2.log; ls -la gpurun_out/r3_*
(time timeout 600 python -m pytest tests/test_gpu_backward.py -q -s -m gpu -k "optimizer") > gpurun_out/r3p_opt.log 2>&1
grep -n "AdamW\|passed\|failed" gpurun_out/r3p_opt.log | cut -c1-600 | head -5
