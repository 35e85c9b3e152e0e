mkdir -p gpurun_out
python profiles/scripts/win_attn_probe.py 80 16 gpurun_out/win80_v2b.txt > gpurun_out/r3b_v2.log 2>&1
(time timeout 900 python -m pytest tests/test_gpu_real_arch.py -q -s -m gpu -x -k "vit_t or vit_h") > gpurun_out/r3b_parity.log 2>&1
python tests/profile_encoder.py vit_h 8 > gpurun_out/r3b_enc.log 2>&1
python tests/profile_encoder.py vit_t 8 > gpurun_out/r3b_enc_t.log 2>&1
(time timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "ops or to_image or amg_against") > gpurun_out/r3b_parity2.log 2>&1
tail -22 gpurun_out/r3b_v2.log; grep -v "^$" gpurun_out/r3b_parity.log | tail -25; tail -n 3 gpurun_out/r3b_enc.log gpurun_out/r3b_enc_t.log; tail -n 8 gpurun_out/r3b_parity2.log
