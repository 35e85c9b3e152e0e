mkdir -p gpurun_out
python profiles/scripts/win_attn_probe.py 80 16 gpurun_out/win80_v2c.txt > gpurun_out/r3c_80.log 2>&1
python profiles/scripts/win_attn_probe.py 64 12 gpurun_out/win64_v2c.txt > gpurun_out/r3c_64.log 2>&1
(time timeout 1200 python -m pytest tests/test_gpu_real_arch.py -q -s -m gpu) > gpurun_out/r3c_parity.log 2>&1
for i in 1 2 3; do python tests/profile_encoder.py vit_h 8 | head -1; done > gpurun_out/r3c_enc.log 2>&1
python tests/profile_encoder.py vit_t 8 | head -1 >> gpurun_out/r3c_enc.log 2>&1
python tests/profile_encoder.py vit_l 8 | head -1 >> gpurun_out/r3c_enc.log 2>&1
(time timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ops or to_image or amg_against") > gpurun_out/r3c_parity2.log 2>&1
grep -h "parity\|timing\|mean" gpurun_out/r3c_80.log gpurun_out/r3c_64.log; grep -v "^$" gpurun_out/r3c_parity.log | grep "rel-L2\|passed\|failed\|Error\|error\|assert" | tail -25; cat gpurun_out/r3c_enc.log; tail -n 6 gpurun_out/r3c_parity2.log
