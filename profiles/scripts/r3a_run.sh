mkdir -p gpurun_out
python profiles/scripts/win_attn_probe.py 80 16 gpurun_out/win80_v2.txt > gpurun_out/r3a_v2.log 2>&1
MSAM_WIN_V1=1 python profiles/scripts/win_attn_probe.py 80 16 gpurun_out/win80_v1.txt > gpurun_out/r3a_v1.log 2>&1
python profiles/scripts/win_attn_probe.py 64 12 gpurun_out/win64_v1.txt > gpurun_out/r3a_64.log 2>&1
python tests/profile_encoder.py vit_h 8 > gpurun_out/r3a_enc_v2.log 2>&1
MSAM_WIN_V1=1 python tests/profile_encoder.py vit_h 8 > gpurun_out/r3a_enc_v1.log 2>&1
(time timeout 600 python -m pytest tests/test_gpu_real_arch.py -q -m gpu -x -k vit_h) > gpurun_out/r3a_parity.log 2>&1
tail -30 gpurun_out/r3a_v2.log; tail -16 gpurun_out/r3a_v1.log; tail -16 gpurun_out/r3a_64.log; tail -2 gpurun_out/r3a_enc_v2.log gpurun_out/r3a_enc_v1.log; tail -5 gpurun_out/r3a_parity.log
