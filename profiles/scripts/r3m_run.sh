mkdir -p gpurun_out
nvidia-smi -L | head -3
(time timeout 1500 python -m pytest tests/test_gpu_multi.py -q -s -m gpu) > gpurun_out/r3m_multi.log 2>&1
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --config cfg5 --steps 5 --warmup 3) > gpurun_out/r3m_cfg5_2gpu.log 2>&1
tail -n 6 gpurun_out/r3m_multi.log | cut -c1-800
grep '^{' gpurun_out/r3m_cfg5_2gpu.log | tail -1 | cut -c1-700; tail -n 4 gpurun_out/r3m_cfg5_2gpu.log | cut -c1-500
