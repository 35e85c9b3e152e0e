mkdir -p gpurun_out
nvidia-smi -L
(time python -m pytest tests/test_gpu_multi.py -q -m gpu -x) > gpurun_out/r2h_multi.log 2>&1
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/r2h_bench2.log 2>&1
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --config cfg4 --gpus 2 --steps 2 --warmup 1) > gpurun_out/r2h_cfg4_2.log 2>&1
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --config cfg3 --gpus 2 --steps 1 --warmup 1) > gpurun_out/r2h_cfg3_2.log 2>&1
tail -6 gpurun_out/r2h_multi.log; for f in bench2 cfg4_2 cfg3_2; do grep '^{' gpurun_out/r2h_$f.log | python -c "import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['metric'][:40], d['n_gpus'], round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; done
