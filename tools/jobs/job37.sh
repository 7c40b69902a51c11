nvidia-smi --query-gpu=index,name --format=csv,noheader | head -4
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/tiles2_final.json 2> gpurun_out/tiles2_final.err; grep '^{' gpurun_out/tiles2_final.json | tail -1 | cut -c1-260
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | grep '^{' | tail -1 | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload canvas --steps 40 2>/dev/null | grep '^{' | tail -1 | cut -c1-260
