# multi-GPU evidence (run under gpurun --gpus N): sharded-canvas bit-exactness + strong scaling of the canvas workloads
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q 2>&1 | tail -3
for n in 1 2 4 8; do
  if [ $n -le $N ]; then
    if [ $n -le 4 ]; then
      if [ $n -eq 1 ]; then python bench.py --workload canvas --steps 40 > gpurun_out/canvas_n$n.json 2> gpurun_out/canvas_n$n.err
      else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --workload canvas --steps 40 > gpurun_out/canvas_n$n.json 2> gpurun_out/canvas_n$n.err; fi
      tail -1 gpurun_out/canvas_n$n.json | cut -c1-400
    fi
    if [ $n -eq 1 ]; then python bench.py --workload export --solve-steps 1 --steps 1 > gpurun_out/export_n$n.json 2> gpurun_out/export_n$n.err
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --workload export --solve-steps 1 --steps 1 > gpurun_out/export_n$n.json 2> gpurun_out/export_n$n.err; fi
    tail -1 gpurun_out/export_n$n.json | cut -c1-400
    if [ "${SKIP_TILES:-0}" = "1" ]; then continue; fi
    if [ $n -eq 1 ]; then python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/tiles_n$n.json 2>/dev/null
    else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 40 --warmup 5 > gpurun_out/tiles_n$n.json 2>/dev/null; fi
    tail -1 gpurun_out/tiles_n$n.json | cut -c1-200
  fi
done
