"""2+-GPU check (torchrun): the striped multi-GPU canvas equals the single-GPU canvas bit for bit.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sharded_gpu.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from oracle import unet as O
from terrain_diffusion_b200.inference import sample_decoder_diffusion_sharded, sample_decoder_diffusion_tiled
from terrain_diffusion_b200.models import EDMUnet2D
from terrain_diffusion_b200.scheduler import EDMDPMSolverMultistepScheduler


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = O.DECODER_CFG
    m = EDMUnet2D(**cfg).eval()
    m.load_state_dict(O.procedural_state_dict(cfg, seed=0))
    m = m.to(dev)
    g = torch.Generator().manual_seed(3)
    h = w = 64 + 48 * 7          # 8 tile rows: stripes of 4 (2 ranks) or 2 (4 ranks) rows
    noise = (torch.randn(1, 1, h, w, generator=g) * 80).to(dev)
    cond = torch.randn(1, 4, h, w, generator=g).to(dev)
    own, (lo, hi) = sample_decoder_diffusion_sharded(m, EDMDPMSolverMultistepScheduler(), cond, noise, 64, 48,
                                                     num_steps=4, tile_batch=4)
    ref = sample_decoder_diffusion_tiled(m, EDMDPMSolverMultistepScheduler(), cond, noise, 64, 48, num_steps=4,
                                         tile_batch=4)[0]
    same = torch.equal(own, ref[:, lo:hi])
    print(f"rank {rank}: rows [{lo},{hi}) sharded == single-GPU: {same}; max diff {float((own - ref[:, lo:hi]).abs().max()):.3e}",
          flush=True)
    ok = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) == 1 else 1)


if __name__ == "__main__":
    main()
