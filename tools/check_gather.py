"""torchrun --nproc-per-node N tools/check_gather.py : FrameGather (p2p / nccl) returns the rank-ordered concatenation."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image2video_synthesis_using_cinns_b200.dist import FrameGather

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
ok = True
for mode in ("p2p", "nccl"):
    g = FrameGather(dev, mode=mode)
    for it in range(5):
        x = torch.full((3, 2, 3, 8, 8), float(rank * 100 + it), device=dev) + torch.arange(8, device=dev)
        g.submit(x)
        torch.zeros(1 << 22, device=dev).normal_()            # later work on the compute stream
        out = g.wait().clone()
        want = torch.cat([torch.full((3, 2, 3, 8, 8), float(r * 100 + it), device=dev) + torch.arange(8, device=dev) for r in range(world)])
        good = torch.equal(out, want)
        ok = ok and good
    print(f"rank {rank} mode requested {mode} -> used {g.mode} {'ok' if ok else 'MISMATCH'} {getattr(g, 'fallback_reason', '')}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
