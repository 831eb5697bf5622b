"""torchrun --nproc-per-node N tools/check_gather.py [--stress] : FrameGather (p2p / nccl) returns the rank-ordered
concatenation.  --stress: bench-like use -- large frames, no host synchronisation between submits, heavy kernels on the compute
stream, NCCL barriers in between."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from image2video_synthesis_using_cinns_b200.dist import FrameGather

stress = "--stress" in sys.argv
modes = [m for m in ("p2p", "nccl") if f"--{m}" in sys.argv] or ["p2p", "nccl"]
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
shape = (64, 16, 3, 64, 64) if stress else (3, 2, 3, 8, 8)
ok = True
a = torch.randn(4096, 4096, device=dev)
for mode in modes:
    g = FrameGather(dev, mode=mode)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    outs = []
    n_it = 12 if stress else 5
    for it in range(n_it):
        if stress:
            for _ in range(20):
                a = (a @ a).clamp_(-1, 1)                     # ~50 ms of compute-stream work per step
        x = torch.full(shape, float(rank * 100 + it), device=dev)
        g.submit(x)
        if not stress or it % 4 == 3:
            out = g.wait().clone()
            want = torch.cat([torch.full(shape, float(r * 100 + it), device=dev) for r in range(world)])
            ok = ok and torch.equal(out, want)
        if stress and it == 5:
            dist.barrier()
    g.wait()
    torch.cuda.synchronize()
    print(f"rank {rank} mode requested {mode} -> used {g.mode} {'ok' if ok else 'MISMATCH'} {time.time() - t0:.2f}s "
          f"{getattr(g, 'fallback_reason', '')}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
