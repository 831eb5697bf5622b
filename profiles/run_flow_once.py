"""Runs the full-size BAIR flow (20 blocks, hidden 512) once per direction at B=64 -- target for ncu captures."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from image2video_synthesis_using_cinns_b200 import synthetic
from image2video_synthesis_using_cinns_b200.modules import ConditionalFlow
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
if "--coop" in sys.argv:
    from image2video_synthesis_using_cinns_b200 import lib as _l
    _l.set_option("flow_cluster", 0)
gen = torch.Generator().manual_seed(0)
sd = synthetic.flow_state_dict(gen, 64, 64, 512, 2, 20)
flow = ConditionalFlow(sd, 64, 64, 512, 2, 20)
x = torch.randn(B, 64, generator=gen).cuda(); c = torch.randn(B, 64, generator=gen).cuda()
for _ in range(3):
    z = flow(x, c, reverse=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); z = flow(x, c, reverse=True); e1.record(); torch.cuda.synchronize()
print("flow reverse B=%d: %.3f ms" % (B, e0.elapsed_time(e1)))

import ctypes
from image2video_synthesis_using_cinns_b200 import lib
L = lib.load()
buf = torch.zeros(32, dtype=torch.int64, device="cuda")
lib.check(L.i2v_debug_flow_timestamps(ctypes.c_void_p(buf.data_ptr())))
z = flow(x, c, reverse=True); torch.cuda.synchronize()
lib.check(L.i2v_debug_flow_timestamps(None))
t = buf.cpu().view(2, 16).double()
cluster = "--coop" not in sys.argv
if cluster:   # cluster-resident kernel (flow_cluster.cu): coupling #4, CTA ranks 0 and 9 of cluster 0
    names = ["start", "L1 weights landed", "L1 pushed", "L1 handed over", "H1 MACs", "H1 partials in smem", "H1 pushed", "H1 handed over",
             "H2 MACs", "H2 partials in smem", "H2 pushed", "H2 handed over", "last pushed", "last handed over", "update done"]
    who = ("rank 0 (scale net)", "rank 9 (translation net)")
else:
    names = ["start", "L1 done", "bar1 passed", "H1 done", "bar2 passed", "H2 done", "bar3 passed", "last done", "bar4 passed", "update done"]
    who = ("CTA 0 (has last-layer rows)", "CTA 100")
for row, w in zip(t, who):
    print(w, " ".join(f"{n}={(row[i] - row[0]) / 1000:.2f}us" for i, n in enumerate(names)))
