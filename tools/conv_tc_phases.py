"""Phase timestamps of the halo conv kernel (i2v_debug_conv_tc_timestamps) for the narrow BAIR g_4 layers,
plain halo form (variant 2) against the kw-stacked form (variant 3).  Run on a B200: python tools/conv_tc_phases.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ops_util as ou  # noqa: E402
from image2video_synthesis_using_cinns_b200 import lib  # noqa: E402

L = lib.load()
NAMES = ["start", "prologue", "first_stage", "last_mma_issued", "acc_complete", "epi_stores", "end"]


def run(name, B, C, T, H, W, Cout, variant, out_mode=0):
    x = torch.randn(B, T, H, W, C, device="cuda")
    w = torch.randn(27, Cout, C, device="cuda") * 0.02
    b = torch.zeros(Cout, device="cuda")
    ncta = 4096
    buf = torch.zeros(ncta * 8, dtype=torch.int64, device="cuda")
    for _ in range(2):
        lib.check(L.i2v_debug_conv_tc_timestamps(ctypes.c_void_p(buf.data_ptr()), ncta), "dbg")
        ou.conv_tc(x, w, b, None, (3, 3, 3), variant=variant, out_mode=out_mode)
        torch.cuda.synchronize()
    lib.check(L.i2v_debug_conv_tc_timestamps(None, 0), "dbg")
    t = buf.cpu().view(ncta, 8).double()
    t = t[t[:, 6] > 0]
    d = (t - t[:, 0:1]) / 1000.0
    span = float((t[:, 6].max() - t[:, 0].min()) / 1000)
    print(f"{name} variant={variant}: {len(t)} CTAs, kernel span {span:.1f} us, "
          f"{span / (len(t) / 148.0):.2f} us per CTA slot")
    print("   " + "  ".join(f"{n}={float(d[:, i].median()):.2f}" for i, n in enumerate(NAMES)))


if __name__ == "__main__":
    for v in (2, 3):
        run("g4_conv1 64->64", 8, 64, 16, 64, 64, 64, v)
        run("g4_conv0 128->64", 8, 128, 16, 64, 64, 64, v)
        run("conv_img 64->3", 8, 64, 16, 64, 64, 3, v, out_mode=1)
