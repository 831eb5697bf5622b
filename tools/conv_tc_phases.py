"""Phase timestamps of the halo conv kernel (i2v_debug_conv_tc_timestamps) on the decoder's big layers.

    python tools/conv_tc_phases.py [narrow|wide|epi|all] [variant ...]

`narrow`: BAIR g_4 / conv_img, plain halo form (variant 2) against the kw-stacked form (variant 3).
`wide`  : g_2 / g_3 layers (N = 128), automatic variant.
`epi`   : epilogue break-down with / without a residual (shortcut through the upsample map, in-place K-split partial).
Run once per setting of the tuning knobs (I2V_TC_MIN_STAGES / I2V_TC_FLAGS are read once per process).  B200 only."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ops_util as ou  # noqa: E402
from image2video_synthesis_using_cinns_b200 import lib  # noqa: E402

L = lib.load()
NAMES = ["producer_start", "prologue", "first_stage", "last_mma_issued", "acc_complete", "epi_done", "end", "-",
         "sub0_out_of_tmem", "sub0_stored", "sub1_out_of_tmem", "sub1_stored", "stats_flushed"]
SLOTS = 16


def run(name, B, C, T, H, W, Cout, variant, out_mode=0, res_up=None):
    x = torch.randn(B, T, H, W, C, device="cuda")
    w = torch.randn(27, Cout, C, device="cuda") * 0.02
    b = torch.zeros(Cout, device="cuda")
    res = None
    if res_up is not None:
        res = torch.randn(B, T // res_up[0], H // res_up[1], W // res_up[2], Cout, device="cuda")
    ncta = 8192
    buf = torch.zeros(ncta * SLOTS, dtype=torch.int64, device="cuda")
    for _ in range(2):
        lib.check(L.i2v_debug_conv_tc_timestamps(ctypes.c_void_p(buf.data_ptr()), ncta), "dbg")
        ou.conv_tc(x, w, b, res, (3, 3, 3), res_up=res_up or (1, 1, 1), variant=variant, out_mode=out_mode)
        torch.cuda.synchronize()
    lib.check(L.i2v_debug_conv_tc_timestamps(None, 0), "dbg")
    t = buf.cpu().view(ncta, SLOTS).double()
    t = t[t[:, 6] > 0]
    # persistent tile loop: "start" of a later tile is the moment the producer turns to it (under the previous tile's
    # main loop), so phases are reported relative to the first landed stage of the tile
    d = (t - t[:, 2:3]) / 1000.0
    span = float((t[:, 6].max() - t[:, 0].min()) / 1000)
    flops = 2.0 * 27 * C * Cout * B * T * H * W
    knobs = f"min_stages={os.environ.get('I2V_TC_MIN_STAGES', '-')} flags={os.environ.get('I2V_TC_FLAGS', '-')}"
    print(f"{name} variant={variant} res_up={res_up} {knobs}: {len(t)} tiles (last launch of the K split), span {span:.1f} us, "
          f"{span / (len(t) / 148.0):.2f} us per tile and SM, {flops / span * 1e-6:.0f} alg. TFLOP/s if 1 launch")
    print("   " + "  ".join(f"{n}={float(d[:, i].median()):.2f}" for i, n in enumerate(NAMES) if n != "-" and float(t[:, i].max()) > 0))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "narrow"
    if what in ("narrow", "all"):
        for v in ([int(a) for a in sys.argv[2:]] or [2, 3]):
            run("g4_conv1 64->64", 8, 64, 16, 64, 64, 64, v)
            run("g4_conv0 128->64", 8, 128, 16, 64, 64, 64, v)
            run("conv_img 64->3", 8, 64, 16, 64, 64, 3, v, out_mode=1)
    if what in ("wide", "all"):
        run("g3_conv1 128->128", 8, 128, 16, 64, 64, 128, 0)
        run("g3_conv0 256->128 (no phase form)", 8, 256, 16, 64, 64, 128, 0)
        run("g2_conv1 256->256", 16, 256, 8, 32, 32, 256, 0)
        run("g4_128 conv1 32->32", 2, 32, 16, 128, 128, 32, 0)
        run("g3_128 conv1 64->64", 8, 64, 16, 64, 64, 64, 0)
    if what in ("epi", "all"):
        run("g3_conv1 128->128 plain", 8, 128, 16, 64, 64, 128, 0)
        run("g3_conv1 128->128 + shortcut x2x2x2", 8, 128, 16, 64, 64, 128, 0, res_up=(2, 2, 2))
        run("g3_conv0 256->128 2 K parts (in-place partial)", 8, 256, 16, 64, 64, 128, 0)
        run("g4_conv1 64->64 stacked plain", 8, 64, 16, 64, 64, 64, 0)
        run("g4_conv1 64->64 stacked + full-res shortcut", 8, 64, 16, 64, 64, 64, 0, res_up=(1, 1, 1))
