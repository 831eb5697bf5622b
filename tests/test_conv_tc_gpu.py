"""Tensor-core conv engine (tcgen05 + TMA, error-compensated fp16 split) against fp32 PyTorch."""
import pytest
import torch
import torch.nn.functional as F

import ops_util as ou
from golden_util import rel_inf, report

pytestmark = pytest.mark.gpu
G = lambda s: torch.Generator().manual_seed(s)

CASES = [
    # name, (B,Cin,T,H,W), Cout, k
    ("g3_like_w64", (1, 64, 2, 8, 64), 64, (3, 3, 3)),        # box 64x2x1x1, kc=64
    ("w32_kc64_n128", (2, 128, 2, 32, 32), 128, (3, 3, 3)),   # box 32x4
    ("w16_n256", (1, 64, 4, 16, 16), 256, (3, 3, 3)),         # box 16x8, N=256
    ("n512_two_ntiles", (1, 64, 1, 8, 8), 512, (3, 3, 3)),    # 2 N tiles, box 8x8x1x2 (bb=2 > B)
    ("head_4x4_batchbox", (3, 64, 1, 4, 4), 64, (3, 3, 3)),   # box 4x4x1x8 over the batch dim
    ("g0_t2", (2, 64, 2, 8, 8), 64, (3, 3, 3)),               # two-frame clip: one-frame tiles of two samples (box 8x8x1x2)
    ("g0_t2_b3", (3, 64, 2, 8, 8), 128, (3, 3, 3)),           # ... with a half-empty last tile
    ("kc32", (1, 32, 2, 16, 16), 32, (3, 3, 3)),              # SWIZZLE_64B rows
    ("kc16", (1, 16, 2, 16, 16), 16, (3, 3, 3)),              # SWIZZLE_32B rows
    ("conv_s_1x1x1", (2, 128, 2, 8, 8), 64, (1, 1, 1)),
    ("spade_gb_2d", (2, 128, 1, 16, 16), 256, (1, 3, 3)),
    ("w128_row", (1, 32, 2, 4, 128), 32, (3, 3, 3)),          # box 128x1
    ("long_k", (1, 256, 2, 8, 8), 128, (3, 3, 3)),            # 108 pipeline iterations, ring wraps many times
    ("long_k_halo", (1, 256, 2, 16, 16), 128, (3, 3, 3)),     # halo kernel, 2 accumulators x 2 sub-tiles
    ("halo_n64_t4", (2, 64, 4, 8, 64), 64, (3, 3, 3)),        # halo kernel, interior t (all 3 temporal taps) + edges
    ("halo_cout3", (1, 64, 2, 4, 64), 3, (3, 3, 3)),          # conv_img shape: N padded to 16
    ("g4_conv0_like", (1, 128, 4, 8, 64), 64, (3, 3, 3)),     # kw-stacked N=192, kc=16, one shared accumulator
    ("wstack_w32_n32", (2, 32, 2, 32, 32), 32, (3, 3, 3)),    # stacked N=96, 2 accumulators, rows of 32 (no cross-warp edge)
    ("wstack_w16_n48", (1, 64, 2, 16, 16), 48, (3, 3, 3)),    # stacked N=144, rows of 16
    ("wstack_2d", (2, 64, 1, 16, 64), 64, (1, 3, 3)),         # 2-D conv (kt=1), stacked
]
HALO_OK = {"g3_like_w64", "w32_kc64_n128", "w16_n256", "kc32", "spade_gb_2d", "w128_row", "long_k_halo",
           "halo_n64_t4", "halo_cout3", "g4_conv0_like", "wstack_w32_n32", "wstack_w16_n48", "wstack_2d"}
# CTA-pair kernel (cta_group::2, two TMEM accumulator sets): variant=5 = plain form (N tiles of 64 / 128 columns),
# variant=4 = kw-stacked form on narrow layers (N = 3 Cout), plain form otherwise
PAIR5_OK = {"g3_like_w64", "w32_kc64_n128", "w16_n256", "spade_gb_2d", "long_k_halo", "halo_n64_t4", "g4_conv0_like", "wstack_2d"}
PAIR_OK = PAIR5_OK | {"kc32", "w128_row", "wstack_w32_n32", "wstack_w16_n48"}
# narrow layers (Cout <= 64, whole w-rows per tile): the halo kernel's kw-stacked form, forced with variant=3
WSTACK_OK = {"g3_like_w64", "kc32", "w128_row", "halo_n64_t4", "halo_cout3", "g4_conv0_like", "wstack_w32_n32",
             "wstack_w16_n48", "wstack_2d"}


@pytest.mark.parametrize("name,xs,cout,k", CASES, ids=[c[0] for c in CASES])
def test_conv_tc_matches_fp32(name, xs, cout, k):
    g = G(sum(map(ord, name)))
    x = torch.randn(xs, generator=g)
    w = torch.randn(cout, xs[1], *k, generator=g) / (xs[1] * k[0] * k[1] * k[2]) ** 0.5
    b = torch.randn(cout, generator=g)
    pad = tuple(kk // 2 for kk in k)
    want = F.conv3d(x.double(), w.double(), b.double(), 1, pad)
    got = ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, k))
    e3 = rel_inf(got, want)
    simt = rel_inf(ou.from_cl(ou.conv(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, (1, 1, 1), pad)), want)
    e1 = rel_inf(ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, terms=1)), want)
    # both kernels explicitly: v1 (per-tap boxes) everywhere, v2 (256-row H-halo) where the shape allows
    ev1 = rel_inf(ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, variant=1)), want)
    ev2 = rel_inf(ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, variant=2)), want) if name in HALO_OK else None
    ev3 = rel_inf(ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, variant=3)), want) if name in WSTACK_OK else None
    ev4 = rel_inf(ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, variant=4)), want) if name in PAIR_OK else None
    ev4f = rel_inf(ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, variant=4, terms=1)), want) if name in PAIR_OK else None
    ev5 = rel_inf(ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, variant=5)), want) if name in PAIR5_OK else None
    report("conv_tc:" + name, split3=e3, fp16_single=e1, simt_fp32=simt, v1=ev1, v2_halo=ev2, v3_wstack=ev3, v4_pair=ev4, v4_pair_fp16=ev4f,
           v5_pair_plain=ev5)
    assert ev4 is None or (ev4 < 1e-5 and ev4f < 1e-3)
    assert ev5 is None or ev5 < 1e-5
    assert ev1 < 6e-6 and (ev2 is None or ev2 < 1e-5)   # halo kernel: 2 accumulators at N=128 instead of 4
    assert ev3 is None or ev3 < 1e-5
    assert got.shape == want.shape
    # fp32-grade: within a small factor of the fp32 SIMT engine's own rounding.  The residual gap is the
    # tensor core's truncating fp32 accumulate (measured ~1e-5 at K=6912 with ONE accumulator), which the
    # 4-way TMEM accumulator round-robin cuts down; see DESIGN.md section 5.
    assert e3 < 1e-5
    assert e1 < 1e-3           # single fp16 product (fast mode)


def test_conv_tc_epilogue_residual_act_and_frames_layout():
    g = G(7)
    x = torch.randn(2, 64, 4, 8, 8, generator=g)
    w = torch.randn(32, 64, 3, 3, 3, generator=g) * 0.03
    b = torch.randn(32, generator=g)
    res = torch.randn(2, 32, 2, 4, 4, generator=g)
    want = F.leaky_relu(F.conv3d(x, w, b, 1, 1) + F.interpolate(res, scale_factor=2.0), 0.2)
    got = ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), ou.to_cl(res), (3, 3, 3), res_up=(2, 2, 2), act=2)
    assert rel_inf(ou.from_cl(got), want) < 1e-5
    w3 = torch.randn(3, 64, 3, 3, 3, generator=g) * 0.03
    b3 = torch.randn(3, generator=g)
    want = torch.tanh(F.conv3d(x, w3, b3, 1, 1)).transpose(1, 2)
    got = ou.conv_tc(ou.to_cl(x), ou.taps(w3), b3.cuda(), None, (3, 3, 3), act=3, out_mode=1)
    assert rel_inf(got.cpu(), want) < 1e-5


@pytest.mark.parametrize("shape", [(2, 64, 4, 16, 32, 128), (1, 32, 2, 32, 64, 64), (3, 64, 2, 16, 16, 192), (1, 64, 2, 4, 128, 64)])
def test_conv_tc_pair_epilogue_residual_act_many_tiles(shape):
    """CTA-pair kernel: bias, residual through the x2x2x2 upsample map, activation, a partial last N tile (Cout = 192),
    and more tiles than CTA pairs would hold at once for the accumulator-set alternation (odd and even tile counts)."""
    B, C, T, H, W, cout = shape
    g = G(B * 1000 + W + cout)
    x = torch.randn(B, C, T, H, W, generator=g)
    w = torch.randn(cout, C, 3, 3, 3, generator=g) * 0.03
    b = torch.randn(cout, generator=g)
    res = torch.randn(B, cout, T // 2, H // 2, W // 2, generator=g)
    want = F.leaky_relu(F.conv3d(x.double(), w.double(), b.double(), 1, 1) + F.interpolate(res.double(), scale_factor=2.0), 0.2)
    got = ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), ou.to_cl(res), (3, 3, 3), res_up=(2, 2, 2), act=2, variant=4)
    assert rel_inf(ou.from_cl(got), want) < 1e-5
    # same launch twice: the barriers / accumulator sets leave no state behind
    got2 = ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), ou.to_cl(res), (3, 3, 3), res_up=(2, 2, 2), act=2, variant=4)
    assert torch.equal(got, got2)


def test_conv_tc_pair_many_tiles_per_pair():
    """A batch large enough that every CTA pair walks > 4 tiles: ring and accumulator-set phases wrap several times."""
    g = G(77)
    x = torch.randn(8, 32, 8, 32, 32, generator=g)          # 8*8*4 = 256 tiles of 256 voxels... x 1 N tile over 74 pairs
    w = torch.randn(64, 32, 3, 3, 3, generator=g) * 0.05
    want = F.conv3d(x, w, None, 1, 1)
    got = ou.conv_tc(ou.to_cl(x), ou.taps(w), None, None, (3, 3, 3), variant=4)
    assert rel_inf(ou.from_cl(got), want) < 1e-5


@pytest.mark.parametrize("variant", [3, 4], ids=["halo", "pair"])
@pytest.mark.parametrize("hw", [(16, 16), (8, 64), (2, 128)])
def test_conv_tc_wstack_epilogue_residual_act_and_frames_layout(hw, variant):
    """kw-stacked form (single-CTA halo kernel / CTA-pair kernel): shifted-sum epilogue with bias, upsampled residual,
    activation; frame layout + tanh (conv_img)."""
    g = G(11)
    H, W = hw
    x = torch.randn(2, 64, 2, H, W, generator=g)
    w = torch.randn(32, 64, 3, 3, 3, generator=g) * 0.03
    b = torch.randn(32, generator=g)
    res = torch.randn(2, 32, 1, H // 2, W // 2, generator=g)
    want = F.leaky_relu(F.conv3d(x, w, b, 1, 1) + F.interpolate(res, scale_factor=2.0), 0.2)
    got = ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), ou.to_cl(res), (3, 3, 3), res_up=(2, 2, 2), act=2, variant=variant)
    assert rel_inf(ou.from_cl(got), want) < 1e-5
    w3 = torch.randn(3, 64, 3, 3, 3, generator=g) * 0.03
    b3 = torch.randn(3, generator=g)
    want = torch.tanh(F.conv3d(x, w3, b3, 1, 1)).transpose(1, 2)
    got = ou.conv_tc(ou.to_cl(x), ou.taps(w3), b3.cuda(), None, (3, 3, 3), act=3, out_mode=1, variant=variant)
    assert rel_inf(got.cpu(), want) < 1e-5
    # single fp16 product mode takes the same path with one accumulator
    got = ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, (3, 3, 3), terms=1, variant=variant)
    assert rel_inf(ou.from_cl(got), F.conv3d(x, w, b, 1, 1)) < 1e-3


@pytest.mark.parametrize("engine", [1])
def test_decoder_tensor_core_engine_matches_oracle(engine, ckpt_cache):
    import oracle_torch as ot
    from image2video_synthesis_using_cinns_b200.get_model import Model
    for dataset, nf, B in (("bair", 32, 3), ("dtdb_fire", 16, 2)):
        mp = ckpt_cache(dataset=dataset, seed=31, nf=nf, n_flows=2, spade_gain=1.0, with_encoder=False)
        m = Model(mp, 16, conv_engine=engine, micro_batch=2)
        om = ot.OracleModel(mp, 16)
        img = m.config.Data["img_size"]
        g = G(3)
        x0 = torch.rand(B, 3, img, img, generator=g) * 2 - 1
        z = torch.randn(B, 64, generator=g)
        want = om.decode(x0, z)
        e = rel_inf(m.decoder(x0.cuda(), z.cuda()).cpu(), want)
        report(f"decoder_tc{engine}:{dataset}", decoder=e)
        assert e < 1e-4


SIDE_CASES = [
    # name, (B, Cmid, T, H, W), Cin2, Cout, variant
    ("bair_g4_stacked", (1, 64, 4, 8, 64), 128, 64, 0),      # BAIR g_4.conv_1: kw-stacked N = 192, kc = 64, 2 side chunks
    ("plain_halo_n128", (1, 128, 2, 16, 32), 256, 128, 0),   # N = 128: plain halo form, weight slab kw = 1 of the stack
    ("small_nf32", (2, 32, 4, 16, 16), 64, 32, 0),           # narrow net: kc = 32 (64-byte rows), stacked N = 96
    ("stacked_forced_off", (1, 64, 2, 8, 64), 128, 64, 2),   # same layer through the plain halo form
    ("single_product", (1, 64, 2, 8, 64), 128, 64, 0),       # terms = 1 (conv_engine 2)
    ("pair_bair_g4", (1, 64, 4, 8, 64), 128, 64, 4),         # CTA-pair kernel: side chunks through the centre row
    ("pair_n128", (2, 32, 2, 16, 32), 64, 128, 4),           # (a side input excludes K-split launches: short reduction)
]


@pytest.mark.parametrize("name,xs,cin2,cout,variant", SIDE_CASES, ids=[c[0] for c in SIDE_CASES])
def test_conv_tc_side_input_is_fused_shortcut(name, xs, cin2, cout, variant):
    """y = conv3x3x3(x) + conv1x1x1(x2) + b in one launch (GeneratorBlock shortcut fused into conv_1, decoder.py:44-50)."""
    g = G(sum(map(ord, name)))
    B, C, T, H, W = xs
    x = torch.randn(xs, generator=g)
    x2 = torch.randn(B, cin2, T, H, W, generator=g)
    w = torch.randn(cout, C, 3, 3, 3, generator=g) / (27 * C) ** 0.5
    w2 = torch.randn(cout, cin2, generator=g) / cin2 ** 0.5
    b = torch.randn(cout, generator=g)
    want = F.conv3d(x.double(), w.double(), b.double(), 1, 1) + F.conv3d(x2.double(), w2.double()[:, :, None, None, None])
    terms = 1 if name == "single_product" else 3
    got = ou.from_cl(ou.conv_tc_side(ou.to_cl(x), ou.taps(w), ou.to_cl(x2), w2.cuda(), b.cuda(), terms=terms, variant=variant))
    e = rel_inf(got, want)
    report("conv_tc_side:" + name, err=e)
    assert e < (1e-3 if terms == 1 else 1e-5)


PHASE_CASES = [
    # name, (B, Cin, T/2, H, W), Cout, variant
    ("g3_conv0_like", (1, 128, 4, 16, 32), 64, 0),      # N = 64 at W = 32: kw-stacked halo form
    ("g2_conv0_like", (2, 64, 2, 16, 16), 128, 0),      # N = 128: plain halo form
    ("t_half_1", (2, 32, 1, 16, 16), 32, 0),            # T = 2: both temporal edges in one tile pair
    ("plain_forced", (1, 64, 2, 8, 64), 64, 2),         # stacked form switched off
    ("pair_g3_conv0", (1, 128, 4, 16, 32), 64, 4),      # CTA-pair kernel
    ("pair_n128", (2, 64, 2, 16, 16), 128, 4),
    ("pair_t_half_1", (2, 64, 1, 8, 64), 64, 4),
    # per-tap kernel (planes below 16x16, i.e. g_0.conv_0 at 2 x 8 x 8: one launch dimension per output phase)
    ("pertap_g0_like", (2, 64, 1, 8, 8), 128, 0),       # one source plane: a single temporal tap per phase survives
    ("pertap_g0_b3", (3, 64, 1, 8, 8), 64, 0),          # tiles of two samples, the last one half empty
    ("pertap_tsrc2", (2, 32, 2, 8, 8), 32, 1),          # two source planes in one tile: both taps live, output frames interleave
    ("pertap_4x4", (2, 32, 4, 4, 4), 32, 1),            # 16-voxel planes: warp slices straddle output frames (generic write-out)
    ("pertap_forced_16x16", (1, 64, 2, 16, 16), 64, 1), # per-tap kernel forced on a halo-eligible shape
]


@pytest.mark.parametrize("name,xs,cout,variant", PHASE_CASES, ids=[c[0] for c in PHASE_CASES])
def test_conv_tc_temporal_phase_form(name, xs, cout, variant):
    """conv_tc(t_phase=1) on the T/2 tensor == F.conv3d on the repeat_interleave'd one (decoder.py:102-111; the phase
    weights are loader.phase_weights, the form csrc/conv_tc.cu::halo_tile documents)."""
    g = G(sum(map(ord, name)))
    x = torch.randn(xs, generator=g)
    w = torch.randn(cout, xs[1], 3, 3, 3, generator=g) / (27 * xs[1]) ** 0.5
    b = torch.randn(cout, generator=g)
    want = F.conv3d(x.double().repeat_interleave(2, dim=2), w.double(), b.double(), 1, 1)
    got = ou.from_cl(ou.conv_tc_phase(ou.to_cl(x), ou.taps(w), b.cuda(), variant=variant))
    e = rel_inf(got, want)
    e1 = rel_inf(ou.from_cl(ou.conv_tc_phase(ou.to_cl(x), ou.taps(w), b.cuda(), terms=1, variant=variant)), want)
    report("conv_tc_phase:" + name, split3=e, fp16_single=e1)
    assert got.shape == want.shape
    assert e < 1e-5 and e1 < 1e-3


def test_conv_tc_two_frame_tiling_switch():
    """tc_t2_split: a two-frame clip runs as one-frame tiles that skip the padded temporal tap; both tilings give the conv."""
    from image2video_synthesis_using_cinns_b200 import lib
    g = G(77)
    x = torch.randn(4, 64, 2, 8, 8, generator=g)
    w = torch.randn(64, 64, 3, 3, 3, generator=g) / (27 * 64) ** 0.5
    b = torch.randn(64, generator=g)
    want = F.conv3d(x.double(), w.double(), b.double(), 1, 1)
    errs = {}
    try:
        for sw in (0, 1):
            lib.set_option("tc_t2_split", sw)
            errs[sw] = rel_inf(ou.from_cl(ou.conv_tc(ou.to_cl(x), ou.taps(w), b.cuda(), None, (3, 3, 3))), want)
    finally:
        lib.set_option("tc_t2_split", 1)
    report("conv_tc:t2_split", whole_clip_tiles=errs[0], one_frame_tiles=errs[1])
    assert errs[0] < 1e-5 and errs[1] < 1e-5


def test_split_saturates_instead_of_overflowing():
    """|x| beyond the fp16 range of the operand split (16 * x > 65504) must clip, not turn into inf / NaN."""
    g = G(5)
    x = torch.randn(1, 64, 2, 16, 16, generator=g)
    x[0, 3, 1, 5, 7] = 1.0e5          # 16 * 1e5 overflows fp16
    x[0, 9, 0, 2, 2] = -3.0e6
    w = torch.randn(32, 64, 3, 3, 3, generator=g) * 0.03
    got = ou.conv_tc(ou.to_cl(x), ou.taps(w), None, None, (3, 3, 3))
    assert torch.isfinite(got).all()
    hi, lo = ou.modulate_split(ou.to_cl(x), None, (2, 16, 16), act=2)
    assert torch.isfinite(hi.float()).all() and torch.isfinite(lo.float()).all()
    assert hi.float().abs().max() <= 65504
    # in-range values are untouched by the clamp
    xs = torch.randn(1, 64, 2, 16, 16, generator=g)
    a = ou.conv_tc(ou.to_cl(xs), ou.taps(w), None, None, (3, 3, 3))
    assert rel_inf(ou.from_cl(a), F.conv3d(xs.double(), w.double(), None, 1, 1)) < 1e-5
