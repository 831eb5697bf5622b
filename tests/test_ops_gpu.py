"""Kernel-level parity: each C-ABI op against the same op in fp32 PyTorch on the CPU."""
import pytest
import torch
import torch.nn.functional as F

import ops_util as ou
from golden_util import rel_inf

pytestmark = pytest.mark.gpu
G = lambda s: torch.Generator().manual_seed(s)
ACTS = {0: lambda v: v, 1: F.relu, 2: lambda v: F.leaky_relu(v, 0.2), 3: torch.tanh}

CONV_CASES = [
    # name, x shape (B,C,T,H,W), Cout, k, stride, pad
    ("3x3x3_same", (2, 32, 4, 8, 8), 48, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("3x3x3_wide", (1, 64, 2, 16, 16), 128, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("1x1x1", (2, 64, 2, 4, 4), 32, (1, 1, 1), (1, 1, 1), (0, 0, 0)),
    ("3x3x3_strided", (2, 16, 8, 16, 16), 32, (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    ("3x3x3_stride_s_only", (1, 16, 4, 16, 16), 16, (3, 3, 3), (1, 2, 2), (1, 1, 1)),
    ("enc_conv1_3x7x7", (1, 3, 15, 32, 32), 64, (3, 7, 7), (2, 2, 2), (1, 3, 3)),
    ("t1_head", (3, 32, 1, 4, 4), 32, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("cout3_img", (1, 16, 4, 16, 16), 3, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("ragged_m", (1, 16, 3, 5, 7), 20, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
]


@pytest.mark.parametrize("name,xs,cout,k,stride,pad", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv3d(name, xs, cout, k, stride, pad):
    g = G(sum(map(ord, name)))
    x = torch.randn(xs, generator=g)
    w = torch.randn(cout, xs[1], *k, generator=g) / (xs[1] * k[0] * k[1] * k[2]) ** 0.5
    b = torch.randn(cout, generator=g)
    want = F.conv3d(x, w, b, stride, pad)
    got = ou.from_cl(ou.conv(ou.to_cl(x), ou.taps(w), b.cuda(), None, k, stride, pad))
    assert got.shape == want.shape
    assert rel_inf(got, want) < 1e-5


def test_conv2d_resnet_stem_and_spade():
    g = G(5)
    x = torch.randn(2, 3, 64, 64, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    want = F.conv2d(x, w, None, 2, 3)
    got = ou.conv(ou.to_cl(x)[:, None], ou.taps(w), None, None, (1, 7, 7), (1, 2, 2), (0, 3, 3))[:, 0]
    assert rel_inf(ou.from_cl(got), want) < 1e-5
    w2 = torch.randn(128, 3, 3, 3, generator=g) * 0.2
    b2 = torch.randn(128, generator=g)
    want = F.leaky_relu(F.conv2d(x, w2, b2, 1, 1), 0.2)
    got = ou.conv(ou.to_cl(x)[:, None], ou.taps(w2), b2.cuda(), None, (1, 3, 3), (1, 1, 1), (0, 1, 1), act=2)[:, 0]
    assert rel_inf(ou.from_cl(got), want) < 1e-5
    # strided 1x1 (resnet downsample) and strided 3x3
    x3 = torch.randn(2, 64, 16, 16, generator=g)
    w3 = torch.randn(128, 64, 1, 1, generator=g) * 0.1
    got = ou.conv(ou.to_cl(x3)[:, None], ou.taps(w3), None, None, (1, 1, 1), (1, 2, 2), (0, 0, 0))[:, 0]
    assert rel_inf(ou.from_cl(got), F.conv2d(x3, w3, None, 2)) < 1e-5
    w4 = torch.randn(64, 64, 3, 3, generator=g) * 0.05
    got = ou.conv(ou.to_cl(x3)[:, None], ou.taps(w4), None, None, (1, 3, 3), (1, 2, 2), (0, 1, 1))[:, 0]
    assert rel_inf(ou.from_cl(got), F.conv2d(x3, w4, None, 2, 1)) < 1e-5


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_conv_epilogue_residual_upsample_and_act(act):
    g = G(11 + act)
    x = torch.randn(2, 32, 4, 8, 8, generator=g)
    w = torch.randn(16, 32, 3, 3, 3, generator=g) * 0.05
    b = torch.randn(16, generator=g)
    res = torch.randn(2, 16, 2, 4, 4, generator=g)
    want = ACTS[act](F.conv3d(x, w, b, 1, 1) + F.interpolate(res, scale_factor=2.0))
    got = ou.conv(ou.to_cl(x), ou.taps(w), b.cuda(), ou.to_cl(res), (3, 3, 3), (1, 1, 1), (1, 1, 1), res_up=(2, 2, 2), act=act)
    assert rel_inf(ou.from_cl(got), want) < 1e-5


def test_conv_frames_layout():
    g = G(21)
    x = torch.randn(2, 16, 4, 8, 8, generator=g)
    w = torch.randn(3, 16, 3, 3, 3, generator=g) * 0.1
    b = torch.randn(3, generator=g)
    want = torch.tanh(F.conv3d(x, w, b, 1, 1)).transpose(1, 2)   # decoder.py:117-120
    got = ou.conv(ou.to_cl(x), ou.taps(w), b.cuda(), None, (3, 3, 3), (1, 1, 1), (1, 1, 1), act=3, out_mode=1)
    assert rel_inf(got.cpu(), want) < 1e-5


@pytest.mark.parametrize("C,shape", [(64, (2, 4, 8, 8)), (256, (1, 1, 4, 4)), (16, (3, 16, 32, 32)), (2048, (2, 1, 2, 2))])
def test_stats_groupnorm_instancenorm(C, shape):
    g = G(C)
    B, T, H, W = shape
    x = torch.randn(B, C, T, H, W, generator=g) * 2 + 0.7
    xc = ou.to_cl(x)
    sums, V = ou.channel_stats(xc)
    assert rel_inf(sums[..., 0].cpu(), x.double().sum(dim=(2, 3, 4))) < 1e-6
    # GroupNorm(16, affine) -> ReLU
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    coef = ou.norm_coeffs(sums, V, 16, gamma.cuda(), beta.cuda())
    got = ou.modulate(xc, coef, (T, H, W), act=1)
    assert rel_inf(ou.from_cl(got), F.relu(F.group_norm(x, 16, gamma, beta, 1e-5))) < 2e-5
    # InstanceNorm + AdaIN modulation + lrelu(0.2)
    mod = torch.randn(B, 2 * C, generator=g)
    coef = ou.norm_coeffs(sums, V, 0, mod=mod.cuda())
    got = ou.modulate(xc, coef, (T, H, W), act=2)
    want = F.leaky_relu(mod[:, :C].reshape(B, C, 1, 1, 1) * F.instance_norm(x, eps=1e-5) + mod[:, C:].reshape(B, C, 1, 1, 1), 0.2)
    assert rel_inf(ou.from_cl(got), want) < 2e-5


def test_modulate_spade_upsample_and_two_branch():
    g = G(33)
    B, C = 2, 32
    x = torch.randn(B, C, 2, 4, 4, generator=g) + 0.3
    gamma = torch.randn(B, C, 8, 8, generator=g) * 0.5
    beta = torch.randn(B, C, 8, 8, generator=g) * 0.5
    xu = F.interpolate(x, scale_factor=2.0)
    want = F.leaky_relu(F.group_norm(xu, 16, eps=1e-5) * (1 + gamma.unsqueeze(2)) + beta.unsqueeze(2), 0.2)
    sums, V = ou.channel_stats(ou.to_cl(x))
    coef = ou.norm_coeffs(sums, V, 16)
    gb = torch.cat((gamma, beta), 1).permute(0, 2, 3, 1).contiguous().cuda()
    got = ou.modulate(ou.to_cl(x), coef, (4, 8, 8), up=(2, 2, 2), gb=gb, act=2)
    assert rel_inf(ou.from_cl(got), want) < 2e-5
    # relu(IN(a) + IN(b)) and relu(IN(a) + b)  (ResNet bottleneck tails)
    a, b = torch.randn(B, C, 1, 8, 8, generator=g), torch.randn(B, C, 1, 8, 8, generator=g) * 3
    sa, V = ou.channel_stats(ou.to_cl(a))
    sb, _ = ou.channel_stats(ou.to_cl(b))
    ca, cb = ou.norm_coeffs(sa, V, 0), ou.norm_coeffs(sb, V, 0)
    got = ou.modulate(ou.to_cl(a), ca, (1, 8, 8), r=ou.to_cl(b), coef2=cb, act=1)
    assert rel_inf(ou.from_cl(got), F.relu(F.instance_norm(a) + F.instance_norm(b))) < 2e-5
    got = ou.modulate(ou.to_cl(a), ca, (1, 8, 8), r=ou.to_cl(b), act=1)
    assert rel_inf(ou.from_cl(got), F.relu(F.instance_norm(a) + b)) < 2e-5


def test_linear_resize_maxpool():
    g = G(44)
    x, w, b = torch.randn(5, 64, generator=g), torch.randn(300, 64, generator=g), torch.randn(300, generator=g)
    assert rel_inf(ou.linear(x.cuda(), w.cuda(), b.cuda()).cpu(), F.linear(x, w, b)) < 1e-5
    x, w = torch.randn(19, 8192, generator=g), torch.randn(128, 8192, generator=g) * 0.01
    assert rel_inf(ou.linear(x.cuda(), w.cuda(), None).cpu(), F.linear(x, w)) < 1e-5
    # the shapes launch_linear meets on the path: the flow's conditioning GEMM, AdaIN / fc widths, the control=True embedding
    # width (160), ragged batch and feature counts, the embedder fc and conv_mu|var (deep K, few features)
    # (K = 64 with N >= 8192 takes the thread-per-feature kernel: full / ragged / several 64-row chunks, ragged feature count)
    for B_, K_, N_ in ((64, 64, 40960), (7, 128, 1000), (3, 160, 513), (200, 64, 256), (64, 2048, 64), (1, 8192, 128),
                       (6, 64, 40960), (1, 64, 16384), (131, 64, 8200), (70, 64, 8193)):
        x, w, b = torch.randn(B_, K_, generator=g), torch.randn(N_, K_, generator=g) * 0.05, torch.randn(N_, generator=g)
        assert rel_inf(ou.linear(x.cuda(), w.cuda(), b.cuda()).cpu(), F.linear(x.double(), w.double(), b.double())) < 1e-5
    img = torch.rand(3, 3, 64, 64, generator=g) * 2 - 1
    for s in (4, 8, 16, 32, 64):
        want = F.interpolate(img, size=(s, s), mode="bilinear", align_corners=True)
        assert rel_inf(ou.resize(img.cuda(), s, s).permute(0, 3, 1, 2).cpu(), want) < 1e-5
    assert torch.equal(ou.resize(img.cuda(), 64, 64).permute(0, 3, 1, 2).cpu(), img)
    x = torch.randn(2, 64, 32, 32, generator=g)
    assert torch.equal(ou.from_cl(ou.maxpool(ou.to_cl(x))), F.max_pool2d(x, 3, 2, 1))


@pytest.mark.parametrize("C,dims,up,with_gb,with_coef", [
    (64, (4, 16, 16), (2, 2, 2), True, True),      # SPADE pass through the full upsample map (8-channel fast path)
    (128, (4, 8, 8), (1, 2, 2), True, True),       # phase form: a0 kept at T/2, spatial upsample only
    (32, (8, 16, 16), (1, 1, 1), False, True),     # AdaIN pass
    (16, (2, 8, 8), (1, 1, 1), False, False),      # plain lrelu split ahead of conv_img
    (24, (2, 8, 8), (2, 1, 1), True, True),        # C/8 not a power of two -> generic kernel
    (64, (8, 16, 16), (2, 1, 1), True, True),      # SPADE kernel: 4 source planes in flight, each written to 2 output planes
    (32, (2, 8, 8), (1, 1, 1), True, True),        # SPADE kernel: two-plane window
    (64, (16, 8, 8), (1, 2, 2), True, True),       # SPADE kernel: window refilled three times
    (32, (3, 8, 8), (1, 1, 1), True, True),        # odd plane count -> generic T-walking kernel
    (64, (4, 8, 8), (1, 1, 1), False, True),       # map-free pass with coefficients (compile-time lrelu)
])
def test_modulate_split_matches_fp32_pass(C, dims, up, with_gb, with_coef):
    """The fp16 (hi, lo) pair is the exact split of 16 x the fp32 pass: hi = fp16(16 v), lo = fp16(16 v - hi)."""
    g = torch.Generator().manual_seed(C)
    B, (T, H, W) = 3, dims
    x = torch.randn(B, T // up[0], H // up[1], W // up[2], C, generator=g).cuda()
    coef = torch.randn(B, C, 2, generator=g).cuda() if with_coef else None
    gb = (0.3 * torch.randn(B, H, W, 2 * C, generator=g)).cuda() if with_gb else None
    want = ou.modulate(x, coef, dims, up=up, gb=gb, act=2)
    hi, lo = ou.modulate_split(x, coef, dims, up=up, gb=gb, act=2, scale=16.0)
    w16 = want * 16.0
    assert torch.equal(hi, w16.half())
    assert torch.equal(lo, (w16 - w16.half().float()).half())


@pytest.mark.parametrize("T", [4, 2, 3])
def test_modulate_split_second_result_shares_the_read(T):
    """a0 = lrelu(SPADE(x)) and the shortcut's GroupNorm-affine input from one pass over x (BAIR g_4 geometry, scaled down;
    T = 4 / 2: SPADE kernel with a 4- / 2-plane window, T = 3: generic T-walking kernel)."""
    g = torch.Generator().manual_seed(9)
    B, H, W, C = 2, 16, 16, 64
    x = torch.randn(B, T, H, W, C, generator=g).cuda()
    coef, coef_b = torch.randn(B, C, 2, generator=g).cuda(), torch.randn(B, C, 2, generator=g).cuda()
    gb = (0.3 * torch.randn(B, H, W, 2 * C, generator=g)).cuda()
    hi, lo, hb, lb = ou.modulate_split(x, coef, (T, H, W), gb=gb, act=2, coef_b=coef_b)
    w16 = ou.modulate(x, coef, (T, H, W), gb=gb, act=2) * 16.0
    assert torch.equal(hi, w16.half()) and torch.equal(lo, (w16 - w16.half().float()).half())
    b16 = ou.modulate(x, coef_b, (T, H, W), act=0) * 16.0
    assert torch.equal(hb, b16.half()) and torch.equal(lb, (b16 - b16.half().float()).half())


@pytest.mark.parametrize("B,H,W", [(3, 4, 4), (2, 8, 8), (2, 16, 16), (2, 32, 32), (2, 64, 64), (1, 128, 128)])
def test_spade_conv3_equals_simt_engine(B, H, W):
    """The dedicated 3->128 SPADE conv (misc.cu) keeps the SIMT engine's summation order: same bits after the split."""
    g = G(60 + H)
    img = torch.rand(B, H, W, 3, generator=g) * 2 - 1
    w = torch.randn(9, 128, 3, generator=g) * 0.2
    b = torch.randn(128, generator=g) * 0.1
    scale = 64.0
    hi, lo = ou.spade_conv3(img.cuda(), w.cuda(), b.cuda(), scale, act=2)
    y = ou.conv(img.cuda().view(B, 1, H, W, 3), w.cuda(), b.cuda(), None, (1, 3, 3), (1, 1, 1), (0, 1, 1), act=2).view(B, H, W, 128)
    f = y * scale
    want_hi = f.half()
    want_lo = (f - want_hi.float()).half()
    assert torch.equal(hi, want_hi) and torch.equal(lo, want_lo)
    ref = F.leaky_relu(F.conv2d(img.permute(0, 3, 1, 2).double(), w.view(3, 3, 128, 3).permute(2, 3, 0, 1).double(), b.double(), padding=1), 0.2)
    assert rel_inf(((hi.double() + lo.double()) / scale).permute(0, 3, 1, 2).cpu(), ref) < 1e-5
