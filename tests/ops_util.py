"""ctypes wrappers over the single-kernel C-ABI entry points, for the parity tests (GPU only)."""
import ctypes

import torch

from image2video_synthesis_using_cinns_b200 import lib as _lib


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def S():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_cl(x):
    """NCTHW / NCHW -> channels-last contiguous cuda tensor."""
    if x.dim() == 5:
        return x.permute(0, 2, 3, 4, 1).contiguous().cuda()
    return x.permute(0, 2, 3, 1).contiguous().cuda()


def from_cl(y):
    if y.dim() == 5:
        return y.permute(0, 4, 1, 2, 3).contiguous().cpu()
    return y.permute(0, 3, 1, 2).contiguous().cpu()


def taps(w):
    co, ci = w.shape[:2]
    perm = (2, 3, 4, 0, 1) if w.dim() == 5 else (2, 3, 0, 1)
    return w.permute(*perm).reshape(-1, co, ci).contiguous().cuda()


def conv(x_cl, w_taps, bias, res_cl, k, stride, pad, res_up=(1, 1, 1), act=0, out_mode=0, engine=0):
    """x_cl [B,T,H,W,Cin] cuda; k/stride/pad are (t,h,w) triples."""
    L = _lib.load()
    B, Ti, Hi, Wi, Cin = x_cl.shape
    Cout = w_taps.shape[1]
    To = (Ti + 2 * pad[0] - k[0]) // stride[0] + 1
    Ho = (Hi + 2 * pad[1] - k[1]) // stride[1] + 1
    Wo = (Wi + 2 * pad[2] - k[2]) // stride[2] + 1
    shape = (B, To, Ho, Wo, Cout) if out_mode == 0 else (B, To, Cout, Ho, Wo)
    y = torch.empty(shape, dtype=torch.float32, device="cuda")
    _lib.check(L.i2v_op_conv(P(x_cl), P(w_taps), P(bias), P(res_cl), P(y), B, Ti, Hi, Wi, Cin, Cout, *k, *stride, *pad,
                             *res_up, act, out_mode, engine, S()), "op_conv")
    return y


def spade_conv3(img_cl, w_taps, bias, scale, act=2):
    """img_cl [B,H,W,3], w_taps [9,128,3] -> (hi, lo) fp16 [B,H,W,128]"""
    L = _lib.load()
    B, H, W, _ = img_cl.shape
    hi = torch.empty(B, H, W, 128, dtype=torch.float16, device="cuda")
    lo = torch.empty_like(hi)
    _lib.check(L.i2v_op_spade_conv3(P(img_cl), P(w_taps), P(bias), P(hi), P(lo), float(scale), B, H, W, act, S()), "op_spade_conv3")
    return hi, lo


def channel_stats(x_cl):
    L = _lib.load()
    B, C = x_cl.shape[0], x_cl.shape[-1]
    V = x_cl.numel() // (B * C)
    sums = torch.empty(B, C, 2, dtype=torch.float64, device="cuda")
    _lib.check(L.i2v_op_channel_stats(P(x_cl), P(sums), B, V, C, S()), "op_channel_stats")
    return sums, V


def norm_coeffs(sums, V, groups, gamma=None, beta=None, mod=None, eps=1e-5):
    L = _lib.load()
    B, C = sums.shape[:2]
    coef = torch.empty(B, C, 2, dtype=torch.float32, device="cuda")
    _lib.check(L.i2v_op_norm_coeffs(P(sums), P(coef), B, C, V, groups, eps, P(gamma), P(beta), P(mod), S()), "op_norm_coeffs")
    return coef


def modulate(x_cl, coef, out_dims, up=(1, 1, 1), gb=None, r=None, coef2=None, act=0):
    L = _lib.load()
    B, C = x_cl.shape[0], x_cl.shape[-1]
    T, H, W = out_dims
    out = torch.empty(B, T, H, W, C, dtype=torch.float32, device="cuda")
    _lib.check(L.i2v_op_modulate(P(x_cl), P(coef), P(gb), P(r), P(coef2), P(out), B, T, H, W, C, *up, act, S()), "op_modulate")
    return out


def modulate_split(x_cl, coef, out_dims, up=(1, 1, 1), gb=None, act=0, scale=16.0, coef_b=None):
    """Conv-ready fp16 pair (hi, lo) of scale * modulate(...); with coef_b also the pair of scale * (A_b x + B_b)."""
    L = _lib.load()
    B, C = x_cl.shape[0], x_cl.shape[-1]
    T, H, W = out_dims
    hi = torch.empty(B, T, H, W, C, dtype=torch.float16, device="cuda")
    lo = torch.empty_like(hi)
    hb = torch.empty_like(hi) if coef_b is not None else None
    lb = torch.empty_like(hi) if coef_b is not None else None
    _lib.check(L.i2v_op_modulate_split(P(x_cl), P(coef), P(gb), P(hi), P(lo), B, T, H, W, C, *up, act, scale, P(coef_b), P(hb),
                                       P(lb), S()), "op_modulate_split")
    return (hi, lo) if coef_b is None else (hi, lo, hb, lb)


def linear(x, w, b, act=0):
    L = _lib.load()
    B, K = x.shape
    N = w.shape[0]
    y = torch.empty(B, N, dtype=torch.float32, device="cuda")
    _lib.check(L.i2v_op_linear(P(x), P(w), P(b), P(y), B, K, N, act, S()), "op_linear")
    return y


def resize(img, H, W):
    L = _lib.load()
    B, C, H0, W0 = img.shape
    out = torch.empty(B, H, W, C, dtype=torch.float32, device="cuda")
    _lib.check(L.i2v_op_resize_bilinear(P(img), P(out), B, C, H0, W0, H, W, S()), "op_resize")
    return out


def maxpool(x_cl):
    L = _lib.load()
    B, H, W, C = x_cl.shape
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = torch.empty(B, Ho, Wo, C, dtype=torch.float32, device="cuda")
    _lib.check(L.i2v_op_maxpool3x3s2(P(x_cl), P(y), B, H, W, C, S()), "op_maxpool")
    return y


def conv_tc(x_cl, w_taps, bias, res_cl, k, res_up=(1, 1, 1), act=0, out_mode=0, terms=3, scale_a=16.0, variant=0):
    """Tensor-core engine on fp32 inputs (the op splits them on the device).  w_taps [taps,Cout,Cin]."""
    import math
    L = _lib.load()
    B, T, H, W, Cin = x_cl.shape
    Cout = w_taps.shape[1]
    cpad = (Cout + 15) // 16 * 16
    wp = torch.zeros(w_taps.shape[0], cpad, Cin, device="cuda")
    wp[:, :Cout] = w_taps
    scale_w = 2.0 ** math.floor(math.log2(2.0 ** 14 / float(w_taps.abs().max())))
    shape = (B, T, H, W, Cout) if out_mode == 0 else (B, T, Cout, H, W)
    y = torch.empty(shape, dtype=torch.float32, device="cuda")
    ws = torch.empty(4 * (x_cl.numel() + wp.numel()) + 4096, dtype=torch.uint8, device="cuda")
    _lib.check(L.i2v_op_conv_tc(P(x_cl), P(wp), P(bias), P(res_cl), P(y), B, T, H, W, Cin, Cout, cpad, *k, *res_up, act,
                                out_mode, terms, variant, scale_a, scale_w, P(ws), ws.numel(), S()), "op_conv_tc")
    return y


def conv_tc_side(x_cl, w_taps, x2_cl, w2, bias, act=0, terms=3, scale_a=16.0, variant=0):
    """3x3x3 tensor-core conv of x plus a fused 1x1x1 conv of the side input x2 (w2 [Cout, Cin2])."""
    import math
    L = _lib.load()
    B, T, H, W, Cin = x_cl.shape
    Cin2 = x2_cl.shape[-1]
    Cout = w_taps.shape[1]
    cpad = (Cout + 15) // 16 * 16
    wp = torch.zeros(27, cpad, Cin, device="cuda")
    wp[:, :Cout] = w_taps
    w2p = torch.zeros(3, cpad, Cin2, device="cuda")
    w2p[1, :Cout] = w2
    wmax = max(float(w_taps.abs().max()), float(w2.abs().max()))
    scale_w = 2.0 ** math.floor(math.log2(2.0 ** 14 / wmax))
    y = torch.empty(B, T, H, W, Cout, dtype=torch.float32, device="cuda")
    ws = torch.empty(4 * (x_cl.numel() + wp.numel() + x2_cl.numel() + w2p.numel()) + 8192, dtype=torch.uint8, device="cuda")
    _lib.check(L.i2v_op_conv_tc_side(P(x_cl), P(wp), P(x2_cl), P(w2p), P(bias), P(y), B, T, H, W, Cin, Cin2, Cout, cpad, act, 0,
                                     terms, variant, scale_a, scale_w, P(ws), ws.numel(), S()), "op_conv_tc_side")
    return y


def conv_tc_phase(x_half_cl, w_taps, bias, terms=3, scale_a=16.0, variant=0):
    """Temporal phase form: x_half_cl [B,T/2,H,W,Cin] is the PRE-upsample tensor, w_taps [27,Cout,Cin] the original
    3x3x3 kernel (phase-combined here with the product's own loader.phase_weights); returns [B,T,H,W,Cout]."""
    import math
    from image2video_synthesis_using_cinns_b200 import loader
    L = _lib.load()
    B, Th, H, W, Cin = x_half_cl.shape
    Cout = w_taps.shape[1]
    cpad = (Cout + 15) // 16 * 16
    wp = torch.zeros(36, cpad, Cin, device="cuda")
    wp[:, :Cout] = loader.phase_weights(w_taps.cpu()).float().cuda()
    scale_w = 2.0 ** math.floor(math.log2(2.0 ** 14 / float(wp.abs().max())))
    y = torch.empty(B, 2 * Th, H, W, Cout, dtype=torch.float32, device="cuda")
    ws = torch.empty(4 * (x_half_cl.numel() + wp.numel()) + 4096, dtype=torch.uint8, device="cuda")
    _lib.check(L.i2v_op_conv_tc_phase(P(x_half_cl), P(wp), P(bias), P(y), B, 2 * Th, H, W, Cin, Cout, cpad, terms, variant,
                                      scale_a, scale_w, P(ws), ws.numel(), S()), "op_conv_tc_phase")
    return y
