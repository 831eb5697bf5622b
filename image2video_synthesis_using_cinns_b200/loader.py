"""Load-time weight preprocessing: reference state-dicts -> kernel layouts.

Everything here runs once per checkpoint on the host in fp32 (float64 where a reduction is folded)
and is pure data movement / constant folding of things the reference recomputes on every forward:

* spectral norm: eval-mode ``W = W_orig / (u . W_mat v)`` (decoder.py:20-25; the reference re-divides
  all 150 M decoder weights per call -- 20 % of its CPU time) is folded once;
* BatchNorm (eval) of the 'bn' embedder variants is folded into the preceding conv;
* conv weights go to ``[taps, Cout, Cin]`` (channels-last implicit GEMM), SPADE's gamma/beta convs are
  concatenated into one conv, the decoder's ``fc`` rows are permuted so its output is already the
  channels-last [B,1,4,4,C] tensor ``decoder.py:99`` reshapes to;
* flow MLPs: the first Linear of every subnet is split into its state part and its conditioning
  part (the latter is hoisted out of the sequential chain), scale/translation nets are stacked.
"""
from __future__ import annotations

import torch


def _taps3(w):
    """(Cout, Cin, kt, kh, kw) -> [kt*kh*kw, Cout, Cin]"""
    co, ci = w.shape[:2]
    return w.permute(2, 3, 4, 0, 1).reshape(-1, co, ci).contiguous()


def _taps2(w):
    """(Cout, Cin, kh, kw) -> [kh*kw, Cout, Cin]"""
    co, ci = w.shape[:2]
    return w.permute(2, 3, 0, 1).reshape(-1, co, ci).contiguous()


def fold_spectral(sd, prefix):
    if prefix + ".weight_orig" in sd:
        w = sd[prefix + ".weight_orig"].double()
        sigma = torch.dot(sd[prefix + ".weight_u"].double(), w.reshape(w.shape[0], -1) @ sd[prefix + ".weight_v"].double())
        return (w / sigma).float()
    return sd[prefix + ".weight"].float()


# ------------------------------------------------------------------------------------------ flow
def pack_flow(sd, n_flows, d, cond_channels, hidden, depth, control):
    """ConditionalFlow state-dict (flow_blocks.py / modules.py key layout) -> packed tensors.

    Returns (tensors, cond_mode list, zc_pad).  The conditioning width is zero-padded to a multiple of
    4 (control=True makes it zc+30 = 94/158)."""
    half = d // 2
    zc_pad = (cond_channels + 3) // 4 * 4
    H = hidden
    w1x = torch.zeros(n_flows, 2, 2 * H, half)
    w1c = torch.zeros(n_flows, 2, 2 * H, zc_pad)
    b1 = torch.zeros(n_flows, 2, 2 * H)
    wh = torch.zeros(n_flows, 2, max(depth, 1), 2, H, H)
    bh = torch.zeros(n_flows, 2, max(depth, 1), 2 * H)
    wo = torch.zeros(n_flows, 2, 2 * half, H)
    bo = torch.zeros(n_flows, 2, 2 * half)
    loc = torch.zeros(n_flows, d)
    scale = torch.zeros(n_flows, d)
    pf = torch.zeros(n_flows, d, dtype=torch.int32)
    pb = torch.zeros(n_flows, d, dtype=torch.int32)
    cond_mode = []
    for fl in range(n_flows):
        p = f"sub_layers.{fl}."
        cm = bool(control) and fl % 4 != 0          # flow_blocks.py:24
        cond_mode.append(1 if cm else 0)
        if int(sd[p + "norm_layer.initialized"]) == 0:
            raise ValueError(
                f"{p}norm_layer.initialized == 0: the reference would data-initialise ActNorm on the first "
                "forward even in eval (modules.py:76-78, quirk Q2); refusing to load an untrained flow")
        loc[fl] = sd[p + "norm_layer.loc"].reshape(-1)
        scale[fl] = sd[p + "norm_layer.scale"].reshape(-1)
        pf[fl] = sd[p + "shuffle.forward_shuffle_idx"].to(torch.int32)
        pb[fl] = sd[p + "shuffle.backward_shuffle_idx"].to(torch.int32)
        for i in range(2):
            for ni, net in enumerate(("s", "t")):
                q = f"{p}coupling.{net}.{i}.main."
                w0 = sd[q + "0.weight"].float()
                rows = slice(ni * H, (ni + 1) * H)
                if cm:
                    assert w0.shape == (H, cond_channels), (w0.shape, cond_channels)
                    w1c[fl, i, rows, :cond_channels] = w0
                else:
                    assert w0.shape == (H, half + cond_channels), (w0.shape, half, cond_channels)
                    w1x[fl, i, rows] = w0[:, :half]
                    w1c[fl, i, rows, :cond_channels] = w0[:, half:]
                b1[fl, i, rows] = sd[q + "0.bias"]
                for l in range(depth):
                    wh[fl, i, l, ni] = sd[f"{q}{2 * (l + 1)}.weight"]
                    bh[fl, i, l, rows] = sd[f"{q}{2 * (l + 1)}.bias"]
                orow = slice(ni * half, (ni + 1) * half)
                wo[fl, i, orow] = sd[f"{q}{2 * (depth + 1)}.weight"]
                bo[fl, i, orow] = sd[f"{q}{2 * (depth + 1)}.bias"]
    t = dict(w1x=w1x, w1c=w1c.reshape(-1, zc_pad), b1=b1.reshape(-1), wh=wh, bh=bh, wo=wo, bo=bo, loc=loc,
             scale=scale, perm_fwd=pf, perm_bwd=pb)
    if d == 64 and H in (256, 512):
        t["wpack"] = pack_flow_chunks(w1x, wh, wo, depth)
    return {k: v.contiguous() for k, v in t.items()}, cond_mode, zc_pad


def pack_flow_chunks(w1x, wh, wo, depth):
    """Weight stream of the cluster-resident flow kernel (csrc/flow_cluster.cu).

    A cluster of 16 CTAs carries 8 batch rows through every coupling: CTAs 0..7 hold the scale net, 8..15 the translation
    net; CTA j of a net owns output columns [j*H/8, (j+1)*H/8) of the first and the hidden Linears and outputs [4j, 4j+4)
    of the last one.  Per coupling and CTA the weights are laid out in the order they are consumed, k-major (column
    fastest, so a warp's shared-memory reads are conflict-free), in chunks of 32 k-rows:
        [first Linear, x part: 32 x H/8] [hidden l: H/32 chunks of 32 x H/8] ... [last Linear: H x 4]
    -> [n_flows*2, 16, 1 + depth*H/32 + 1, 32*H/8] float32."""
    n_flows, _, twoH, half = w1x.shape
    H = twoH // 2
    Cc = H // 8
    assert half == 32 and H % 32 == 0
    nc = n_flows * 2
    w1 = w1x.reshape(nc, 2, 8, Cc, half)                                   # [c, net, j, col, k]
    parts = [w1.permute(0, 1, 2, 4, 3).reshape(nc, 16, 1, half * Cc)]      # k-major: [k][col]
    if depth > 0:
        h = wh.reshape(nc, depth, 2, 8, Cc, H // 32, 32)                    # [c, l, net, j, col, chunk, k]
        h = h.permute(0, 2, 3, 1, 5, 6, 4)                                  # [c, net, j, l, chunk, k, col]
        parts.append(h.reshape(nc, 16, depth * (H // 32), 32 * Cc))
    o = wo.reshape(nc, 2, 8, 4, H)                                          # [c, net, j, out, k]
    parts.append(o.permute(0, 1, 2, 4, 3).reshape(nc, 16, 1, H * 4))        # [k][4]
    return torch.cat(parts, dim=2).contiguous()


# --------------------------------------------------------------------------------------- decoder
DEC_BLOCKS = ("head_0", "g_0", "g_1", "g_2", "g_3", "g_4")
ACT_SPLIT_SCALE = 16.0      # must equal ACT_SPLIT_SCALE in csrc/api.cu


def split_fp16(w, in_scale, wmax=None):
    """fp32 weights [taps, Cout, Cin] -> (hi, lo, ws) for the tensor-core engine (csrc/conv_tc.cu).
    ``wmax`` overrides max|w| when several tensors must share one scale (fused shortcut, see pack_decoder).

    hi = fp16(s*w), lo = fp16(s*w - hi) with s the power of two that puts max|w| in [2^13, 2^14): both
    words stay in fp16's normal range for everything that matters and hi+lo carries ~22 significand
    bits.  Rows are zero-padded to a multiple of 16 (UMMA N granularity).  ws = 1/(in_scale*s) is the
    exact power-of-two factor the epilogue applies to the fp32 accumulator."""
    import math
    m = float(w.abs().max()) if wmax is None else float(wmax)
    s = 2.0 ** math.floor(math.log2(2.0 ** 14 / m)) if m > 0 else 1.0
    ws = w.double() * s
    hi = ws.to(torch.float16)
    lo = (ws - hi.double()).to(torch.float16)
    cout = w.shape[1]
    cpad = (cout + 15) // 16 * 16
    if cpad != cout:
        z = torch.zeros(w.shape[0], cpad - cout, w.shape[2], dtype=torch.float16)
        hi, lo = torch.cat((hi, z), 1), torch.cat((lo, z), 1)
    return hi.contiguous(), lo.contiguous(), torch.tensor([1.0 / (in_scale * s)], dtype=torch.float32)


def phase_weights(w_taps):
    """[27, Cout, Cin] (kt, kh, kw) -> [2 phases * 2 taps * 9, Cout, Cin] for a conv whose input was
    nearest-upsampled x2 in time: frames 2j and 2j+1 of the input are identical, so
        out[2j]   = W0 a[j-1] + (W1+W2) a[j]        out[2j+1] = (W0+W1) a[j] + W2 a[j+1]
    (csrc/conv_tc.cu, t_phase).  Sums are formed in float64."""
    w = w_taps.double().reshape(3, 9, *w_taps.shape[1:])
    return torch.stack((w[0], w[1] + w[2], w[0] + w[1], w[2])).reshape(36, *w_taps.shape[1:])


def pack_decoder(sd, nf, engine=0, upsample_t=(2, 1), upsample_s=(2, 2)):
    """Returns (tensors, scalars).  engine >= 1 adds the split fp16 weights of the tensor-core engine.

    Blocks that do not upsample but change the channel count (BAIR / iPER g_4) get their learned shortcut fused into
    conv_1 (csrc/conv_tc.cu, side input): ``<block>.conv_1x`` holds conv_s as a [3 (kw), Cout, Cin] stack whose kw = 1
    slab is the 1x1x1 kernel, split with the scale conv_1 uses (one accumulator, one epilogue factor)."""
    import math
    t, scalars = {}, {}
    c0 = 16 * nf
    # fc output index n = c*16 + (h*4+w)  ->  channels-last index (h*4+w)*C + c
    fw = sd["fc.weight"].float().reshape(c0, 16, -1).permute(1, 0, 2).reshape(16 * c0, -1)
    fb = sd["fc.bias"].float().reshape(c0, 16).t().reshape(-1)
    t["fc.w"], t["fc.b"] = fw, fb
    for name in DEC_BLOCKS:
        t[f"{name}.conv_0.w"] = _taps3(fold_spectral(sd, f"{name}.conv_0"))
        t[f"{name}.conv_0.b"] = sd[f"{name}.conv_0.bias"].float()
        t[f"{name}.conv_1.w"] = _taps3(fold_spectral(sd, f"{name}.conv_1"))
        t[f"{name}.conv_1.b"] = sd[f"{name}.conv_1.bias"].float()
        if f"{name}.conv_s.weight_orig" in sd or f"{name}.conv_s.weight" in sd:
            t[f"{name}.conv_s.w"] = _taps3(fold_spectral(sd, f"{name}.conv_s"))
            t[f"{name}.norm_s.w"] = sd[f"{name}.norm_s.bn.weight"].float()
            t[f"{name}.norm_s.b"] = sd[f"{name}.norm_s.bn.bias"].float()
        t[f"{name}.spade.conv.w"] = _taps2(sd[f"{name}.norm_0.conv.weight"].float())
        t[f"{name}.spade.conv.b"] = sd[f"{name}.norm_0.conv.bias"].float()
        t[f"{name}.spade.gb.w"] = _taps2(torch.cat((sd[f"{name}.norm_0.conv_gamma.weight"],
                                                    sd[f"{name}.norm_0.conv_beta.weight"]), 0).float())
        t[f"{name}.spade.gb.b"] = torch.cat((sd[f"{name}.norm_0.conv_gamma.bias"],
                                             sd[f"{name}.norm_0.conv_beta.bias"]), 0).float()
        t[f"{name}.adain.w"] = sd[f"{name}.norm_1.linear.weight"].float()
        t[f"{name}.adain.b"] = sd[f"{name}.norm_1.linear.bias"].float()
    t["conv_img.w"] = _taps3(sd["conv_img.weight"].float())
    t["conv_img.b"] = sd["conv_img.bias"].float()
    if engine >= 1:
        tc = {}
        for name in DEC_BLOCKS:
            convs = ["conv_0", "conv_1"] + (["conv_s"] if f"{name}.conv_s.w" in t else [])
            # conv_0 of the blocks that run on 16x16+ planes behind a x2 temporal upsample: phase-combined weights
            ut = {"g_0": 2, "g_1": 2, "g_2": 2, "g_3": upsample_t[0], "g_4": upsample_t[1]}.get(name, 0)
            if ut == 2:
                ph, pl, ps = split_fp16(phase_weights(t[f"{name}.conv_0.w"]), ACT_SPLIT_SCALE)   # float64 sums -> split
                tc[f"{name}.conv_0.wph"], tc[f"{name}.conv_0.wpl"], tc[f"{name}.conv_0.wps"] = ph, pl, ps
            us = {"g_0": 2, "g_1": 2, "g_2": 2, "g_3": upsample_s[0], "g_4": upsample_s[1]}.get(name, 0)
            if "conv_s" in convs and ut == 1 and us == 1:
                w1, wsc = t[f"{name}.conv_1.w"], t[f"{name}.conv_s.w"]                # [27,Cout,Cmid], [1,Cout,Cin]
                wmax = max(float(w1.abs().max()), float(wsc.abs().max()))
                stack = torch.zeros(3, *wsc.shape[1:])
                stack[1] = wsc[0]
                tc[f"{name}.conv_1x.wh"], tc[f"{name}.conv_1x.wl"], _ = split_fp16(stack, ACT_SPLIT_SCALE, wmax)
                tc[f"{name}.conv_1.wh"], tc[f"{name}.conv_1.wl"], tc[f"{name}.conv_1.ws"] = split_fp16(t.pop(f"{name}.conv_1.w"),
                                                                                                   ACT_SPLIT_SCALE, wmax)
                convs = [c for c in convs if c != "conv_1"]
            for c in convs:
                tc[f"{name}.{c}.wh"], tc[f"{name}.{c}.wl"], tc[f"{name}.{c}.ws"] = split_fp16(t.pop(f"{name}.{c}.w"), ACT_SPLIT_SCALE)
            # SPADE hidden map h = lrelu(conv(img) + b) with |img| <= 1 after the bilinear resize:
            # |h| <= max_c (sum|W_c| + |b_c|).  Its split scale targets 2^11 at that bound (x8 input headroom).
            w, b = t[f"{name}.spade.conv.w"], t[f"{name}.spade.conv.b"]
            bound = float((w.abs().sum(dim=(0, 2)) + b.abs()).max())
            sa = 2.0 ** math.floor(math.log2(2.0 ** 11 / max(bound, 1e-30)))
            scalars[f"{name}.spade.sa"] = sa
            tc[f"{name}.spade.gb.wh"], tc[f"{name}.spade.gb.wl"], tc[f"{name}.spade.gb.ws"] = split_fp16(t.pop(f"{name}.spade.gb.w"), sa)
        tc["conv_img.wh"], tc["conv_img.wl"], tc["conv_img.ws"] = split_fp16(t.pop("conv_img.w"), ACT_SPLIT_SCALE)
        t.update(tc)
    return {k: v.contiguous() for k, v in t.items()}, scalars


# -------------------------------------------------------------------------------------- embedder
def pack_embedder(sd, zc, norm, tensor_core=True):
    """torchvision-resnet50 keys under ``model.`` (AE.py:109) -> channels-last conv stacks.

    Every stride-1 conv also gets the split fp16 weights of the tensor-core engine (``.wh/.wl/.ws``, csrc/conv_tc.cu; for
    the BatchNorm variant these are the BN-folded weights); the kernel side decides per call which layers use them
    (csrc/api.cu, embedder_run), so the fp32 copy stays registered too."""
    t = {}

    def put(dst, conv_key, bn_key, stride=1):
        w = sd[conv_key].double()
        if norm == "bn":
            g, b = sd[bn_key + ".weight"].double(), sd[bn_key + ".bias"].double()
            mean, var = sd[bn_key + ".running_mean"].double(), sd[bn_key + ".running_var"].double()
            sc = g / torch.sqrt(var + 1e-5)
            w = w * sc.reshape(-1, 1, 1, 1)
            t[dst + ".b"] = (b - mean * sc).float()
        t[dst + ".w"] = _taps2(w.float())
        if tensor_core and stride == 1 and w.shape[1] % 16 == 0:
            t[dst + ".wh"], t[dst + ".wl"], t[dst + ".ws"] = split_fp16(t[dst + ".w"], ACT_SPLIT_SCALE)

    put("conv1", "model.conv1.weight", "model.bn1", stride=2)
    for li, nb in enumerate((3, 4, 6, 3)):
        for bi in range(nb):
            p, q = f"model.layer{li + 1}.{bi}.", f"layer{li + 1}.{bi}."
            down = 2 if (li > 0 and bi == 0) else 1          # resnet v1.5: the stride sits on conv2 and the downsample conv
            for j in (1, 2, 3):
                put(f"{q}conv{j}", f"{p}conv{j}.weight", f"{p}bn{j}", stride=down if j == 2 else 1)
            if bi == 0:
                put(f"{q}ds", f"{p}downsample.0.weight", f"{p}downsample.1", stride=down)
    fcw = sd["model.fc.sub_layers.0.weight"].float()
    if fcw.shape[2] != 1 or fcw.shape[3] != 1:
        raise ValueError("embedder fc kernel is not 1x1: input size does not reduce to 1x1 before the fc")
    t["fc.w"] = fcw.reshape(fcw.shape[0], -1)[:zc]      # mode() keeps the first zc channels
    t["fc.b"] = sd["model.fc.sub_layers.0.bias"].float()[:zc]
    return {k: v.contiguous() for k, v in t.items()}


# ------------------------------------------------------------------------------------ 3-D encoder
def pack_encoder3d(sd, n_stages=4, blocks=2, tensor_core=True):
    """resnet3D.Encoder keys -> channels-last conv stacks [27, Cout, Cin].  Every 3x3x3 conv of the blocks also gets the split
    fp16 weights of the tensor-core engine (``.wh/.wl/.ws``, csrc/conv_tc.cu); the kernel side uses them for the stride-1
    convs whose GEMM fills the machine (csrc/api.cu, encoder3d_run) and the fp32 copy everywhere else."""
    t = {"conv1.w": _taps3(sd["conv1.weight"].float()), "norm1.w": sd["norm1.weight"].float(),
         "norm1.b": sd["norm1.bias"].float()}

    def put_split(dst):
        if tensor_core and t[dst + ".w"].shape[2] % 16 == 0:
            t[dst + ".wh"], t[dst + ".wl"], t[dst + ".ws"] = split_fp16(t[dst + ".w"], ACT_SPLIT_SCALE)

    for l in range(n_stages):
        for b in range(blocks):
            p = f"layer.{l}.{b}."
            t[p + "conv1.w"] = _taps3(sd[p + "conv1.weight"].float())
            t[p + "conv2.w"] = _taps3(sd[p + "conv2.weight"].float())
            put_split(p + "conv1")
            put_split(p + "conv2")
            for nm in ("bn1", "bn2"):
                t[p + nm + ".w"], t[p + nm + ".b"] = sd[p + nm + ".weight"].float(), sd[p + nm + ".bias"].float()
            if p + "downsample.0.weight" in sd:
                t[p + "ds.w"] = _taps3(sd[p + "downsample.0.weight"].float())
                put_split(p + "ds")
                t[p + "ds.gn.w"] = sd[p + "downsample.1.weight"].float()
                t[p + "ds.gn.b"] = sd[p + "downsample.1.bias"].float()
    # conv_mu / conv_var: (z, C, 4, 4) valid conv on the 4x4 map -> Linear over the (h, w, c) flattening
    mv = torch.cat((sd["conv_mu.weight"], sd["conv_var.weight"]), 0).float()
    t["muvar.w"] = mv.permute(0, 2, 3, 1).reshape(mv.shape[0], -1)
    t["muvar.b"] = torch.cat((sd["conv_mu.bias"], sd["conv_var.bias"]), 0).float()
    return {k: v.contiguous() for k, v in t.items()}
