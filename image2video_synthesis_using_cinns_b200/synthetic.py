"""Synthetic, well-conditioned checkpoints in the reference's exact on-disk format.

No pretrained weights ship with the reference (README.md:45 links Google-Drive files) and there is
no network, so benchmarks and parity tests use random-init weights written in the layout
``get_model.Model.__init__`` reads (get_model.py:15-45, stage2_cINN/modules/INN.py:36-41):

    <root>/stage2/config_stage2.yaml, cINN.pth                      (ConditionalFlow keys)
    <root>/stage1/s1/config_stage1.yaml, best_PFVD_GEN.pth, best_PFVD_ENC.pth.tar
    <root>/ae/ae/config_stage2_AE.yaml, Encoder_stage2.pth          (ResnetEncoder keys)

Every file is ``torch.save({'state_dict': ...})``.  Key names, shapes and dtypes follow the
reference constructors (tests/test_synthetic_layout.py checks them against the real constructors
whenever the reference tree is present); the *values* follow SURVEY.md section 8d so the pipeline is
numerically meaningful: ActNorm ``initialized=1`` with mild loc/scale (quirk Q2), spectral-norm
``u``/``v`` converged by power iteration (otherwise sigma ~ 0.003 and tanh saturates), SPADE
convolutions Xavier-uniform with the reference's gain (decoder.py:86-95).

This module never touches the GPU and has no dependency on the reference tree or the oracle.
"""
from __future__ import annotations

import math
import os

import torch
import yaml

from .config import DATASETS


# --------------------------------------------------------------------------- initialisers
def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def _linear(gen, out_f, in_f):
    b = 1.0 / math.sqrt(in_f)
    return _uniform(gen, (out_f, in_f), b), _uniform(gen, (out_f,), b)


def _conv(gen, cout, cin, *k, bias=True):
    fan_in = cin * math.prod(k)
    b = 1.0 / math.sqrt(fan_in)
    w = _uniform(gen, (cout, cin, *k), b)
    return (w, _uniform(gen, (cout,), b)) if bias else (w, None)


def _kaiming_fan_out(gen, cout, cin, *k):
    std = math.sqrt(2.0 / (cout * math.prod(k)))
    return torch.randn((cout, cin, *k), generator=gen) * std


def _xavier_uniform(gen, cout, cin, *k, gain=0.02):
    rf = math.prod(k)
    bound = gain * math.sqrt(6.0 / (cin * rf + cout * rf))
    return _uniform(gen, (cout, cin, *k), bound)


def _power_iterate(w, gen, n_iter=15, eps=1e-12):
    """u, v of torch's legacy spectral_norm after ``n_iter`` train-mode forward passes."""
    wm = w.reshape(w.shape[0], -1)
    u = torch.nn.functional.normalize(torch.randn(wm.shape[0], generator=gen), dim=0, eps=eps)
    v = torch.nn.functional.normalize(torch.randn(wm.shape[1], generator=gen), dim=0, eps=eps)
    for _ in range(n_iter):
        v = torch.nn.functional.normalize(wm.t() @ u, dim=0, eps=eps)
        u = torch.nn.functional.normalize(wm @ v, dim=0, eps=eps)
    return u, v


# --------------------------------------------------------------------------- state dicts
def flow_state_dict(gen, in_channels=64, cond_channels=64, hidden=512, depth=2, n_flows=20,
                    control=False):
    """Keys of ``ConditionalFlow`` (flow_blocks.py:8-29,63-76,108-115,142-148; modules.py:9-41)."""
    sd = {}
    half = in_channels // 2
    for fl in range(n_flows):
        p = f"sub_layers.{fl}."
        mode_cond = bool(control) and fl % 4 != 0
        d_in = cond_channels if mode_cond else half + cond_channels
        sd[p + "norm_layer.loc"] = (torch.randn(1, in_channels, 1, 1, generator=gen) * 0.1)
        sd[p + "norm_layer.scale"] = 0.8 + torch.rand(1, in_channels, 1, 1, generator=gen) * 0.45
        sd[p + "norm_layer.initialized"] = torch.tensor(1, dtype=torch.uint8)
        for net in ("s", "t"):
            for i in range(2):
                q = f"{p}coupling.{net}.{i}.main."
                dims = [d_in] + [hidden] * (depth + 1) + [half]
                for li in range(depth + 2):
                    w, b = _linear(gen, dims[li + 1], dims[li])
                    sd[f"{q}{2 * li}.weight"], sd[f"{q}{2 * li}.bias"] = w, b
        idx = torch.randperm(in_channels, generator=gen)
        sd[p + "shuffle.forward_shuffle_idx"] = idx
        sd[p + "shuffle.backward_shuffle_idx"] = torch.argsort(idx)
    return sd


def decoder_state_dict(gen, nf=64, z_dim=64, spectral=True, spade_gain=0.02):
    """Keys of ``Generator`` (decoder.py:7-31,55-84; normalization_layer.py:5-16,27-45)."""
    sd = {}
    sd["fc.weight"], sd["fc.bias"] = _linear(gen, 16 * 16 * nf, z_dim)

    def conv3(prefix, cout, cin, k, bias):
        w, b = _conv(gen, cout, cin, k, k, k, bias=bias)
        if bias:
            sd[prefix + ".bias"] = b
        if spectral:
            u, v = _power_iterate(w, gen)
            sd[prefix + ".weight_orig"], sd[prefix + ".weight_u"], sd[prefix + ".weight_v"] = w, u, v
        else:
            sd[prefix + ".weight"] = w

    blocks = [("head_0", 16, 16), ("g_0", 16, 16), ("g_1", 16, 8), ("g_2", 8, 4), ("g_3", 4, 2),
              ("g_4", 2, 1)]
    for name, a, b in blocks:
        n_in, n_out = a * nf, b * nf
        n_mid = min(n_in, n_out)
        conv3(f"{name}.conv_0", n_mid, n_in, 3, True)
        conv3(f"{name}.conv_1", n_out, n_mid, 3, True)
        if n_in != n_out:
            conv3(f"{name}.conv_s", n_out, n_in, 1, False)
            sd[f"{name}.norm_s.bn.weight"] = 1.0 + 0.1 * torch.randn(n_in, generator=gen)
            sd[f"{name}.norm_s.bn.bias"] = 0.1 * torch.randn(n_in, generator=gen)
        sd[f"{name}.norm_0.conv.weight"] = _xavier_uniform(gen, 128, 3, 3, 3, gain=spade_gain)
        sd[f"{name}.norm_0.conv.bias"] = torch.zeros(128)
        for gb in ("conv_gamma", "conv_beta"):
            sd[f"{name}.norm_0.{gb}.weight"] = _xavier_uniform(gen, n_in, 128, 3, 3, gain=spade_gain)
            sd[f"{name}.norm_0.{gb}.bias"] = torch.zeros(n_in)
        sd[f"{name}.norm_1.linear.weight"], sd[f"{name}.norm_1.linear.bias"] = _linear(
            gen, 2 * n_mid, z_dim)
    sd["conv_img.weight"], sd["conv_img.bias"] = _conv(gen, 3, nf, 3, 3, 3)
    return sd


def encoder3d_state_dict(gen, channels, stride_s, z_dim=64, layers=(2, 2, 2, 2)):
    """Keys of the 3-D ``Encoder`` (resnet3D.py:138-200): ResNet-18, GroupNorm(16), bias-free.

    Two reference facts shape the key set: ``self.inplanes = 64`` is hard-coded (resnet3D.py:140), so
    ``channels[0]`` must be 64, and a block gets a (3x3x3) downsample branch only when its *spatial*
    stride != 1 or the width changes (resnet3D.py:184, the temporal stride is not consulted)."""
    if channels[0] != 64:
        raise ValueError("reference Encoder hard-codes inplanes=64 (resnet3D.py:140)")
    sd = {}

    def gn(prefix, c):
        sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=gen)
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=gen)

    sd["conv1.weight"] = _kaiming_fan_out(gen, channels[0], 3, 3, 7, 7)
    gn("norm1", channels[0])
    inpl = channels[0]
    for li, ch in enumerate(channels[1:]):
        for bi in range(layers[li]):
            p = f"layer.{li}.{bi}."
            sd[p + "conv1.weight"] = _kaiming_fan_out(gen, ch, inpl if bi == 0 else ch, 3, 3, 3)
            gn(p + "bn1", ch)
            sd[p + "conv2.weight"] = _kaiming_fan_out(gen, ch, ch, 3, 3, 3)
            gn(p + "bn2", ch)
        if stride_s[li] != 1 or inpl != ch:
            sd[f"layer.{li}.0.downsample.0.weight"] = _kaiming_fan_out(gen, ch, inpl, 3, 3, 3)
            gn(f"layer.{li}.0.downsample.1", ch)
        inpl = ch
    for nm in ("conv_mu", "conv_var"):
        sd[nm + ".weight"], sd[nm + ".bias"] = _conv(gen, z_dim, channels[-1], 4, 4)
    return sd


RESNET50_LAYERS = (3, 4, 6, 3)
RESNET50_PLANES = (64, 128, 256, 512)


def embedder_state_dict(gen, z_dim=64, norm="in"):
    """Keys of ``ResnetEncoder`` (AE.py:91-124): torchvision resnet50 v1.5 under ``model.``, the fc
    replaced by a 1x1 ``DenseEncoderLayer`` conv to 2*z_dim (AE.py:54-81,121-124)."""
    sd = {}

    def nrm(prefix, c):
        if norm == "bn":
            # eval-mode BatchNorm does not re-normalise, so random weights must keep the residual
            # trunk's variance bounded: the last norm of every branch gets a small gain.
            gain = 0.3 if prefix.endswith("bn3") else 1.0
            sd[prefix + ".weight"] = gain * (1.0 + 0.1 * torch.randn(c, generator=gen))
            sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=gen)
            sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=gen)
            sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=gen)
            sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        elif norm != "in":
            raise ValueError(f"unsupported embedder norm {norm!r} (reference ships 'in' and 'bn')")

    def cw(cout, cin, k):
        if norm == "bn":   # variance-preserving (fan-in) init, see nrm()
            return torch.randn((cout, cin, k, k), generator=gen) * math.sqrt(2.0 / (cin * k * k))
        return _kaiming_fan_out(gen, cout, cin, k, k)   # torchvision's default resnet init

    sd["model.conv1.weight"] = cw(64, 3, 7)
    nrm("model.bn1", 64)
    inpl = 64
    for li, (nb, pl) in enumerate(zip(RESNET50_LAYERS, RESNET50_PLANES)):
        for bi in range(nb):
            p = f"model.layer{li + 1}.{bi}."
            sd[p + "conv1.weight"] = cw(pl, inpl, 1)
            nrm(p + "bn1", pl)
            sd[p + "conv2.weight"] = cw(pl, pl, 3)
            nrm(p + "bn2", pl)
            sd[p + "conv3.weight"] = cw(4 * pl, pl, 1)
            nrm(p + "bn3", 4 * pl)
            if bi == 0:
                sd[p + "downsample.0.weight"] = cw(4 * pl, inpl, 1)
                nrm(p + "downsample.1", 4 * pl)
            inpl = 4 * pl
    w, b = _conv(gen, 2 * z_dim, 2048, 1, 1)
    sd["model.fc.sub_layers.0.weight"], sd["model.fc.sub_layers.0.bias"] = w, b
    return sd


# --------------------------------------------------------------------------- files
def write_synthetic_checkpoints(root, dataset="bair", seed=0, nf=None, n_flows=None, hidden_factor=None,
                                enc_channels=None, control=None, spade_gain=0.02, with_encoder=True,
                                img_size=None):
    """Write the 3 YAML + 4 checkpoint files; return the ``model_path`` for ``Model(...)``.

    ``nf`` / ``n_flows`` / ``hidden_factor`` / ``enc_channels`` shrink the model for fast tests (the
    architecture and key set stay authentic); ``control`` overrides ``Training.control``.
    """
    g = DATASETS[dataset]
    nf = g["nf"] if nf is None else nf
    n_flows = 20 if n_flows is None else n_flows
    hidden_factor = 8 if hidden_factor is None else hidden_factor
    enc = dict(g["enc"])
    if enc_channels is not None:
        enc["channels"] = list(enc_channels)
    control = g["control"] if control is None else control
    img_size = g["img_size"] if img_size is None else img_size
    z_dim = 64
    gen = torch.Generator().manual_seed(seed)

    root = os.path.abspath(root)
    d2, d1, dae = (os.path.join(root, "stage2"), os.path.join(root, "stage1", "s1"),
                   os.path.join(root, "ae", "ae"))
    for d in (d2, d1, dae):
        os.makedirs(d, exist_ok=True)

    training = {"bs": 50}
    if control is not None:
        training["control"] = bool(control)
    cfg2 = {
        "Flow": {"n_flows": n_flows, "flow_hidden_depth": 2, "flow_mid_channels_factor": hidden_factor},
        "Conditioning_Model": {"z_dim": g["cond_z"], "checkpoint_name": "Encoder_stage2",
                               "model_name": "ae", "model_path": os.path.join(root, "ae") + "/"},
        "First_stage_model": {"checkpoint_encoder": "best_PFVD_ENC", "checkpoint_decoder": "best_PFVD_GEN",
                              "model_name": "s1", "model_path": os.path.join(root, "stage1") + "/"},
        "Training": training,
        "Data": {"sequence_length": 17, "img_size": img_size, "dataset": g["dataset"]},
    }
    cfg1 = {
        "Decoder": {"channel_factor": nf, "z_dim": z_dim, "upsample_s": list(g["upsample_s"]),
                    "upsample_t": list(g["upsample_t"]), "spectral_norm": True},
        "Encoder": {"res_type_encoder": "resnet18", "deterministic": False, "use_max_pool": False,
                    "z_dim": z_dim, "channels": enc["channels"], "stride_t": enc["stride_t"],
                    "stride_s": enc["stride_s"]},
        "Data": {"sequence_length": 17, "img_size": img_size, "dataset": g["dataset"]},
    }
    cfgae = {"AE": {"deterministic": False, "in_size": img_size, "norm": g["ae_norm"],
                    "encoder_type": "resnet50", "use_actnorm_in_dec": False, "z_dim": g["cond_z"]}}
    for path, cfg in ((os.path.join(d2, "config_stage2.yaml"), cfg2),
                      (os.path.join(d1, "config_stage1.yaml"), cfg1),
                      (os.path.join(dae, "config_stage2_AE.yaml"), cfgae)):
        with open(path, "w") as f:
            yaml.safe_dump(cfg, f)

    cond_ch = g["cond_z"] + (30 if control else 0)
    torch.save({"state_dict": flow_state_dict(gen, z_dim, cond_ch, z_dim * hidden_factor, 2, n_flows,
                                              control=bool(control))},
               os.path.join(d2, "cINN.pth"))
    torch.save({"state_dict": decoder_state_dict(gen, nf, z_dim, True, spade_gain)},
               os.path.join(d1, "best_PFVD_GEN.pth"))
    if with_encoder:
        torch.save({"state_dict": encoder3d_state_dict(gen, enc["channels"], enc["stride_s"], z_dim)},
                   os.path.join(d1, "best_PFVD_ENC.pth.tar"))
    torch.save({"state_dict": embedder_state_dict(gen, g["cond_z"], g["ae_norm"])},
               os.path.join(dae, "Encoder_stage2.pth"))
    return d2 + "/"
