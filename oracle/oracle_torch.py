"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's sampling path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this file, and only as the checker / CPU baseline.  The product path
(``image2video_synthesis_using_cinns_b200``) never imports it and fails loudly without its CUDA
library.

The reference is pure PyTorch (no native code), so the oracle is a *functional* fp32 PyTorch
restatement of the arithmetic, written against state-dict tensors (no ``nn.Module`` graph), every
function citing the reference lines it follows.  It runs wherever torch runs (CPU by default).

Pinning: the reference holds NO golden vectors or tests for this path (SURVEY.md section 4 / 8c:
"parity unpinned" by the reference itself).  The oracle is therefore pinned against the reference's
own modules executed in the build container: ``tests/test_oracle_vs_reference.py`` (runs whenever
``/root/reference`` exists) and the committed fixtures ``tests/golden/*.npz`` produced from the real
reference by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------ flow
def _mlp(sd, prefix, x, depth=2):
    """BasicFullyConnectedNet (stage2_cINN/modules/modules.py:9-30): Linear, then depth x
    (LeakyReLU(0.01), Linear), LeakyReLU, Linear; no tanh / bn (flow_blocks.py:68-75)."""
    n = depth + 2
    for li in range(n):
        x = F.linear(x, sd[f"{prefix}.main.{2 * li}.weight"], sd[f"{prefix}.main.{2 * li}.bias"])
        if li < n - 1:
            x = F.leaky_relu(x, 0.01)
    return x


def flow_block_modes(n_flows, control):
    """'cond' blocks see only the conditioning (flow_blocks.py:24): fl % 4 != 0 and control."""
    return [bool(control) and fl % 4 != 0 for fl in range(n_flows)]


def flow_reverse(sd, x, cond, n_flows, control=False, depth=2, trace=None):
    """ConditionalFlow.forward(reverse=True) (flow_blocks.py:52-57): blocks in reverse order, each
    Shuffle^-1 -> coupling^-1 -> InvLeakyRelu^-1 -> ActNorm^-1 (flow_blocks.py:130-136)."""
    modes = flow_block_modes(n_flows, control)
    x = x.reshape(x.shape[0], -1)
    half = x.shape[1] // 2
    for fl in reversed(range(n_flows)):
        p = f"sub_layers.{fl}."
        x = x[:, sd[p + "shuffle.backward_shuffle_idx"]]                # flow_blocks.py:153-154
        for i in (1, 0):                                                  # flow_blocks.py:96-105
            if i % 2 == 0:
                x = torch.cat((x[:, half:], x[:, :half]), dim=1)
            xa, xk = x[:, :half], x[:, half:]
            ci = cond if modes[fl] else torch.cat((xa, cond), dim=1)
            s = _mlp(sd, f"{p}coupling.s.{i}", ci, depth)
            t = _mlp(sd, f"{p}coupling.t.{i}", ci, depth)
            x = torch.cat((xa, (xk - t) * s.neg().exp()), dim=1)
        scaling = (x >= 0).to(x) + (x < 0).to(x) * 0.9                    # flow_blocks.py:184-187
        x = x / scaling
        x = x / sd[p + "norm_layer.scale"].reshape(1, -1) - sd[p + "norm_layer.loc"].reshape(1, -1)
        if trace is not None:
            trace.append(x.clone())
    return x


def flow_forward(sd, x, cond, n_flows, control=False, depth=2):
    """ConditionalFlow.forward (flow_blocks.py:42-51): ActNorm -> InvLeakyRelu -> coupling ->
    Shuffle per block; logdet = sum of ActNorm log|scale| (modules.py:86-88, H=W=1) and coupling
    sum(s) (flow_blocks.py:93); InvLeakyRelu and Shuffle report 0 (quirk Q3)."""
    modes = flow_block_modes(n_flows, control)
    x = x.reshape(x.shape[0], -1)
    half = x.shape[1] // 2
    logdet = torch.zeros(x.shape[0], dtype=x.dtype, device=x.device)
    for fl in range(n_flows):
        p = f"sub_layers.{fl}."
        scale = sd[p + "norm_layer.scale"].reshape(1, -1)
        x = scale * (x + sd[p + "norm_layer.loc"].reshape(1, -1))       # modules.py:80
        logdet = logdet + torch.sum(torch.log(torch.abs(scale)))
        x = x * ((x >= 0).to(x) + (x < 0).to(x) * 0.9)                    # flow_blocks.py:180-182
        for i in (0, 1):                                                  # flow_blocks.py:82-95
            if i % 2 != 0:
                x = torch.cat((x[:, half:], x[:, :half]), dim=1)
            xa, xk = x[:, :half], x[:, half:]
            ci = cond if modes[fl] else torch.cat((xa, cond), dim=1)
            s = _mlp(sd, f"{p}coupling.s.{i}", ci, depth)
            t = _mlp(sd, f"{p}coupling.t.{i}", ci, depth)
            x = torch.cat((xa, xk * s.exp() + t), dim=1)
            logdet = logdet + s.sum(dim=1)
        x = x[:, sd[p + "shuffle.forward_shuffle_idx"]]                 # flow_blocks.py:150-152
    return x, logdet


def embed_pos(pos, cond_size=10):
    """SupervisedTransformer.embed_pos (INN.py:49-57): three separate 10-way one-hots of (pos*10-1e-4).long(),
    concatenated (a negative bin wraps inside its own block, as the reference's indexing does)."""
    pos = pos.detach().cpu() * cond_size - 1e-4
    idx = pos.long()
    blocks = []
    for k in range(3):
        one_hot = torch.zeros(pos.shape[0], cond_size)
        one_hot[torch.arange(pos.shape[0]), idx[:, k]] = 1
        blocks.append(one_hot)
    return torch.cat(blocks, dim=1)


# ------------------------------------------------------------------------------------ embedder
def _norm2d(sd, prefix, x, norm):
    if norm == "in":      # InstanceNorm2d(affine=False, no running stats), eps 1e-5
        return F.instance_norm(x, eps=1e-5)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.0, 1e-5)


def embedder_mean(sd, x, norm="in"):
    """ResnetEncoder.encode(x).mode() (AE.py:126-141,163-166; distributions.py:9,41): torchvision
    resnet50 v1.5 trunk (stride on the 3x3), adaptive avg-pool to 1x1, 1x1 conv to 2*zc, keep the
    first zc channels.  The ImageNet normalisation built at AE.py:111-114 is never applied."""
    h = F.conv2d(x, sd["model.conv1.weight"], None, stride=2, padding=3)
    h = F.relu(_norm2d(sd, "model.bn1", h, norm))
    h = F.max_pool2d(h, 3, 2, 1)
    for li, nb in enumerate((3, 4, 6, 3)):
        for bi in range(nb):
            p = f"model.layer{li + 1}.{bi}."
            stride = 2 if (li > 0 and bi == 0) else 1
            idt = h
            o = F.relu(_norm2d(sd, p + "bn1", F.conv2d(h, sd[p + "conv1.weight"]), norm))
            o = F.relu(_norm2d(sd, p + "bn2", F.conv2d(o, sd[p + "conv2.weight"], None, stride, 1), norm))
            o = _norm2d(sd, p + "bn3", F.conv2d(o, sd[p + "conv3.weight"]), norm)
            if bi == 0:
                idt = _norm2d(sd, p + "downsample.1",
                              F.conv2d(h, sd[p + "downsample.0.weight"], None, stride), norm)
            h = F.relu(o + idt)
    h = h.mean(dim=(2, 3), keepdim=True)
    enc = F.conv2d(h, sd["model.fc.sub_layers.0.weight"], sd["model.fc.sub_layers.0.bias"])
    zc = enc.shape[1] // 2
    return enc[:, :zc].reshape(x.shape[0], -1)


# ------------------------------------------------------------------------------------- decoder
def spectral_weight(sd, prefix):
    """Legacy ``torch.nn.utils.spectral_norm`` in eval mode: W = W_orig / (u . (W_mat v)), no power
    iteration (decoder.py:20-25).  Plain ``.weight`` when the conv is not wrapped."""
    if prefix + ".weight_orig" in sd:
        w = sd[prefix + ".weight_orig"]
        sigma = torch.dot(sd[prefix + ".weight_u"], w.reshape(w.shape[0], -1) @ sd[prefix + ".weight_v"])
        return w / sigma
    return sd[prefix + ".weight"]


def _spade(sd, p, x, img):
    """Spade.forward (normalization_layer.py:18-24)."""
    c = x.shape[1]
    groups = 16
    while c % groups != 0:
        groups -= 1
    nrm = F.group_norm(x, groups, eps=1e-5)
    y = F.interpolate(img, size=x.shape[-2:], mode="bilinear", align_corners=True)
    y = F.leaky_relu(F.conv2d(y, sd[p + ".conv.weight"], sd[p + ".conv.bias"], 1, 1), 0.2)
    gamma = F.conv2d(y, sd[p + ".conv_gamma.weight"], sd[p + ".conv_gamma.bias"], 1, 1).unsqueeze(2)
    beta = F.conv2d(y, sd[p + ".conv_beta.weight"], sd[p + ".conv_beta.bias"], 1, 1).unsqueeze(2)
    return nrm * (1 + gamma) + beta


def _adain(sd, p, x, z):
    """ADAIN.forward (normalization_layer.py:47-51): InstanceNorm3d then gamma*x+beta."""
    c = x.shape[1]
    nrm = F.instance_norm(x, eps=1e-5)
    gb = F.linear(z, sd[p + ".linear.weight"], sd[p + ".linear.bias"])
    return gb[:, :c].reshape(-1, c, 1, 1, 1) * nrm + gb[:, c:].reshape(-1, c, 1, 1, 1)


def _gen_block(sd, name, x, z, img):
    """GeneratorBlock.forward (decoder.py:33-49)."""
    if name + ".conv_s.weight_orig" in sd or name + ".conv_s.weight" in sd:
        xs = F.group_norm(x, 16, sd[name + ".norm_s.bn.weight"], sd[name + ".norm_s.bn.bias"], 1e-5)
        xs = F.conv3d(xs, spectral_weight(sd, name + ".conv_s"))
    else:
        xs = x
    dx = F.conv3d(F.leaky_relu(_spade(sd, name + ".norm_0", x, img), 0.2),
                  spectral_weight(sd, name + ".conv_0"), sd[name + ".conv_0.bias"], 1, 1)
    dx = F.conv3d(F.leaky_relu(_adain(sd, name + ".norm_1", dx, z), 0.2),
                  spectral_weight(sd, name + ".conv_1"), sd[name + ".conv_1.bias"], 1, 1)
    return xs + dx


def decoder_forward(sd, img, z, upsample_s, upsample_t, trace=None):
    """Generator.forward (decoder.py:97-120) -> (B, 16, 3, H, W)."""
    b = img.shape[0]
    x = F.linear(z, sd["fc.weight"], sd["fc.bias"]).reshape(b, -1, 1, 4, 4)
    x = _gen_block(sd, "head_0", x, z, img)
    if trace is not None:
        trace["head_0"] = x
    scales = [(2, 2, 2), (2, 2, 2), (2, 2, 2), (upsample_t[0], upsample_s[0], upsample_s[0]),
              (upsample_t[1], upsample_s[1], upsample_s[1])]
    for i, sc in enumerate(scales):
        x = F.interpolate(x, scale_factor=tuple(float(s) for s in sc))   # nearest (decoder.py:102-114)
        x = _gen_block(sd, f"g_{i}", x, z, img)
        if trace is not None:
            trace[f"g_{i}"] = x
    x = F.conv3d(F.leaky_relu(x, 0.2), sd["conv_img.weight"], sd["conv_img.bias"], 1, 1)
    return torch.tanh(x).transpose(1, 2)


# ---------------------------------------------------------------------------------- 3-D encoder
def _encoder3d_trunk(sd, x, stride_s, stride_t, layers=(2, 2, 2, 2)):
    """conv1 -> GN -> ReLU -> 4 stages x 2 BasicBlocks -> squeeze(T)  (resnet3D.py:208-219, :101-135)."""
    if x.shape[1] > x.shape[2]:
        x = x.transpose(1, 2)
    h = F.conv3d(x, sd["conv1.weight"], None, (2, 2, 2), (1, 3, 3))
    h = F.relu(F.group_norm(h, 16, sd["norm1.weight"], sd["norm1.bias"], 1e-5))
    for li in range(len(stride_s)):
        for bi in range(layers[li]):
            p = f"layer.{li}.{bi}."
            st = (stride_t[li], stride_s[li], stride_s[li]) if bi == 0 else (1, 1, 1)
            o = F.conv3d(h, sd[p + "conv1.weight"], None, st, 1)
            o = F.relu(F.group_norm(o, 16, sd[p + "bn1.weight"], sd[p + "bn1.bias"], 1e-5))
            o = F.conv3d(o, sd[p + "conv2.weight"], None, 1, 1)
            o = F.group_norm(o, 16, sd[p + "bn2.weight"], sd[p + "bn2.bias"], 1e-5)
            res = h
            if p + "downsample.0.weight" in sd:
                res = F.conv3d(h, sd[p + "downsample.0.weight"], None, st, 1)
                res = F.group_norm(res, 16, sd[p + "downsample.1.weight"], sd[p + "downsample.1.bias"], 1e-5)
            h = F.relu(o + res)
    return h.squeeze(2)


def encoder3d_mu(sd, x, stride_s, stride_t, layers=(2, 2, 2, 2)):
    """Encoder.forward (resnet3D.py:208-219) on (B, 3, T, H, W); returns mu only -- the sample and
    logvar are not consumed by the transfer path (get_model.py:87)."""
    h = _encoder3d_trunk(sd, x, stride_s, stride_t, layers)
    return F.conv2d(h, sd["conv_mu.weight"], sd["conv_mu.bias"]).reshape(h.shape[0], -1)


def encoder3d_posterior(sd, x, stride_s, stride_t, layers=(2, 2, 2, 2)):
    """Encoder.forward with its reparameterisation (resnet3D.py:202-206): (mu + exp(logvar/2) * eps, mu, logvar),
    eps drawn from the CPU generator exactly like ``torch.FloatTensor(size).normal_()`` does."""
    h = _encoder3d_trunk(sd, x, stride_s, stride_t, layers)
    mu = F.conv2d(h, sd["conv_mu.weight"], sd["conv_mu.bias"]).reshape(h.shape[0], -1)
    logvar = F.conv2d(h, sd["conv_var.weight"], sd["conv_var.bias"]).reshape(h.shape[0], -1)
    eps = torch.FloatTensor(logvar.size()).normal_().to(logvar)
    return eps.mul(logvar.mul(0.5).exp()).add(mu), mu, logvar


def flow_nll(gauss, logdet):
    """FlowLoss.forward (stage2_cINN/modules/loss.py:9-28) without logging: mean(0.5*|gauss|^2) - mean(logdet)."""
    g = gauss.reshape(gauss.shape[0], -1)
    return torch.mean(0.5 * torch.sum(g * g, dim=1)) - torch.mean(logdet)


# -------------------------------------------------------------------------------------- facade
class OracleModel:
    """Functional twin of get_model.Model (get_model.py:10-103) built from the same files."""

    def __init__(self, model_path, vid_length, transfer=False, load_yaml=None):
        import yaml

        def _ld(p):
            with open(p) as f:
                return yaml.safe_load(f)

        opt = _ld(model_path + "config_stage2.yaml")
        fs = opt["First_stage_model"]
        p1 = fs["model_path"] + fs["model_name"] + "/"
        c1 = _ld(p1 + "config_stage1.yaml")
        cm = opt["Conditioning_Model"]
        pae = cm["model_path"] + cm["model_name"] + "/"
        cae = _ld(pae + "config_stage2_AE.yaml")
        ld = lambda p: torch.load(p, map_location="cpu")["state_dict"]
        self.dec = ld(p1 + fs["checkpoint_decoder"] + ".pth")
        self.enc = ld(p1 + fs["checkpoint_encoder"] + ".pth.tar") if transfer else None
        self.flow = ld(model_path + "cINN.pth")
        self.emb = ld(pae + cm["checkpoint_name"] + ".pth")
        self.opt, self.c1, self.cae = opt, c1, cae
        self.control = bool((opt.get("Training") or {}).get("control"))
        self.n_flows = opt["Flow"]["n_flows"]
        self.depth = opt["Flow"]["flow_hidden_depth"]
        self.z_dim = c1["Decoder"]["z_dim"]
        self.vid_length = vid_length

    def to(self, device=None, dtype=None):
        """Move (and/or cast) every floating-point state-dict tensor: ``to(dtype=torch.float64)`` gives the fp64
        evaluation of the same algorithm (the "truth" the fp32 reference and the CUDA path are both measured against in
        tests/test_parity_truth_gpu.py), ``to('cuda')`` the eager PyTorch-on-GPU leg of bench.py."""
        def mv(sd):
            if sd is None:
                return None
            return {k: (v.to(device=device, dtype=dtype if v.is_floating_point() else None)) for k, v in sd.items()}
        self.dec, self.enc, self.flow, self.emb = mv(self.dec), mv(self.enc), mv(self.flow), mv(self.emb)
        return self

    def embed(self, x_0, cond=None):
        e = embedder_mean(self.emb, x_0, self.cae["AE"]["norm"])
        if self.control:
            e = torch.cat((e, embed_pos(cond).to(e)), dim=1)
        return e

    def decode(self, x_0, z, trace=None):
        d = self.c1["Decoder"]
        return decoder_forward(self.dec, x_0, z, d["upsample_s"], d["upsample_t"], trace)

    def _extend(self, x_0, z):
        seq = self.decode(x_0, z)
        while seq.shape[1] < self.vid_length:                              # get_model.py:71-73
            seq = torch.cat((seq, self.decode(seq[:, -1], z)), dim=1)
        return seq

    @torch.no_grad()
    def forward(self, x_0, residual, cond=None, return_latent=False, batch_slice=True):
        """get_model.py:51-75 with the residual passed in (the caller draws it on the CPU RNG, Q5)."""
        z = flow_reverse(self.flow, residual, self.embed(x_0, cond), self.n_flows, self.control, self.depth)
        seq = self._extend(x_0, z)
        if batch_slice:
            seq = seq[: self.vid_length]                                   # quirk Q1 (batch slice)
        return (seq, z) if return_latent else seq

    @torch.no_grad()
    def transfer(self, seq_query, x_0, return_latent=False):
        """get_model.py:77-103."""
        e = self.c1["Encoder"]
        mu = encoder3d_mu(self.enc, seq_query[:, 1:].transpose(1, 2), e["stride_s"], e["stride_t"])
        res, logdet = flow_forward(self.flow, mu, self.embed(seq_query[:, 0]), self.n_flows,
                                   self.control, self.depth)
        res = res.reshape(mu.shape[0], -1).repeat(x_0.shape[0], 1)
        z = flow_reverse(self.flow, res, self.embed(x_0), self.n_flows, self.control, self.depth)
        seq = self._extend(x_0, z)
        return (seq, z, mu, res, logdet) if return_latent else seq

    @torch.no_grad()
    def validation_step(self, seq, cond=None):
        """stage2_cINN/main.py:55-63: posterior sample -> forward flow on the first frame -> NLL.
        Returns (loss, gauss, logdet, post)."""
        e = self.c1["Encoder"]
        post, mu, logvar = encoder3d_posterior(self.enc, seq[:, 1:].transpose(1, 2), e["stride_s"], e["stride_t"])
        gauss, logdet = flow_forward(self.flow, post.reshape(post.shape[0], -1), self.embed(seq[:, 0], cond), self.n_flows,
                                     self.control, self.depth)
        return flow_nll(gauss, logdet), gauss, logdet, post

    @torch.no_grad()
    def reconstruct(self, seq):
        """utils/auxiliaries.py:73-75: decoder(seq[:, 0], Encoder(seq[:, 1:]).sample)."""
        e = self.c1["Encoder"]
        post, _, _ = encoder3d_posterior(self.enc, seq[:, 1:].transpose(1, 2), e["stride_s"], e["stride_t"])
        return self.decode(seq[:, 0], post)
