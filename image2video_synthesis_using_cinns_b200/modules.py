"""Host-side mirrors of the reference's module seams, backed by the C-ABI library.

Class names, attribute names and call signatures follow the reference so that code written against
it keeps working:

    ConditionalFlow.forward(x, embedding, reverse=False)      flow_blocks.py:31-57
    SupervisedTransformer.forward(input, cond, reverse, train) INN.py:59-73   (.flow, .embedder, .control)
    ResnetEncoder.encode(x).mode()                             AE.py:163-166, distributions.py:41
    Generator.forward(img, motion)                             decoder.py:97-120
    Encoder.forward(x) -> (sample, mu, logvar)                 resnet3D.py:208-219

They are thin: tensors stay torch CUDA tensors (device memory + stream plumbing only), the
arithmetic happens in libi2v_b200.so.  There is no CPU path: constructing any of these without a
CUDA device or without the built library raises.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import lib as _lib
from . import loader


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device=None):
    """The caller's current stream ON `device` (not on whatever device happens to be current)."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t, device):
    """fp32, contiguous, on `device` (the reference feeds fp32 NCHW tensors)."""
    if t.device != device or t.dtype != torch.float32:
        t = t.to(device=device, dtype=torch.float32)
    return t.contiguous()


class _Native:
    """Owns a C handle, its registered weight tensors and a grow-only workspace."""

    _destroy = None

    def __init__(self, device):
        if not torch.cuda.is_available():
            raise RuntimeError("image2video_synthesis_using_cinns_b200 needs a CUDA device (B200); "
                               "there is no CPU fallback for the sampling path")
        self.device = torch.device(device)
        self.L = _lib.load()
        self.h = None
        self._tensors = {}
        self._ws = None

    def _guard(self):
        """Every native call runs with this handle's device current: the library reads cudaGetDevice() for its
        per-device function attributes, and the stream handed over must belong to the same device as the pointers."""
        return torch.cuda.device(self.device)

    def _stream(self):
        return _stream(self.device)

    def _register(self, set_fn, tensors):
        for name, t in tensors.items():
            t = t.to(self.device).contiguous()
            self._tensors[name] = t
            _lib.check(set_fn(self.h, name.encode(), _ptr(t), t.numel() * t.element_size()), f"set_tensor({name})")

    def _workspace(self, nbytes):
        if nbytes == 0:
            raise RuntimeError(f"workspace query failed: {self.L.i2v_last_error().decode()}")
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    # ---- CUDA-graph replay of one native call (launch-bound small batches: the scripts' -bs 6 and B = 1) -------------
    # The path is ~280 kernel launches per pass, each with host-side argument checks and TMA descriptor encoding; at
    # B = 64 the device hides that, at B <= 6 the host does not keep up.  `_graphed` captures the launches of one call
    # on fixed-address buffers once per shape and replays them; inputs are copied in, the result copied out.
    graph = False
    graph_replays = 0          # replays so far
    graph_kernels = 0          # kernels replayed so far (they bypass i2v_launch_count)

    def _graphed(self, key, shapes_in, shape_out, call):
        """`call(*static_inputs, static_out, workspace_bytes_fn -> ws, stream)` is captured once per `key`."""
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        ent = self._graphs.get(key)
        if ent is None:
            ins = [torch.empty(sh, dtype=torch.float32, device=self.device) for sh in shapes_in]
            for t in ins:
                t.zero_()
            out = torch.empty(shape_out, dtype=torch.float32, device=self.device)
            holder = {}
            call(*ins, out, holder, self._stream())                  # eager warm-up: lazy initialisation happens here
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = self.L.i2v_launch_count()
            with torch.cuda.graph(g):
                call(*ins, out, holder, self._stream())
            ent = (g, ins, out, holder, self.L.i2v_launch_count() - n0)
            self._graphs[key] = ent
        return ent

    def weight_bytes(self):
        return sum(t.numel() * t.element_size() for t in self._tensors.values())

    def __del__(self):
        try:
            if self.h and self._destroy:
                getattr(self.L, self._destroy)(self.h)
                self.h = None
        except Exception:
            pass


class ConditionalFlow(_Native):
    """stage2_cINN ConditionalFlow (flow_blocks.py:8-60) as one persistent kernel per direction."""

    _destroy = "i2v_flow_destroy"

    def __init__(self, state_dict, in_channels, embedding_dim, hidden_dim, hidden_depth, n_flows, control=False,
                 device="cuda"):
        super().__init__(device)
        self.in_channels, self.cond_channels = in_channels, embedding_dim
        self.mid_channels, self.num_blocks, self.n_flows = hidden_dim, hidden_depth, n_flows
        tensors, cond_mode, self.zc_pad = loader.pack_flow(state_dict, n_flows, in_channels, embedding_dim, hidden_dim,
                                                           hidden_depth, control)
        cm = (ctypes.c_ubyte * n_flows)(*cond_mode)
        self.h = self.L.i2v_flow_create(n_flows, in_channels, self.zc_pad, hidden_dim, hidden_depth, cm)
        if not self.h:
            raise RuntimeError(self.L.i2v_last_error().decode())
        self._register(self.L.i2v_flow_set_tensor, tensors)

    def _cond(self, embedding):
        n = embedding.shape[0]
        if embedding.numel() != n * self.cond_channels:
            raise ValueError(f"embedding has {embedding.numel() // max(n, 1)} channels, flow expects {self.cond_channels}")
        e = _f32c(embedding.reshape(n, self.cond_channels), self.device)
        if self.zc_pad != e.shape[1]:
            e = torch.nn.functional.pad(e, (0, self.zc_pad - e.shape[1]))
        return e

    def forward(self, x, embedding, reverse=False):
        B = x.shape[0]
        if x.numel() != B * self.in_channels:
            raise ValueError(f"flow input has {x.numel() // max(B, 1)} channels, expected {self.in_channels}")
        x2 = _f32c(x.reshape(B, self.in_channels), self.device)
        if B == 0:
            out = x2.new_zeros(0, self.in_channels, 1, 1)
            return out if reverse else (out, x2.new_zeros(0))
        cond = self._cond(embedding)
        with self._guard():
            ws = self._workspace(self.L.i2v_flow_workspace_bytes(self.h, B))
            out = torch.empty_like(x2)
            if reverse:
                _lib.check(self.L.i2v_flow_reverse(self.h, _ptr(x2), _ptr(cond), _ptr(out), B, _ptr(ws), ws.numel(),
                                                   self._stream()), "flow_reverse")
                return out[:, :, None, None]
            logdet = torch.empty(B, dtype=torch.float32, device=self.device)
            _lib.check(self.L.i2v_flow_forward(self.h, _ptr(x2), _ptr(cond), _ptr(out), _ptr(logdet), B, _ptr(ws),
                                               ws.numel(), self._stream()), "flow_forward")
        return out[:, :, None, None], logdet

    __call__ = forward

    def reverse(self, out, xcond):
        return self.forward(out, xcond, reverse=True)


class _Mode:
    """What ``ResnetEncoder.encode`` hands back: only ``.mode()`` is used on this path (INN.py:62)."""

    def __init__(self, mean):
        self.mean = mean

    def mode(self):
        return self.mean


class ResnetEncoder(_Native):
    """Start-frame conditioning embedder (AE.py:91-166): ResNet-50 trunk, returns the posterior mean."""

    _destroy = "i2v_embedder_destroy"

    def __init__(self, state_dict, config, device="cuda", tc_mode=1, tc_min_ctas=None, graph=False):
        """``tc_mode``: 0 = fp32 SIMT convs only, 1 = tensor-core convs (fp32-grade fp16 split) where the GEMM fills
        the machine (InstanceNorm variant), 2 = wherever the shape is supported.  ``tc_min_ctas``: fewest 128x128
        output tiles for which mode 1 picks the tensor-core engine (library default 8)."""
        super().__init__(device)
        self.graph = bool(graph)
        self.config = config
        self.z_dim = config["z_dim"]
        norm = config["norm"]
        if norm not in ("in", "bn"):
            raise ValueError(f"embedder norm {norm!r} is not used by any reference config ('in' / 'bn')")
        if config["encoder_type"] != "resnet50":
            raise ValueError("only the resnet50 embedder of the reference configs is implemented")
        self.h = self.L.i2v_embedder_create(self.z_dim, 0 if norm == "in" else 1)
        if not self.h:
            raise RuntimeError(self.L.i2v_last_error().decode())
        self._register(self.L.i2v_embedder_set_tensor, loader.pack_embedder(state_dict, self.z_dim, norm, tensor_core=tc_mode != 0))
        self.set_tc_mode(tc_mode)
        if tc_min_ctas is not None:
            _lib.check(self.L.i2v_embedder_set_scalar(self.h, b"tc_min_ctas", float(tc_min_ctas)),
                       "embedder_set_scalar(tc_min_ctas)")

    def set_tc_mode(self, tc_mode):
        _lib.check(self.L.i2v_embedder_set_scalar(self.h, b"tc_mode", float(tc_mode)), "embedder_set_scalar(tc_mode)")
        self.tc_mode = int(tc_mode)

    def forward(self, x):
        x = _f32c(x, self.device)
        B, C, H, W = x.shape
        if C != 3:
            raise ValueError("embedder expects (B,3,H,W)")
        out = torch.empty(B, self.z_dim, dtype=torch.float32, device=self.device)
        if B == 0:
            return out
        with self._guard():
            if self.graph and not self.L.i2v_prof_is_enabled():
                def call(s_x, s_out, holder, stream):
                    if "ws" not in holder:
                        holder["ws"] = torch.empty(int(self.L.i2v_embedder_workspace_bytes(self.h, B, H, W)), dtype=torch.uint8,
                                                   device=self.device)
                    ws = holder["ws"]
                    _lib.check(self.L.i2v_embedder_forward(self.h, _ptr(s_x), _ptr(s_out), B, H, W, _ptr(ws), ws.numel(), stream),
                               "embedder_forward")
                g, (s_x,), s_out, _, nk = self._graphed(("emb", B, H, W, self.tc_mode), [(B, 3, H, W)], (B, self.z_dim), call)
                s_x.copy_(x)
                g.replay()
                out.copy_(s_out)
                self.graph_replays += 1
                self.graph_kernels += nk
                return out
            ws = self._workspace(self.L.i2v_embedder_workspace_bytes(self.h, B, H, W))
            _lib.check(self.L.i2v_embedder_forward(self.h, _ptr(x), _ptr(out), B, H, W, _ptr(ws), ws.numel(), self._stream()),
                       "embedder_forward")
        return out

    def encode(self, x):
        return _Mode(self.forward(x))


class Generator(_Native):
    """stage1_VAE 3-D conv decoder (decoder.py:55-120)."""

    _destroy = "i2v_decoder_destroy"

    def __init__(self, state_dict, dic, device="cuda", conv_engine=1, micro_batch=16, streams=1, graph=False):
        super().__init__(device)
        self.graph = bool(graph)
        self.nf, self.z_dim = dic["channel_factor"], dic["z_dim"]
        self.upsample_s, self.upsample_t = list(dic["upsample_s"]), list(dic["upsample_t"])
        self.micro_batch = micro_batch
        # micro-batches are independent: with streams > 1 they alternate over side streams (own workspace each), so one
        # micro-batch's HBM-bound passes and kernel tails overlap the other's tensor-bound convs
        self.n_streams = max(1, int(streams))
        self._side = None
        # optional hook ``f(row_lo, row_hi, out)`` called after each micro-batch has been enqueued on the caller's stream
        # (dist.py overlaps the all-gather of finished rows with the next micro-batch through it)
        self.on_micro_batch = None
        if conv_engine >= 1 and self.nf % 16 != 0:
            conv_engine = 0      # tensor-core tiles need channel counts that are multiples of 16
        us = (ctypes.c_int * 2)(*self.upsample_s)
        ut = (ctypes.c_int * 2)(*self.upsample_t)
        self.h = self.L.i2v_decoder_create(self.nf, self.z_dim, us, ut, conv_engine)
        if not self.h:
            raise RuntimeError(self.L.i2v_last_error().decode())
        tensors, scalars = loader.pack_decoder(state_dict, self.nf, conv_engine, self.upsample_t, self.upsample_s)
        self._register(self.L.i2v_decoder_set_tensor, tensors)
        for name, v in scalars.items():
            _lib.check(self.L.i2v_decoder_set_scalar(self.h, name.encode(), float(v)), f"set_scalar({name})")
        self.conv_engine = conv_engine
        self.frames = 8 * self.upsample_t[0] * self.upsample_t[1]
        self.size = 32 * self.upsample_s[0] * self.upsample_s[1]

    def forward(self, img, motion, out=None):
        img = _f32c(img, self.device)
        z = _f32c(motion.reshape(motion.shape[0], self.z_dim), self.device)
        B, _, H, W = img.shape
        if H != self.size or W != self.size:
            raise ValueError(f"decoder geometry renders {self.size}x{self.size} frames, start frame is {H}x{W}")
        if out is None:
            out = torch.empty(B, self.frames, 3, H, W, dtype=torch.float32, device=self.device)
        mb = max(1, min(self.micro_batch, B))
        if B == 0:
            return out
        with self._guard():
            if self.graph and self.n_streams <= 1 and not self.L.i2v_prof_is_enabled():
                return self._forward_graphed(img, z, out, B, H, W, mb)
            return self._forward_batches(img, z, out, B, H, W, mb)

    def _forward_graphed(self, img, z, out, B, H, W, mb):
        for b0 in range(0, B, mb):
            n = min(mb, B - b0)

            def call(s_img, s_z, s_out, holder, stream, n=n):
                if "ws" not in holder:
                    nbytes = self.L.i2v_decoder_workspace_bytes(self.h, n, H, W)
                    if nbytes == 0:
                        raise RuntimeError(f"workspace query failed: {self.L.i2v_last_error().decode()}")
                    holder["ws"] = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
                ws = holder["ws"]
                _lib.check(self.L.i2v_decoder_forward(self.h, _ptr(s_img), _ptr(s_z), _ptr(s_out), n, H, W, _ptr(ws), ws.numel(),
                                                      stream), "decoder_forward")

            g, (s_img, s_z), s_out, _, nk = self._graphed(("dec", n, H, W), [(n, 3, H, W), (n, self.z_dim)],
                                                         (n, self.frames, 3, H, W), call)
            s_img.copy_(img[b0:b0 + n])
            s_z.copy_(z[b0:b0 + n])
            g.replay()
            out[b0:b0 + n].copy_(s_out)
            self.graph_replays += 1
            self.graph_kernels += nk
            if self.on_micro_batch is not None:
                self.on_micro_batch(b0, b0 + n, out)
        return out

    def _forward_batches(self, img, z, out, B, H, W, mb):
        nbytes = self.L.i2v_decoder_workspace_bytes(self.h, mb, H, W)
        ns = min(self.n_streams, (B + mb - 1) // mb)
        if ns <= 1:
            ws = self._workspace(nbytes)
            for b0 in range(0, B, mb):
                n = min(mb, B - b0)
                _lib.check(self.L.i2v_decoder_forward(self.h, _ptr(img[b0:b0 + n]), _ptr(z[b0:b0 + n]), _ptr(out[b0:b0 + n]),
                                                      n, H, W, _ptr(ws), ws.numel(), self._stream()), "decoder_forward")
                if self.on_micro_batch is not None:
                    self.on_micro_batch(b0, b0 + n, out)
            return out
        if nbytes == 0:
            raise RuntimeError(f"workspace query failed: {self.L.i2v_last_error().decode()}")
        if self._side is None or len(self._side) < ns or self._side[0][1].numel() < nbytes:
            self._side = [(torch.cuda.Stream(device=self.device),
                           torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)) for _ in range(ns)]
        cur = torch.cuda.current_stream(self.device)
        for st, _ in self._side[:ns]:
            st.wait_stream(cur)                       # inputs were produced on the caller's stream
        for i, b0 in enumerate(range(0, B, mb)):
            n = min(mb, B - b0)
            st, ws = self._side[i % ns]
            _lib.check(self.L.i2v_decoder_forward(self.h, _ptr(img[b0:b0 + n]), _ptr(z[b0:b0 + n]), _ptr(out[b0:b0 + n]),
                                                  n, H, W, _ptr(ws), ws.numel(), ctypes.c_void_p(st.cuda_stream)),
                       "decoder_forward")
        for st, _ in self._side[:ns]:
            cur.wait_stream(st)                       # the caller's stream owns the frames again
        return out

    __call__ = forward


class Encoder(_Native):
    """stage1_VAE 3-D ResNet-18 video encoder (resnet3D.py:138-219), used by the transfer path."""

    _destroy = "i2v_encoder3d_destroy"

    def __init__(self, state_dict, dic, device="cuda", tc_mode=1, tc_min_ctas=None):
        super().__init__(device)
        if dic["res_type_encoder"] != "resnet18" or dic["use_max_pool"]:
            raise ValueError("only the resnet18 / no-max-pool encoder of the reference configs is implemented")
        self.z_dim = dic["z_dim"]
        ch = (ctypes.c_int * 5)(*dic["channels"])
        ss = (ctypes.c_int * 4)(*dic["stride_s"])
        st = (ctypes.c_int * 4)(*dic["stride_t"])
        self.h = self.L.i2v_encoder3d_create(ch, ss, st, self.z_dim)
        if not self.h:
            raise RuntimeError(self.L.i2v_last_error().decode())
        self._register(self.L.i2v_encoder3d_set_tensor, loader.pack_encoder3d(state_dict, tensor_core=tc_mode != 0))
        # tc_mode 0: fp32 SIMT convs only; 1: tensor-core engine for the stride-1 convs that fill the machine; 2: wherever supported
        if tc_min_ctas is not None:
            _lib.check(self.L.i2v_encoder3d_set_scalar(self.h, b"tc_min_ctas", float(tc_min_ctas)), "encoder3d_set_scalar(tc_min_ctas)")
        _lib.check(self.L.i2v_encoder3d_set_scalar(self.h, b"tc_mode", float(tc_mode)), "encoder3d_set_scalar(tc_mode)")

    def mu_logvar(self, x):
        # the reference accepts (B,3,T,H,W) and transposes (B,T,3,H,W) itself (resnet3D.py:209-210)
        if x.size(1) > x.size(2):
            x = x.transpose(1, 2)            # -> (B,3,T,H,W)
        seq = _f32c(x.transpose(1, 2), self.device)   # kernel layout (B,T,3,H,W)
        B, T, C, H, W = seq.shape
        if C != 3:
            raise ValueError("encoder expects RGB clips")
        out = torch.empty(B, 2 * self.z_dim, dtype=torch.float32, device=self.device)
        with self._guard():
            ws = self._workspace(self.L.i2v_encoder3d_workspace_bytes(self.h, B, T, H, W))
            _lib.check(self.L.i2v_encoder3d_forward(self.h, _ptr(seq), _ptr(out), B, T, H, W, _ptr(ws), ws.numel(),
                                                    self._stream()), "encoder3d_forward")
        return out[:, : self.z_dim], out[:, self.z_dim:]

    def forward(self, x):
        mu, logvar = self.mu_logvar(x)
        eps = torch.FloatTensor(logvar.size()).normal_().to(self.device)   # CPU RNG, resnet3D.py:204
        return eps.mul(logvar.mul(0.5).exp()).add(mu), mu, logvar

    __call__ = forward


class SupervisedTransformer:
    """INN.SupervisedTransformer (INN.py:8-73): embed the start frame, run the flow."""

    def __init__(self, flow, embedder, control=False):
        self.flow, self.embedder = flow, embedder
        self.control = bool(control)
        self.cond_size = 10 if self.control else 0

    def embed_pos(self, pos):
        # INN.py:49-57: three 10-way one-hots of floor(pos*10 - 1e-4), built on the host
        # Three separate (B, 10) one-hots like the reference, so a negative bin (pos <= -0.1: .long() truncates towards
        # zero) wraps INSIDE its own block exactly as the reference's negative indexing does, never into a neighbour's.
        pos = pos.detach().cpu() * self.cond_size - 1e-4
        idx = pos.long()
        rows = np.arange(pos.size(0))
        blocks = []
        for k in range(3):
            one_hot = torch.zeros(pos.size(0), self.cond_size)
            one_hot[rows, idx[:, k]] = 1
            blocks.append(one_hot)
        return torch.cat(blocks, dim=1).to(self.flow.device)

    def embed(self, cond):
        e = self.embedder.encode(cond[0]).mode().reshape(cond[0].size(0), self.embedder.z_dim)   # INN.py:62 (B, -1)
        if self.control:
            e = torch.cat((e, self.embed_pos(cond[1])), dim=1)
        return e

    def forward(self, input, cond, reverse=False, train=False):
        embed = self.embed(cond)
        if reverse:
            return self.flow(input, embed, reverse=True)
        return self.flow(input, embed)

    __call__ = forward

    def reverse(self, out, cond):
        return self.flow(out, cond, reverse=True)

    def eval(self):
        return self
