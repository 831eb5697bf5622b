"""Pre/post-processing of the reference's CLI scripts, restated without kornia / imageio / natsort
(generate_samples.py:36-62, generate_transfer.py:30-68, utils/auxiliaries.py:15-22,53-55).

The arithmetic runs on the device: ``load_images`` uploads the decoded uint8 image and lets
``i2v_op_preprocess_u8`` do RGB order, /255, Normalize(0.5, 0.5) and the bilinear resize straight into a slot of
the ``x_0`` batch; ``convert_seq2gif_u8`` reduces the clip's maximum and writes the uint8 GIF canvas on the
device, so only 1 byte per colour sample crosses PCIe instead of a float.  The host only decodes / encodes image
files.  ``load_image`` / ``convert_seq2gif`` are the same formulas in plain torch: the checker of the tests."""
from __future__ import annotations

import glob
import os
import re

import ctypes

import numpy as np
import torch
import torch.nn.functional as F

IMG_SUFFIX = ("jpg", "png", "jpeg")


def natural_key(s):
    return [int(t) if t.isdigit() else t.lower() for t in re.split(r"(\d+)", s)]


def list_images(folder):
    out = []
    for suf in IMG_SUFFIX:
        out.extend(glob.glob(os.path.join(folder, f"*.{suf}")))
    return out


def load_image(path, size):
    """cv2.imread -> RGB -> [0,1] -> Normalize(0.5,0.5) -> Resize((size,size)) (bilinear, align_corners=False,
    no antialias: kornia 0.5's Resize is F.interpolate), as generate_samples.py:36-41."""
    import cv2
    bgr = cv2.imread(path)
    if bgr is None:
        raise FileNotFoundError(path)
    rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
    t = torch.from_numpy(rgb).permute(2, 0, 1).float().div(255.0)
    t = (t - 0.5) / 0.5
    return F.interpolate(t[None], size=(size, size), mode="bilinear", align_corners=False)[0]


def denorm(x):
    return ((x + 1) / 2).clamp(0, 1)                                   # utils/auxiliaries.py:53-55


def convert_seq2gif(sequence):
    """(N, T, 3, H, W) in [-1,1] -> (T, H, N*W, 3) float frames, videos side by side, scaled to 255/max
    (utils/auxiliaries.py:15-22)."""
    imgs = denorm(sequence).permute(0, 1, 3, 4, 2).detach().cpu().numpy()
    gif = np.concatenate(list(imgs), axis=2)
    return 255 * gif / np.max(gif)


def save_gif(path, frames, fps=3):
    """imageio.mimsave(path, frames.astype(uint8), fps=3) without imageio (Pillow)."""
    from PIL import Image
    frames = [Image.fromarray(f) for f in np.asarray(frames).astype(np.uint8)]
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    frames[0].save(path, save_all=True, append_images=frames[1:], duration=int(1000 / fps), loop=0)


# ------------------------------------------------------------------------------------------------ device versions
def _native(device):
    from . import lib
    if not torch.cuda.is_available():
        raise RuntimeError("image2video_synthesis_using_cinns_b200.cli: the device pre/post-processing needs a CUDA device")
    return lib, lib.load(), torch.device(device)


def read_image_u8(path):
    """cv2.imread: uint8 HWC in BGR order (generate_samples.py:39 converts it to RGB; the kernel does that)."""
    import cv2
    bgr = cv2.imread(path)
    if bgr is None:
        raise FileNotFoundError(path)
    return np.ascontiguousarray(bgr)


def preprocess_u8(images, size, device="cuda", bgr=True):
    """List of uint8 HWC arrays (any sizes) -> (N, 3, size, size) fp32 start frames in [-1, 1] on `device`."""
    lib, L, dev = _native(device)
    out = torch.empty(len(images), 3, size, size, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        keep = []
        for i, im in enumerate(images):
            if im.ndim != 3 or im.shape[2] != 3 or im.dtype != np.uint8:
                raise ValueError("preprocess_u8 expects uint8 HWC images with 3 channels")
            d = torch.from_numpy(im).to(dev, non_blocking=False)
            keep.append(d)
            lib.check(L.i2v_op_preprocess_u8(ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(out[i].data_ptr()), im.shape[0],
                                             im.shape[1], size, size, 1 if bgr else 0, stream), "preprocess_u8")
        torch.cuda.current_stream(dev).synchronize()      # the uploaded images may be freed now
    return out


def load_images(paths, size, device="cuda"):
    """generate_samples.py:36-42 for a list of files, arithmetic on the device."""
    return preprocess_u8([read_image_u8(p) for p in paths], size, device)


def frames_max(sequence):
    """Device scalar max(denorm(sequence)) (utils/auxiliaries.py:21)."""
    lib, L, dev = _native(sequence.device)
    seq = sequence.contiguous()
    mx = torch.empty(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        lib.check(L.i2v_op_frames_max(ctypes.c_void_p(seq.data_ptr()), ctypes.c_void_p(mx.data_ptr()), seq.numel(),
                                      ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "frames_max")
    return mx


def frames_to_u8(sequence, mx, layout="gif"):
    """(N, T, 3, H, W) device frames -> uint8 RGB: layout "gif" = the (T, H, N*W, 3) canvas of convert_seq2gif,
    "video" = (N, T, H, W, 3).  `mx` is the device scalar the clip is normalised by (frames_max, possibly
    all-reduced over ranks)."""
    lib, L, dev = _native(sequence.device)
    seq = sequence.contiguous()
    N, T, C, H, W = seq.shape
    if C != 3:
        raise ValueError("frames_to_u8 expects RGB frames (N, T, 3, H, W)")
    if layout == "gif":
        out = torch.empty(T, H, N * W, 3, dtype=torch.uint8, device=dev)
        sn, sh, st = 3 * W, 3 * N * W, 3 * H * N * W
    elif layout == "video":
        out = torch.empty(N, T, H, W, 3, dtype=torch.uint8, device=dev)
        sn, st, sh = 3 * T * H * W, 3 * H * W, 3 * W
    else:
        raise ValueError(f"unknown layout {layout!r}")
    with torch.cuda.device(dev):
        lib.check(L.i2v_op_frames_to_u8(ctypes.c_void_p(seq.data_ptr()), ctypes.c_void_p(mx.data_ptr()),
                                        ctypes.c_void_p(out.data_ptr()), N, T, H, W, sn, st, sh,
                                        ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "frames_to_u8")
    return out


def convert_seq2gif_u8(sequence):
    """utils/auxiliaries.py:15-22 + ``.astype(np.uint8)`` (generate_samples.py:61) on the device:
    (N, T, 3, H, W) in [-1, 1] -> uint8 (T, H, N*W, 3) host array."""
    return frames_to_u8(sequence, frames_max(sequence), "gif").cpu().numpy()
