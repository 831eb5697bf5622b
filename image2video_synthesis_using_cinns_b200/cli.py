"""Pre/post-processing of the reference's CLI scripts, restated without kornia / imageio / natsort
(generate_samples.py:36-62, generate_transfer.py:30-68, utils/auxiliaries.py:15-22,53-55)."""
from __future__ import annotations

import glob
import os
import re

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
