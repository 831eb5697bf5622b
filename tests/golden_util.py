"""Helpers shared by the parity tests: golden fixture loading and the inputs they were made with."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["bair_small", "bair_small_refgain", "landscape_small_bn", "bair_control"]


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(d["meta"]).decode())
    arrays = {k: torch.from_numpy(d[k]) for k in d.files if k != "meta"}
    return meta, arrays


def golden_inputs(meta, img_size):
    """Same construction as oracle/make_golden.py::inputs_for."""
    g = torch.Generator().manual_seed(meta["seed_inputs"])
    x0 = torch.rand(meta["B"], 3, img_size, img_size, generator=g) * 2 - 1
    q = torch.rand(1, 16, 3, img_size, img_size, generator=g) * 2 - 1
    pos = torch.rand(meta["B"], 3, generator=g)
    return x0, q, pos


def rel_inf(a, b):
    """The parity metric of BASELINE.md section 5: ||a-b||_inf / ||b||_inf."""
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()
