"""SURVEY f1 on the device: start-frame preprocessing and the denorm -> uint8 GIF canvas, against the torch / numpy
formulas the reference scripts use (generate_samples.py:36-41,57-62; utils/auxiliaries.py:15-22,53-55), and the two CLI
scripts end to end on PNG files."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from golden_util import rel_inf

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_preprocess(rgb_u8, size):
    """kornia 0.5: image_to_tensor(rgb)/255 -> Normalize(0.5, 0.5) -> Resize == F.interpolate(bilinear, align_corners=False)."""
    t = torch.from_numpy(rgb_u8).permute(2, 0, 1).float().div(255.0)
    t = (t - 0.5) / 0.5
    return F.interpolate(t[None], size=(size, size), mode="bilinear", align_corners=False)[0]


@pytest.mark.parametrize("hw,size", [((64, 64), 64), ((100, 150), 64), ((37, 53), 128), ((256, 256), 128), ((480, 640), 64)])
def test_preprocess_u8_matches_torch_formula(hw, size):
    from image2video_synthesis_using_cinns_b200 import cli
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    bgr = rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8)
    got = cli.preprocess_u8([bgr, bgr[::-1].copy()], size)
    want = torch.stack([_ref_preprocess(np.ascontiguousarray(bgr[:, :, ::-1]), size),
                        _ref_preprocess(np.ascontiguousarray(bgr[::-1, :, ::-1]), size)])
    assert got.shape == (2, 3, size, size)
    assert (got.cpu() - want).abs().max().item() < 2e-6
    # cli.load_image (the host formula used as checker) is the same computation
    assert rel_inf(got.cpu(), want) < 2e-6


@pytest.mark.parametrize("shape", [(1, 2, 3, 8, 8), (3, 16, 3, 64, 64), (5, 4, 3, 32, 48)])
def test_frames_to_u8_is_bit_exact_with_numpy(shape):
    from image2video_synthesis_using_cinns_b200 import cli
    g = torch.Generator().manual_seed(sum(shape))
    seq = torch.rand(shape, generator=g) * 2.4 - 1.2            # beyond [-1, 1]: the clamp of denorm matters
    seq[0, 0, :, 0, 0] *= 0.3
    want = cli.convert_seq2gif(seq).astype(np.uint8)             # the reference formula in numpy (host)
    got = cli.convert_seq2gif_u8(seq.cuda())
    assert got.shape == want.shape and got.dtype == np.uint8
    assert np.array_equal(got, want)
    # a dim clip (max < 1) is stretched to 255 exactly like 255 * gif / np.max(gif)
    dim = seq * 0.25 - 0.5
    assert np.array_equal(cli.convert_seq2gif_u8(dim.cuda()), cli.convert_seq2gif(dim).astype(np.uint8))
    # per-video layout used by the uint8 all-gather
    mx = cli.frames_max(seq.cuda())
    vid = cli.frames_to_u8(seq.cuda(), mx, "video").cpu().numpy()
    N, T, _, H, W = shape
    assert np.array_equal(vid, want.reshape(T, H, N, W, 3).transpose(2, 0, 1, 3, 4))


def _write_pngs(folder, n, hw, seed):
    import cv2
    os.makedirs(folder, exist_ok=True)
    rng = np.random.default_rng(seed)
    for i in range(n):
        cv2.imwrite(os.path.join(folder, f"img{i}.png"), rng.integers(0, 256, size=(hw[0], hw[1], 3), dtype=np.uint8))


def test_generate_samples_cli_end_to_end(tmp_path, ckpt_cache, monkeypatch):
    """generate_samples.py on PNG start frames: same flags as the reference, GIF frames equal the oracle's rendering of
    the reference pipeline (host formulas) up to one grey level."""
    import oracle_torch as ot
    from image2video_synthesis_using_cinns_b200 import cli
    from PIL import Image
    monkeypatch.syspath_prepend(ROOT)
    import generate_samples
    mp = ckpt_cache(dataset="landscape", seed=5, nf=16, n_flows=3, spade_gain=1.0, with_encoder=False)
    img_dir, out_dir = str(tmp_path / "GT") + "/", str(tmp_path / "res") + "/"
    _write_pngs(img_dir, 3, (96, 160), 1)
    gpu = os.environ.get("CUDA_VISIBLE_DEVICES", "0")
    torch.manual_seed(17)
    videos = generate_samples.main(["-gpu", gpu, "-dataset", "landscape", "-ckpt_path", mp, "-bs", "2",
                                    "-img_path", img_dir, "-save_path", out_dir])
    assert videos.shape == (3, 16, 3, 128, 128)
    # the reference pipeline on the host: load -> Model.forward (batches of 2, CPU-RNG residual per batch) -> gif
    names = cli.list_images(img_dir)
    imgs = torch.stack([cli.load_image(n, 128) for n in names])
    om = ot.OracleModel(mp, 16)
    torch.manual_seed(17)
    want = torch.cat([om.forward(imgs[i:i + 2], torch.randn(imgs[i:i + 2].size(0), 64)) for i in range(0, 3, 2)])
    assert rel_inf(videos.cpu(), want) < 1e-4
    gif = Image.open(os.path.join(out_dir, "results.gif"))
    assert gif.n_frames == 16 and gif.size == (3 * 128, 128)
    canvas = cli.convert_seq2gif_u8(videos)
    ref_canvas = cli.convert_seq2gif(want).astype(np.uint8)
    assert canvas.shape == ref_canvas.shape == (16, 128, 3 * 128, 3)
    assert np.abs(canvas.astype(int) - ref_canvas.astype(int)).max() <= 1


def test_generate_transfer_cli_end_to_end(tmp_path, ckpt_cache, monkeypatch):
    import oracle_torch as ot
    from image2video_synthesis_using_cinns_b200 import cli
    from PIL import Image
    monkeypatch.syspath_prepend(ROOT)
    import generate_transfer
    mp = ckpt_cache(dataset="bair", seed=6, nf=16, n_flows=3, spade_gain=1.0, enc_channels=[64, 32, 32, 64, 64])
    root = str(tmp_path / "transfer") + "/"
    for v in range(2):
        _write_pngs(os.path.join(root, f"vid{v}"), 16, (64, 64), 10 + v)
    gpu = os.environ.get("CUDA_VISIBLE_DEVICES", "0")
    out_dir = str(tmp_path / "res") + "/"
    results = generate_transfer.main(["-gpu", gpu, "-dataset", "bair", "-ckpt_path", mp, "-img_path", root, "-save_path", out_dir])
    assert len(results) == 2 and results[0].shape == (3, 16, 3, 64, 64)      # query + one transfer per clip
    clips = torch.stack([torch.stack([cli.load_image(n, 64) for n in sorted(cli.list_images(os.path.join(root, f"vid{v}")),
                                                                             key=cli.natural_key)]) for v in range(2)])
    om = ot.OracleModel(mp, 16, transfer=True)
    want = om.transfer(clips[1][None], clips[:, 0])
    assert rel_inf(results[1][1:].cpu(), want) < 5e-4     # 64x64 InstanceNorm embedder: see golden_util.embed_tolerance
    assert rel_inf(results[1][:1].cpu(), clips[1][None]) < 2e-6
    assert Image.open(os.path.join(out_dir, "transfer_1.gif")).n_frames == 16
