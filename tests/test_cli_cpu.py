"""Host side of the CLI surface (no GPU): flags, file discovery / natural order, GIF writer, the host formulas."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_flags_match_the_reference_scripts(monkeypatch):
    """generate_samples.py:11-19 / generate_transfer.py:12-18: same flags, same defaults."""
    monkeypatch.syspath_prepend(ROOT)
    import generate_samples
    import generate_transfer
    a = generate_samples.parse(["-gpu", "0", "-dataset", "DTDB", "-texture", "fire"])
    assert (a.gpu, a.dataset, a.texture, a.ckpt_path, a.seq_length, a.bs) == ("0", "DTDB", "fire", None, 16, 6)
    b = generate_transfer.parse(["-gpu", "1", "-dataset", "iPER", "-seq_length", "24"])
    assert (b.gpu, b.dataset, b.ckpt_path, b.seq_length) == ("1", "iPER", None, 24)
    with pytest.raises(SystemExit):
        generate_samples.parse(["-dataset", "bair"])          # -gpu is required, like the reference


def test_natural_order_and_listing(tmp_path):
    from image2video_synthesis_using_cinns_b200 import cli
    for n in ("f10.png", "f2.png", "f1.jpg", "notes.txt", "f3.jpeg"):
        (tmp_path / n).write_bytes(b"x")
    names = sorted((os.path.basename(p) for p in cli.list_images(str(tmp_path))), key=cli.natural_key)
    assert names == ["f1.jpg", "f2.png", "f3.jpeg", "f10.png"]


def test_host_formulas_and_gif_roundtrip(tmp_path):
    from PIL import Image
    from image2video_synthesis_using_cinns_b200 import cli
    seq = torch.rand(2, 3, 3, 8, 8, generator=torch.Generator().manual_seed(0)) * 2 - 1
    gif = cli.convert_seq2gif(seq)
    assert gif.shape == (3, 8, 16, 3) and abs(float(gif.max()) - 255.0) < 1e-3 and gif.min() >= 0
    # videos side by side along the width (utils/auxiliaries.py:18-20)
    d = cli.denorm(seq)
    assert np.allclose(gif[:, :, 8:], 255 * d[1].permute(0, 2, 3, 1).numpy() / float(d.max()), atol=1e-4)
    path = str(tmp_path / "out" / "r.gif")
    cli.save_gif(path, gif, fps=3)
    im = Image.open(path)
    assert im.n_frames == 3 and im.size == (16, 8)


def test_device_functions_refuse_to_run_without_cuda():
    from image2video_synthesis_using_cinns_b200 import cli
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="CUDA"):
        cli.preprocess_u8([np.zeros((4, 4, 3), np.uint8)], 8)
