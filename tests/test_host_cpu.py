"""CPU-side checks: the C-ABI library loads and exports every declared symbol, host logic, loaders."""
import os
import re

import pytest
import torch

from image2video_synthesis_using_cinns_b200 import lib, loader, synthetic
from image2video_synthesis_using_cinns_b200.config import load_yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    lib.build()
    return lib.load()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "i2v_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(i2v_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(lib.SIGNATURES), declared ^ set(lib.SIGNATURES)
    for name in declared:
        assert hasattr(built, name), name
    assert built.i2v_abi_version() == 1


def test_create_rejects_bad_geometry(built):
    assert not built.i2v_flow_create(20, 64, 63, 512, 2, None)       # zc not padded
    assert b"flow_create" in built.i2v_last_error()
    h = built.i2v_flow_create(20, 64, 64, 512, 2, None)
    assert h
    assert built.i2v_flow_workspace_bytes(h, 64) > 64 * 20 * 2048 * 4
    built.i2v_flow_destroy(h)


def test_no_cuda_device_fails_loudly(ckpt_cache):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from image2video_synthesis_using_cinns_b200.get_model import Model
    mp = ckpt_cache(dataset="bair", seed=4, nf=16, n_flows=2, with_encoder=False)
    with pytest.raises(RuntimeError, match="CUDA"):
        Model(mp, 16)


def test_yaml_missing_keys_read_none(tmp_path):
    p = tmp_path / "c.yaml"
    p.write_text("Training:\n  bs: 3\nData:\n  img_size: 64\n")
    c = load_yaml(str(p))
    assert c.Training["control"] is None and c.Training.bs == 3 and c.Data["img_size"] == 64
    assert c.Missing is None


def test_pack_flow_layout_roundtrip():
    """The packed flow tensors reproduce the MLPs: evaluate one coupling from the packed layout on the
    CPU and compare with the state-dict evaluation."""
    import oracle_torch as ot
    gen = torch.Generator().manual_seed(2)
    for control in (False, True):
        cc = 64 + (30 if control else 0)
        sd = synthetic.flow_state_dict(gen, 64, cc, 128, 2, 5, control)
        t, cond_mode, zc_pad = loader.pack_flow(sd, 5, 64, cc, 128, 2, control)
        assert zc_pad % 4 == 0 and cond_mode == [int(control and fl % 4 != 0) for fl in range(5)]
        x = torch.randn(3, 32, generator=gen)
        c = torch.randn(3, cc, generator=gen)
        cp = torch.nn.functional.pad(c, (0, zc_pad - cc))
        for fl in (0, 1, 4):
            for i in (0, 1):
                c1 = cp @ t["w1c"].reshape(5, 2, 256, zc_pad)[fl, i].t() + t["b1"].reshape(5, 2, 256)[fl, i]
                h = torch.nn.functional.leaky_relu(c1 + (0 if cond_mode[fl] else x @ t["w1x"][fl, i].t()), 0.01)
                for l in range(2):
                    hs = torch.cat([h[:, n * 128:(n + 1) * 128] @ t["wh"][fl, i, l, n].t() for n in (0, 1)], 1)
                    h = torch.nn.functional.leaky_relu(hs + t["bh"][fl, i, l], 0.01)
                st = torch.cat([h[:, n * 128:(n + 1) * 128] @ t["wo"][fl, i, n * 32:(n + 1) * 32].t() for n in (0, 1)], 1)
                st = st + t["bo"][fl, i]
                ci = c if cond_mode[fl] else torch.cat((x, c), 1)
                s = ot._mlp(sd, f"sub_layers.{fl}.coupling.s.{i}", ci)
                tt = ot._mlp(sd, f"sub_layers.{fl}.coupling.t.{i}", ci)
                assert torch.allclose(st[:, :32], s, atol=1e-5) and torch.allclose(st[:, 32:], tt, atol=1e-5)


def test_pack_decoder_folds_spectral_norm_and_permutes_fc():
    import oracle_torch as ot
    gen = torch.Generator().manual_seed(3)
    sd = synthetic.decoder_state_dict(gen, 16)
    t, _ = loader.pack_decoder(sd, 16)
    w = ot.spectral_weight(sd, "g_1.conv_0")
    assert torch.allclose(t["g_1.conv_0.w"], w.permute(2, 3, 4, 0, 1).reshape(27, w.shape[0], w.shape[1]), atol=1e-6)
    z = torch.randn(2, 64, generator=gen)
    ref = torch.nn.functional.linear(z, sd["fc.weight"], sd["fc.bias"]).reshape(2, 256, 1, 4, 4)
    mine = torch.nn.functional.linear(z, t["fc.w"], t["fc.b"]).reshape(2, 1, 4, 4, 256)
    assert torch.allclose(mine.permute(0, 4, 1, 2, 3), ref, atol=1e-6)


def test_pack_embedder_bn_fold():
    gen = torch.Generator().manual_seed(4)
    sd = synthetic.embedder_state_dict(gen, 64, "bn")
    t = loader.pack_embedder(sd, 64, "bn")
    x = torch.randn(1, 3, 16, 16, generator=gen)
    want = torch.nn.functional.batch_norm(
        torch.nn.functional.conv2d(x, sd["model.conv1.weight"], None, 2, 3), sd["model.bn1.running_mean"],
        sd["model.bn1.running_var"], sd["model.bn1.weight"], sd["model.bn1.bias"], False, 0.0, 1e-5)
    w = t["conv1.w"].reshape(7, 7, 64, 3).permute(2, 3, 0, 1)
    got = torch.nn.functional.conv2d(x, w, t["conv1.b"], 2, 3)
    assert torch.allclose(got, want, atol=1e-5)
    assert t["fc.w"].shape == (64, 2048)


def test_split_fp16_is_fp32_grade():
    gen = torch.Generator().manual_seed(5)
    w = torch.randn(27, 20, 64, generator=gen) * 0.03
    hi, lo, ws = loader.split_fp16(w, 16.0)
    assert hi.shape == (27, 32, 64) and hi[:, 20:].abs().max() == 0
    s = 1.0 / (16.0 * float(ws))
    rec = (hi[:, :20].double() + lo[:, :20].double()) / s
    assert (rec - w.double()).abs().max() / w.abs().max() < 2 ** -21
    assert hi.float().abs().max() < 2 ** 14.01 and torch.isfinite(hi.float()).all()
    sd = synthetic.decoder_state_dict(gen, 16)
    t, scalars = loader.pack_decoder(sd, 16, engine=1)
    assert "g_1.conv_0.wh" in t and "g_1.conv_0.w" not in t and "g_1.spade.sa" in scalars
    assert t["conv_img.wh"].shape == (27, 16, 16)


def test_pack_embedder_in_adds_tensor_core_split():
    gen = torch.Generator().manual_seed(6)
    sd = synthetic.embedder_state_dict(gen, 64, "in")
    t = loader.pack_embedder(sd, 64, "in")
    # stride-1 convs carry the split fp16 copy next to the fp32 weights; strided ones and the 3-channel stem do not
    assert "layer1.0.conv1.wh" in t and "layer1.0.ds.wh" in t and "layer2.1.conv2.wh" in t
    assert "conv1.wh" not in t and "layer2.0.conv2.wh" not in t and "layer2.0.ds.wh" not in t
    w = t["layer3.2.conv2.w"]
    hi, lo, ws = t["layer3.2.conv2.wh"], t["layer3.2.conv2.wl"], t["layer3.2.conv2.ws"]
    assert hi.shape == w.shape and hi.dtype == torch.float16
    s = 1.0 / (loader.ACT_SPLIT_SCALE * float(ws))
    assert ((hi.double() + lo.double()) / s - w.double()).abs().max() / w.abs().max() < 2 ** -21
    assert not any(k.endswith(".wh") for k in loader.pack_embedder(sd, 64, "in", tensor_core=False))
    # BatchNorm variant (round 2): the BN-folded stride-1 weights carry the split too, next to weight + bias
    tb = loader.pack_embedder(synthetic.embedder_state_dict(gen, 64, "bn"), 64, "bn")
    assert "layer1.0.conv1.wh" in tb and "layer1.0.conv1.b" in tb and "layer2.0.conv2.wh" not in tb
    wb, hb, lb, sb = tb["layer2.1.conv2.w"], tb["layer2.1.conv2.wh"], tb["layer2.1.conv2.wl"], tb["layer2.1.conv2.ws"]
    sb = 1.0 / (loader.ACT_SPLIT_SCALE * float(sb))
    assert ((hb.double() + lb.double()) / sb - wb.double()).abs().max() / wb.abs().max() < 2 ** -21


def test_embed_pos_wraps_inside_its_block_like_the_reference():
    """three separate 10-wide one-hots (INN.py:49-57): a negative bin wraps inside its own block."""
    import oracle_torch as ot
    from image2video_synthesis_using_cinns_b200.modules import SupervisedTransformer

    class _Flow:
        device = torch.device("cpu")

    st = SupervisedTransformer(_Flow(), None, control=True)
    pos = torch.tensor([[0.05, 0.55, 0.999], [-0.25, 0.0, 1.0], [0.31, -0.95, 0.1]])
    got = st.embed_pos(pos)
    assert got.shape == (3, 30) and torch.equal(got.sum(1), torch.full((3,), 3.0))
    idx = (pos * 10 - 1e-4).long()
    for b in range(3):
        for k in range(3):
            assert got[b, k * 10 + (idx[b, k].item() % 10)] == 1        # python's % = the reference's negative indexing
    assert torch.equal(got, ot.embed_pos(pos))


def test_pack_flow_chunks_is_the_per_cta_k_major_stream():
    """`wpack` of the cluster-resident flow kernel (csrc/flow_cluster.cu): per coupling and CTA rank the x-part of the first
    Linear, the hidden Linears and the last Linear, k-major, in consumption order."""
    gen = torch.Generator().manual_seed(8)
    n_flows, H, depth = 3, 256, 2
    sd = synthetic.flow_state_dict(gen, 64, 64, H, depth, n_flows, False)
    t, _, _ = loader.pack_flow(sd, n_flows, 64, 64, H, depth, False)
    wp = t["wpack"]
    Cc = H // 8
    assert wp.shape == (n_flows * 2, 16, 1 + depth * (H // 32) + 1, 32 * Cc)
    flat = wp.reshape(n_flows * 2, 16, -1)
    for c, r in ((0, 0), (3, 5), (5, 12)):
        fl, i, net, j = c // 2, c % 2, r // 8, r % 8
        cols = slice(j * Cc, (j + 1) * Cc)
        s = flat[c, r]
        l1 = s[: 32 * Cc].reshape(32, Cc)                                   # [k][col]
        assert torch.equal(l1, t["w1x"][fl, i, net * H:(net + 1) * H][cols].t())
        off = 32 * Cc
        for l in range(depth):
            hid = s[off: off + H * Cc].reshape(H, Cc)                       # all k-rows of the layer, chunk after chunk
            assert torch.equal(hid, t["wh"][fl, i, l, net][cols].t())
            off += H * Cc
        last = s[off: off + 4 * H].reshape(H, 4)
        assert torch.equal(last, t["wo"][fl, i, net * 32 + 4 * j: net * 32 + 4 * j + 4].t())
    # geometries the cluster kernel does not take keep the cooperative kernel's tensors only
    t128, _, _ = loader.pack_flow(synthetic.flow_state_dict(gen, 64, 64, 128, 2, 2, False), 2, 64, 64, 128, 2, False)
    assert "wpack" not in t128


def test_pack_encoder3d_adds_tensor_core_split():
    """Every 3x3x3 conv of the 3-D encoder's blocks carries the split fp16 copy (csrc/api.cu, encoder3d_run picks the engine per
    conv); the Cin = 3 stem does not; hi + lo reproduces the fp32 weights to fp32 grade."""
    from image2video_synthesis_using_cinns_b200.config import DATASETS
    gen = torch.Generator().manual_seed(8)
    e = DATASETS["bair"]["enc"]
    sd = synthetic.encoder3d_state_dict(gen, [64, 32, 32, 64, 64], e["stride_s"])
    t = loader.pack_encoder3d(sd)
    assert "conv1.wh" not in t
    for l in range(4):
        for b in range(2):
            for c in ("conv1", "conv2"):
                k = f"layer.{l}.{b}.{c}"
                w, hi, lo, ws = t[k + ".w"], t[k + ".wh"], t[k + ".wl"], t[k + ".ws"]
                assert hi.shape == w.shape and hi.dtype == torch.float16 and w.shape[0] == 27
                s = 1.0 / (loader.ACT_SPLIT_SCALE * float(ws))
                assert ((hi.double() + lo.double()) / s - w.double()).abs().max() / w.abs().max() < 2 ** -21
            assert (f"layer.{l}.{b}.ds.w" in t) == (f"layer.{l}.{b}.ds.wh" in t)
    assert any(k.endswith("ds.wh") for k in t)
    assert not any(k.endswith(".wh") for k in loader.pack_encoder3d(sd, tensor_core=False))


def test_pack_decoder_phase_weights_cover_every_temporally_upsampled_block():
    """conv_0 of g_0, g_1, g_2 (always x2 in time, decoder.py:102-108) and of g_3 / g_4 when the config upsamples them carries the
    phase-combined weights [2 phases x 2 taps x 9, Cout, Cin]: out[2j] = W0 a[j-1] + (W1+W2) a[j], out[2j+1] = (W0+W1) a[j] + W2 a[j+1]."""
    gen = torch.Generator().manual_seed(9)
    sd = synthetic.decoder_state_dict(gen, 16)
    t, _ = loader.pack_decoder(sd, 16, engine=1, upsample_t=(2, 1), upsample_s=(2, 1))
    assert {k.split(".")[0] for k in t if k.endswith("conv_0.wph")} == {"g_0", "g_1", "g_2", "g_3"}
    t0, _ = loader.pack_decoder(sd, 16)                       # fp32 taps of the same checkpoint
    w = t0["g_0.conv_0.w"].double().reshape(3, 9, *t0["g_0.conv_0.w"].shape[1:])
    want = torch.stack((w[0], w[1] + w[2], w[0] + w[1], w[2])).reshape(36, *w.shape[2:])
    hi, lo, ws = t["g_0.conv_0.wph"], t["g_0.conv_0.wpl"], t["g_0.conv_0.wps"]
    s = 1.0 / (loader.ACT_SPLIT_SCALE * float(ws))
    assert hi.shape == want.shape
    assert ((hi.double() + lo.double()) / s - want).abs().max() / want.abs().max() < 2 ** -21
    # a clip of identical frame pairs through the 3-tap conv == the 2-tap phase form on the half-rate clip
    a = torch.randn(1, w.shape[3], 3, 4, 4, generator=gen, dtype=torch.float64)
    w5 = t0["g_0.conv_0.w"].double().reshape(3, 3, 3, *w.shape[2:]).permute(3, 4, 0, 1, 2)
    full = torch.nn.functional.conv3d(a.repeat_interleave(2, dim=2), w5, padding=1)
    wp = want.reshape(2, 2, 3, 3, *w.shape[2:]).permute(0, 4, 5, 1, 2, 3)           # [phase][Cout][Cin][tap][kh][kw]
    ap = torch.nn.functional.pad(a, (1, 1, 1, 1, 1, 1))
    for p in range(2):
        got = torch.nn.functional.conv3d(ap[:, :, p:p + a.shape[2] + 1], wp[p])     # source planes j - 1 + p + q, q = 0, 1
        assert torch.allclose(got, full[:, :, p::2], atol=1e-10)


def test_bench_arguments_default_to_the_headline_config(monkeypatch):
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = bench.parse()
    assert (a.gpus, a.config, a.dataset, a.batch, a.seq_length, a.impl, a.legs) == (1, "bair_b64", "bair", 64, 16, "b200", 3)
    assert a.steps >= 1 and a.warmup >= 3 and not a.e2e_probe
    monkeypatch.setattr(sys, "argv", ["bench.py", "--config", "iper128_transfer_b64", "--legs", "1"])
    a = bench.parse()
    assert a.mode == "transfer" and a.dataset == "iper128" and a.legs == 1
