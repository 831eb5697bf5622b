"""Pin the oracle port (and the synthetic checkpoint layout) against the reference's own modules.
Runs only where /root/reference exists (the build container); skipped on the GPU box."""
import pytest
import torch

import oracle_torch as ot
import ref_harness as rh
from golden_util import rel_inf

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("dataset,kw", [
    ("bair", dict(nf=16, n_flows=3, enc_channels=[64, 32, 32, 64, 64], spade_gain=1.0)),
    ("dtdb_waterfall", dict(nf=16, n_flows=2, enc_channels=[64, 32, 32, 64, 64], spade_gain=0.5)),
])
def test_full_model_matches_reference(dataset, kw, ckpt_cache):
    mp = ckpt_cache(dataset=dataset, seed=3, **kw)
    ref = rh.build_reference_model(mp, 16, transfer=True)
    om = ot.OracleModel(mp, 16, transfer=True)
    img = om.opt["Data"]["img_size"]
    g = torch.Generator().manual_seed(5)
    x0 = torch.rand(2, 3, img, img, generator=g) * 2 - 1
    q = torch.rand(1, 16, 3, img, img, generator=g) * 2 - 1
    want = rh.run_reference_forward(ref, x0, seed=9)
    torch.manual_seed(9)
    got = om.forward(x0, torch.randn(2, om.z_dim))
    assert rel_inf(got, want) < 1e-6
    assert rel_inf(om.transfer(q, x0), rh.run_reference_transfer(ref, q, x0)) < 1e-6


def test_validation_step_and_reconstruction_match_reference(ckpt_cache):
    """posterior sample (CPU-RNG eps) -> forward flow -> NLL, and the stage-1 reconstruction."""
    mp = ckpt_cache(dataset="bair", seed=3, nf=16, n_flows=3, enc_channels=[64, 32, 32, 64, 64], spade_gain=1.0)
    ref = rh.build_reference_model(mp, 16, transfer=True)
    om = ot.OracleModel(mp, 16, transfer=True)
    img = om.opt["Data"]["img_size"]
    seq = torch.rand(2, 16, 3, img, img, generator=torch.Generator().manual_seed(11)) * 2 - 1
    w_loss, w_gauss, w_logdet, w_post = rh.run_reference_validation_step(ref, seq, seed=21)
    torch.manual_seed(21)
    loss, gauss, logdet, post = om.validation_step(seq)
    assert rel_inf(post, w_post) < 1e-6 and rel_inf(gauss.reshape(2, -1), w_gauss.reshape(2, -1)) < 1e-6
    assert rel_inf(logdet, w_logdet) < 1e-6 and abs(loss.item() - w_loss.item()) < 1e-5 * abs(w_loss.item())
    want = rh.run_reference_reconstruction(ref, seq, seed=22)
    torch.manual_seed(22)
    assert rel_inf(om.reconstruct(seq), want) < 1e-6


def test_synthetic_layout_matches_reference_constructors(ckpt_cache):
    """Key set, shapes and dtypes of every synthetic state-dict equal the reference constructors'
    (full-size BAIR geometry, flow with control to cover the 'cond' blocks)."""
    from image2video_synthesis_using_cinns_b200 import synthetic
    from image2video_synthesis_using_cinns_b200.config import DATASETS
    ns = rh.import_reference()
    gen = torch.Generator().manual_seed(0)

    def same(mine, theirs):
        assert set(mine) == set(theirs.keys()), set(mine) ^ set(theirs.keys())
        for k, v in theirs.items():
            assert tuple(mine[k].shape) == tuple(v.shape) and mine[k].dtype == v.dtype, k

    for control in (False, True):
        cc = 64 + (30 if control else 0)
        same(synthetic.flow_state_dict(gen, 64, cc, 512, 2, 5, control),
             ns.flow_blocks.ConditionalFlow(64, cc, 512, 2, 5, "None", control=control).state_dict())
    dcfg = {"channel_factor": 16, "z_dim": 64, "upsample_s": [2, 1], "upsample_t": [2, 1], "spectral_norm": True}
    same(synthetic.decoder_state_dict(gen, 16), ns.decoder.Generator(dcfg).state_dict())
    for ds in ("bair", "landscape", "dtdb_fire"):
        e = DATASETS[ds]["enc"]
        ecfg = dict(res_type_encoder="resnet18", deterministic=False, use_max_pool=False, z_dim=64, **e)
        same(synthetic.encoder3d_state_dict(gen, e["channels"], e["stride_s"]),
             ns.resnet3D.Encoder(ecfg).state_dict())
    for norm in ("in", "bn"):
        acfg = dict(deterministic=False, in_size=64, norm=norm, encoder_type="resnet50", z_dim=64)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            same(synthetic.embedder_state_dict(gen, 64, norm), ns.AE.ResnetEncoder(acfg).state_dict())


def test_embed_pos_matches_reference_including_negative_bins():
    ns = rh.import_reference()

    class _Self:
        cond_size = 10
    pos = torch.tensor([[0.05, 0.55, 0.999], [-0.25, 0.0, 1.0], [0.31, -0.95, 0.1]])
    with rh.cpu_cuda_identity():
        want = ns.INN.SupervisedTransformer.embed_pos(_Self(), pos.clone())
    assert torch.equal(ot.embed_pos(pos), want)
