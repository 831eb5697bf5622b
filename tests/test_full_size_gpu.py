"""BASELINE.json-size checks (BAIR nf=64, 152.8 M-parameter decoder, 20-block flow with hidden 512).

Direct oracle comparison at B=2 (the CPU oracle needs ~8 s per sample at this size) plus size-independent
properties at the benchmark batch: engine-vs-engine agreement, determinism, flow invertibility, batch-split
invariance (micro-batching must not change any sample)."""
import pytest
import torch

import oracle_torch as ot
from golden_util import assert_parity, conditioned_tolerance, rel_inf, report, truth_model

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_ckpt(tmp_path_factory):
    from image2video_synthesis_using_cinns_b200 import synthetic
    d = tmp_path_factory.mktemp("full_bair")
    return synthetic.write_synthetic_checkpoints(str(d), "bair", seed=0, with_encoder=False)


def test_full_size_bair_matches_oracle(full_ckpt):
    from image2video_synthesis_using_cinns_b200.get_model import Model
    m = Model(full_ckpt, 16)
    om = ot.OracleModel(full_ckpt, 16)
    g = torch.Generator().manual_seed(8)
    x0 = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    z = torch.randn(2, 64, generator=g)
    residual = torch.randn(2, 64, generator=g)
    # decoder alone at full width (K up to 27648 per output)
    e_dec = rel_inf(m.decoder(x0.cuda(), z.cuda()).cpu(), om.decode(x0, z))
    wf, wz = om.forward(x0, residual, return_latent=True, batch_slice=False)
    gf, gz = m.sample(x0, residual=residual, return_latent=True)
    e_z, e_f = rel_inf(gz.cpu(), wz), rel_inf(gf.cpu(), wf)
    tol_f = conditioned_tolerance(lambda a: om.forward(a, residual, batch_slice=False), (x0,))
    # decoder at the flow's own latent (|z| is an order of magnitude larger than randn), both engines
    e_dec_z = rel_inf(m.decoder(x0.cuda(), wz.cuda()).cpu(), wf)
    m0 = Model(full_ckpt, 16, conv_engine=0, micro_batch=2)
    e_simt_z = rel_inf(m0.decoder(x0.cuda(), wz.cuda()).cpu(), wf)
    e_simt = rel_inf(m0.decoder(x0.cuda(), z.cuda()).cpu(), om.decode(x0, z))
    report("full_size:bair", decoder=e_dec, decoder_at_flow_z=e_dec_z, simt_decoder=e_simt, simt_decoder_at_flow_z=e_simt_z,
           z=e_z, frames=e_f, frames_tol=tol_f, zmax=float(wz.abs().max()))
    # at the flow's own latents (|z| ~ 75) the decoder itself is ill-conditioned: the fp32 SIMT engine, which only
    # differs from the oracle in summation order, already sits at ~8e-5.  Bar: the reference's own noise floor.
    assert e_dec < 1e-4 and e_simt < 1e-4 and e_z < 1e-4 and e_f < tol_f
    assert e_simt_z < 1e-4 and e_dec_z < 1e-4


def test_full_size_bair_against_fp64_truth(full_ckpt):
    """the BAIR frames gap (4.9e-4 against the fp32 oracle) is the reference's own fp32 noise.

    The oracle evaluated in float64 is the exact result of the reference's algorithm.  (a) Through the whole path the
    CUDA frames / embedding must be as close to it as the fp32 reference is (factor 1.5).  (b) With the 64x64
    InstanceNorm embedder's conditioning taken out -- the oracle's own fp32 embedding injected into the CUDA flow --
    the frames must meet the flat 1e-4 bar against the exact result."""
    from image2video_synthesis_using_cinns_b200.get_model import Model
    m = Model(full_ckpt, 16)
    om = ot.OracleModel(full_ckpt, 16)
    tm = truth_model(full_ckpt, 16)
    g = torch.Generator().manual_seed(8)
    B = 2
    x0 = torch.rand(B, 3, 64, 64, generator=g) * 2 - 1
    residual = torch.randn(B, 64, generator=g)
    # (a) whole path
    wf, wz = om.forward(x0, residual, return_latent=True, batch_slice=False)
    tf, tz = tm.forward(x0.double(), residual.double(), return_latent=True, batch_slice=False)
    gf, gz = m.sample(x0, residual=residual, return_latent=True)
    emb = m.flow.embedder.encode(x0.cuda()).mode().cpu()
    assert_parity("truth:bair_full:embedding", emb, om.embed(x0), tm.embed(x0.double()))
    assert_parity("truth:bair_full:z", gz.cpu(), wz, tz)
    assert_parity("truth:bair_full:frames", gf.cpu(), wf, tf, zmax=float(wz.abs().max()))
    # (b) oracle embedding injected: flow + decoder alone, flat bar against the exact result
    e32 = om.embed(x0)
    z_inj = m.flow.flow(residual.cuda(), e32.cuda(), reverse=True).view(B, -1)
    f_inj = m.decoder(x0.cuda(), z_inj)
    tz_inj = ot.flow_reverse(tm.flow, residual.double(), e32.double(), tm.n_flows, tm.control, tm.depth)
    tf_inj = tm.decode(x0.double(), tz_inj)
    wz_inj = ot.flow_reverse(om.flow, residual, e32, om.n_flows, om.control, om.depth)
    wf_inj = om.decode(x0, wz_inj)
    e_z, e_f = rel_inf(z_inj.cpu(), tz_inj), rel_inf(f_inj.cpu(), tf_inj)
    report("truth:bair_full:injected_embedding", z_vs_truth=e_z, frames_vs_truth=e_f, frames_vs_reference=rel_inf(f_inj.cpu(), wf_inj),
           reference_frames_vs_truth=rel_inf(wf_inj, tf_inj))
    assert e_z < 1e-4 and e_f < 1e-4


def test_full_size_properties_at_benchmark_batch(full_ckpt):
    from image2video_synthesis_using_cinns_b200.get_model import Model
    B = 64
    g = torch.Generator().manual_seed(9)
    x0 = (torch.rand(B, 3, 64, 64, generator=g) * 2 - 1).cuda()
    residual = torch.randn(B, 64, generator=g).cuda()
    m = Model(full_ckpt, 16, micro_batch=64)
    seq, z = m.sample(x0, residual=residual, return_latent=True)
    assert seq.shape == (B, 16, 3, 64, 64) and torch.isfinite(seq).all() and seq.abs().max() <= 1.0
    # determinism + micro-batch invariance: every sample is independent of its batch neighbours
    m16 = Model(full_ckpt, 16, micro_batch=16)
    seq16 = m16.sample(x0, residual=residual)
    assert torch.equal(seq16, m16.sample(x0, residual=residual))
    assert rel_inf(seq16.cpu(), seq.cpu()) < 2e-5
    # a different batch composition changes the embedder's split-K partition (1e-7-level rounding), which the
    # ill-conditioned 64x64 InstanceNorm embedder amplifies: the decoder and the flow are
    # exactly batch-invariant, the full path only up to the reference's own noise floor
    zs = m.sample(x0[5:9], residual=residual[5:9], return_latent=True)[1]
    assert rel_inf(m.decoder(x0[5:9], z[5:9]).cpu(), seq[5:9].cpu()) < 2e-5
    assert rel_inf(zs.cpu(), z[5:9].cpu()) < 1e-4
    # flow invertibility at B=64 (reference itself: 2-3e-6)
    back, logdet = m.flow(z, [x0])
    assert (back.view(B, -1) - residual).abs().max().item() < 1e-4 and torch.isfinite(logdet).all()
    # tensor-core engine vs the fp32 SIMT engine on the same weights
    m0 = Model(full_ckpt, 16, conv_engine=0, micro_batch=8)
    e = rel_inf(m.decoder(x0[:8], z[:8]).cpu(), m0.decoder(x0[:8], z[:8]).cpu())
    report("full_size:tc_vs_simt", decoder=e)
    assert e < 2e-4        # two engines, each within 1e-4 of the oracle at these latents (see the oracle test)
