"""SURVEY 8 rows f3/f4 on the GPU: the stage-2 validation step (posterior sample -> forward flow -> NLL), the
stage-1 reconstruction and the evaluation sampling loops, against the oracle (pinned to the reference by
tests/test_oracle_vs_reference.py::test_validation_step_and_reconstruction_match_reference)."""
import pytest
import torch

import oracle_torch as ot
from golden_util import rel_inf, report

pytestmark = pytest.mark.gpu
TOL = 1e-4
CK = dict(dataset="landscape", seed=5, nf=16, n_flows=4, spade_gain=1.0, enc_channels=[64, 32, 32, 64, 64])


@pytest.fixture(scope="module")
def models(ckpt_cache):
    from image2video_synthesis_using_cinns_b200.get_model import Model
    mp = ckpt_cache(**CK)
    return Model(mp, 16, transfer=True), ot.OracleModel(mp, 16, transfer=True)


def test_validation_step_matches_oracle(models):
    from image2video_synthesis_using_cinns_b200 import eval_loops as el
    m, om = models
    seq = torch.rand(3, 16, 3, 128, 128, generator=torch.Generator().manual_seed(2)) * 2 - 1
    torch.manual_seed(31)
    w_loss, w_gauss, w_logdet, w_post = om.validation_step(seq)
    torch.manual_seed(31)                       # the posterior's eps comes from the CPU generator (resnet3D.py:204)
    log = []
    loss, gauss, logdet = el.flow_validation_step(m, seq, logger=log)
    e = dict(gauss=rel_inf(gauss.reshape(3, -1).cpu(), w_gauss.reshape(3, -1)), logdet=rel_inf(logdet.cpu(), w_logdet),
             loss=abs(loss.item() - w_loss.item()) / abs(w_loss.item()))
    report("eval_loops:validation_step", **e)
    assert gauss.shape == (3, 64, 1, 1) and logdet.shape == (3,)
    assert e["gauss"] < TOL and e["logdet"] < TOL and e["loss"] < TOL
    assert abs(log[0]["Loss"] - w_loss.item()) < TOL * abs(w_loss.item())


def test_reconstruction_matches_oracle(models):
    from image2video_synthesis_using_cinns_b200 import eval_loops as el
    m, om = models
    seq = torch.rand(2, 16, 3, 128, 128, generator=torch.Generator().manual_seed(3)) * 2 - 1
    torch.manual_seed(32)
    want = om.reconstruct(seq)
    torch.manual_seed(32)
    got, orig = el.reconstruct_posterior(m, seq)
    e = rel_inf(got, want)
    report("eval_loops:reconstruct", frames=e)
    assert got.shape == (2, 16, 3, 128, 128) and torch.equal(orig, seq[:, 1:])
    assert e < TOL


def test_sampling_loops(models):
    from image2video_synthesis_using_cinns_b200 import eval_loops as el
    m, om = models
    g = torch.Generator().manual_seed(4)
    batches = [{"seq": torch.rand(2, 17, 3, 128, 128, generator=g) * 2 - 1} for _ in range(2)]
    torch.manual_seed(33)
    fake, real = el.collect_synthesis_pairs(m, batches, "landscape")
    assert fake.shape == real.shape == (4, 16, 3, 128, 128)
    # batch 1 against the oracle with the residual the CPU generator hands out second
    torch.manual_seed(33)
    torch.randn(2, 64)
    want = om.forward(batches[1]["seq"][:, 0], torch.randn(2, 64))
    assert rel_inf(fake[2:], want) < TOL
    torch.manual_seed(34)
    r1 = el.collect_realizations(m, batches, 2)
    torch.manual_seed(34)
    r2 = el.collect_realizations(m, batches, 2)
    assert r1.shape == (4, 2, 16, 3, 128, 128) and torch.equal(r1, r2)        # deterministic given the seed
    assert (r1[:, 0] - r1[:, 1]).abs().max() > 1e-3                             # realisations differ
    torch.manual_seed(35)
    gen, orig = el.sample_prior(m, batches[0]["seq"])
    torch.manual_seed(35)
    want = om.forward(batches[0]["seq"][:, 0], torch.randn(2, 64))
    assert gen.shape == (2, 16, 3, 128, 128) and orig.shape[1] == 16 and rel_inf(gen, want) < TOL
