"""Host logic of eval_loops.py with a stub model: stack assembly per dataset, realisation
stacking, FlowLoss arithmetic.  No GPU, no native library."""
import pytest
import torch

from image2video_synthesis_using_cinns_b200 import eval_loops as el


class _StubModel:
    """Generates T frames that encode (call index, sample index) so that the assembly can be checked."""

    def __init__(self, T=16):
        self.device = torch.device("cpu")
        self.T, self.calls = T, 0

    def __call__(self, x0):
        self.calls += 1
        B, C, H, W = x0.shape
        out = x0[:, None].repeat(1, self.T, 1, 1, 1) + torch.arange(1, self.T + 1).view(1, -1, 1, 1, 1) + 100.0 * self.calls
        return out


def _batches(n, B, T):
    g = torch.Generator().manual_seed(0)
    return [{"seq": torch.rand(B, T, 3, 4, 4, generator=g).double()} for _ in range(n)]


@pytest.mark.parametrize("dataset,t_fake", [("bair", 16), ("iPER", 17), ("landscape", 16)])
def test_collect_synthesis_pairs_assembly(dataset, t_fake):
    bs = _batches(3, 2, 17)
    fake, real = el.collect_synthesis_pairs(_StubModel(), bs, dataset)
    assert fake.shape == real.shape == (6, t_fake, 3, 4, 4) and fake.dtype == torch.float32
    seq1 = bs[1]["seq"].float()
    gen1 = seq1[:, :1] + torch.arange(1, 17).view(1, -1, 1, 1, 1) + 200.0
    if dataset == "bair":      # eval_synthesis_quality.py:45-49
        assert torch.equal(fake[2:4, 0], seq1[:, 0]) and torch.equal(fake[2:4, 1:], gen1[:, :-1])
        assert torch.equal(real[2:4], seq1[:, :-1])
    elif dataset == "iPER":    # :50-54
        assert torch.equal(fake[2:4, 0], seq1[:, 0]) and torch.equal(fake[2:4, 1:], gen1)
        assert torch.equal(real[2:4], seq1)
    else:                      # :55-57
        assert torch.equal(fake[2:4], gen1) and torch.equal(real[2:4], seq1[:, :-1])


def test_collect_synthesis_pairs_rejects_wrong_loader_length():
    with pytest.raises(ValueError):
        el.collect_synthesis_pairs(_StubModel(), _batches(1, 2, 16), "bair")


def test_collect_realizations_layout():
    bs = _batches(2, 3, 16)
    m = _StubModel()
    out = el.collect_realizations(m, bs, n_realiz=4)
    assert out.shape == (6, 4, 16, 3, 4, 4) and m.calls == 8
    # realisation r, batch k came from call r*2 + k + 1  (eval_diversity.py:42-48: outer loop over realisations)
    assert torch.allclose(out[3:, 2, 0] - bs[1]["seq"].float()[:, 0], torch.full((3, 3, 4, 4), 1.0 + 100.0 * 6))


def test_flow_loss_matches_formula():
    g = torch.Generator().manual_seed(1)
    sample, logdet = torch.randn(5, 64, 1, 1, generator=g), torch.randn(5, generator=g)
    log = []
    loss = el.FlowLoss()(sample, logdet, log)
    want = (0.5 * sample.view(5, -1).pow(2).sum(1)).mean() - logdet.mean()
    assert torch.allclose(loss, want) and set(log[0]) == {"Loss", "reference_nll_loss", "nlogdet_loss", "nll_loss"}
    with pytest.raises(AssertionError):
        el.FlowLoss()(sample, logdet[:, None])
