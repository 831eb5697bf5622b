"""cINN parity: the persistent flow kernel against the oracle port (flow_blocks.py semantics)."""
import pytest
import torch

import oracle_torch as ot
from golden_util import rel_inf
from image2video_synthesis_using_cinns_b200 import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-4   # ||a-b||_inf / ||b||_inf on z


def _mk(n_flows, zc, hidden, control, seed):
    from image2video_synthesis_using_cinns_b200.modules import ConditionalFlow
    gen = torch.Generator().manual_seed(seed)
    cc = zc + (30 if control else 0)
    sd = synthetic.flow_state_dict(gen, 64, cc, hidden, 2, n_flows, control)
    return sd, ConditionalFlow(sd, 64, cc, hidden, 2, n_flows, control=control), cc, gen


@pytest.mark.parametrize("B", [1, 3, 64, 70])
@pytest.mark.parametrize("n_flows,zc,hidden,control", [(20, 64, 512, False), (6, 128, 512, False), (8, 64, 256, True)])
def test_flow_reverse_forward_match_oracle(B, n_flows, zc, hidden, control):
    sd, flow, cc, gen = _mk(n_flows, zc, hidden, control, seed=B + n_flows)
    x = torch.randn(B, 64, generator=gen)
    cond = torch.randn(B, cc, generator=gen) * 0.7
    want = ot.flow_reverse(sd, x, cond, n_flows, control)
    got = flow(x.cuda(), cond.cuda(), reverse=True)
    assert got.shape == (B, 64, 1, 1)
    assert rel_inf(got.view(B, -1).cpu(), want) < TOL
    # forward direction + logdet on the oracle's latent
    wf, wld = ot.flow_forward(sd, want, cond, n_flows, control)
    gf, gld = flow(want.cuda(), cond.cuda())
    assert rel_inf(gf.view(B, -1).cpu(), wf) < TOL
    assert rel_inf(gld.cpu(), wld) < TOL
    # invertibility on the device: forward(reverse(x)) == x  (reference itself: 2-3e-6)
    rt, _ = flow(got, cond.cuda())
    assert (rt.view(B, -1).cpu() - x).abs().max().item() < 1e-4


def test_flow_rejects_uninitialised_actnorm():
    from image2video_synthesis_using_cinns_b200.modules import ConditionalFlow
    sd = synthetic.flow_state_dict(torch.Generator().manual_seed(0), 64, 64, 128, 2, 2)
    sd["sub_layers.1.norm_layer.initialized"] = torch.tensor(0, dtype=torch.uint8)
    with pytest.raises(ValueError, match="initialized"):
        ConditionalFlow(sd, 64, 64, 128, 2, 2)


def test_flow_empty_batch():
    sd, flow, cc, gen = _mk(2, 64, 128, False, 1)
    out = flow(torch.zeros(0, 64).cuda(), torch.zeros(0, 64).cuda(), reverse=True)
    assert out.shape == (0, 64, 1, 1)
