"""a Model built for ``cuda:1`` must work while ``cuda:0`` is the current device -- native calls run
under the handle's device with that device's stream, and the >48 KB dynamic shared-memory attributes are set per device.
Needs two GPUs (skipped on the single-GPU box; run once with ``gpurun --gpus 2``)."""
import pytest
import torch

import oracle_torch as ot
from golden_util import rel_inf

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_second_device_while_first_is_current(ckpt_cache):
    from image2video_synthesis_using_cinns_b200.get_model import Model
    mp = ckpt_cache(dataset="landscape", seed=5, nf=16, n_flows=4, spade_gain=1.0, enc_channels=[64, 32, 32, 64, 64])
    torch.cuda.set_device(0)
    m0 = Model(mp, 16, transfer=True, device="cuda:0")
    m1 = Model(mp, 16, transfer=True, device="cuda:1")       # second device: every function attribute once more
    om = ot.OracleModel(mp, 16, transfer=True)
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(2, 3, 128, 128, generator=g) * 2 - 1
    q = torch.rand(1, 16, 3, 128, 128, generator=g) * 2 - 1
    residual = torch.randn(2, 64, generator=g)
    want = om.forward(x0, residual, batch_slice=False)
    assert torch.cuda.current_device() == 0
    got1 = m1.sample(x0, residual=residual)                   # inputs on the host, current device 0, model on device 1
    got0 = m0.sample(x0, residual=residual)
    assert got1.device == torch.device("cuda", 1) and got0.device == torch.device("cuda", 0)
    assert rel_inf(got1.cpu(), want) < 1e-4 and rel_inf(got0.cpu(), want) < 1e-4
    assert torch.equal(got0.cpu(), got1.cpu())                # same kernels, same arithmetic on both devices
    t1 = m1.transfer(q, x0)
    assert rel_inf(t1.cpu(), om.transfer(q, x0)) < 1e-4
    assert torch.cuda.current_device() == 0
