"""The oracle port reproduces the outputs recorded from the REAL reference (tests/golden/*.npz,
made by oracle/make_golden.py).  CPU only."""
import pytest
import torch

import oracle_torch as ot
from golden_util import GOLDEN_CASES, golden_inputs, load_golden, rel_inf

TOL = 2e-6   # same arithmetic, same library; slack only for CPU ISA differences between hosts


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_fixture(name, ckpt_cache):
    meta, g = load_golden(name)
    mp = ckpt_cache(**meta["ck"])
    om = ot.OracleModel(mp, meta["vid_length"], transfer=meta["transfer"])
    img = om.opt["Data"]["img_size"]
    x0, q, pos = golden_inputs(meta, img)
    torch.manual_seed(meta["seed_residual"])
    residual = torch.randn(meta["B"], om.z_dim)
    assert torch.equal(residual, g["residual"]), "CPU RNG stream differs from the fixture's"
    cond = pos if om.control else None
    assert rel_inf(om.embed(x0, cond)[:, : g["embed"].shape[1]], g["embed"]) < TOL
    frames, z = om.forward(x0, residual, cond, return_latent=True)
    assert rel_inf(z, g["z"]) < TOL
    kt, ks = meta["keep"]
    assert list(frames.shape) == g["frames_shape"].tolist()
    assert rel_inf(frames[:, ::kt, :, ::ks, ::ks], g["frames"]) < TOL
    assert rel_inf(frames.double().sum(dim=(2, 3, 4)), g["frames_sum"]) < 1e-5
    if "fwd_res" in g:
        r, ld = ot.flow_forward(om.flow, z, om.embed(x0), om.n_flows, om.control, om.depth)
        assert rel_inf(r, g["fwd_res"]) < 1e-5 and rel_inf(ld, g["fwd_logdet"]) < 1e-5
        # invertibility: forward(reverse(residual)) == residual
        assert (r - residual).abs().max().item() < 1e-4
    if meta["transfer"]:
        tr, zt, mu, res, logdet = om.transfer(q, x0, return_latent=True)
        assert rel_inf(mu, g["t_mu"]) < TOL
        assert rel_inf(res[:1], g["t_res"]) < 1e-5 and rel_inf(logdet, g["t_logdet"]) < 1e-5
        assert list(tr.shape) == g["t_frames_shape"].tolist()
        assert rel_inf(tr[:, ::kt, :, ::ks, ::ks], g["t_frames"]) < 1e-5
