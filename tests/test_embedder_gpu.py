"""Start-frame embedder (AE.py:126-166): every conv engine setting against the oracle.

tc_mode 0 = fp32 SIMT convs, 1 = tensor-core convs where the GEMM fills the machine (the default), 2 = tensor-core convs
wherever the shape is supported (forces the small-plane / batch-overhang tiles of the per-tap kernel)."""
import pytest
import torch

import oracle_torch as ot
from golden_util import embed_tolerance, rel_inf, report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dataset,B", [("bair", 3), ("bair", 40), ("dtdb_fire", 2), ("landscape", 2)])
def test_embedder_engines_match_oracle(dataset, B, ckpt_cache):
    from image2video_synthesis_using_cinns_b200.get_model import Model
    mp = ckpt_cache(dataset=dataset, seed=33, nf=16, n_flows=2, with_encoder=False)
    m = Model(mp, 16, transfer=False)
    om = ot.OracleModel(mp, 16, transfer=False)
    img = m.config.Data["img_size"]
    g = torch.Generator().manual_seed(7)
    x0 = torch.rand(B, 3, img, img, generator=g) * 2 - 1
    want = om.embed(x0)
    tol = embed_tolerance(om, x0)
    errs = {}
    for mode in (0, 1, 2):
        m.flow.embedder.set_tc_mode(mode)
        got = m.flow.embedder.encode(x0.cuda()).mode().cpu()
        assert got.shape == want.shape
        errs[f"tc{mode}"] = rel_inf(got, want)
    report(f"embedder:{dataset}:B{B}", tol=tol, **errs)
    for k, e in errs.items():
        assert e < tol, (k, e, tol)


def test_embedder_rejects_unknown_option(ckpt_cache):
    from image2video_synthesis_using_cinns_b200.get_model import Model
    mp = ckpt_cache(dataset="bair", seed=33, nf=16, n_flows=2, with_encoder=False)
    m = Model(mp, 16, transfer=False)
    with pytest.raises(RuntimeError, match="unknown option"):
        m.flow.embedder.set_tc_mode(7)
