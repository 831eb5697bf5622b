"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the REAL reference.

Run in the build container (where /root/reference exists):

    python oracle/make_golden.py

For each case it writes synthetic checkpoints (deterministic from the seed, regenerated identically
by the tests through ``image2video_synthesis_using_cinns_b200.synthetic``), runs the unmodified
reference modules on CPU through ``oracle/ref_harness.py`` and stores inputs' seeds + the reference's
outputs.  The reference itself has no golden vectors for this path; these
fixtures are what pins both the oracle port and the CUDA path to the reference's arithmetic.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402
from image2video_synthesis_using_cinns_b200 import synthetic  # noqa: E402

# name -> (synthetic kwargs, batch, vid_length, transfer, frame stride kept in the fixture)
CASES = {
    "bair_small": dict(ck=dict(dataset="bair", seed=11, nf=16, n_flows=4, spade_gain=1.0,
                               enc_channels=[64, 32, 32, 64, 64]), B=2, vid_length=16, transfer=True,
                       keep=(1, 1)),
    "bair_small_refgain": dict(ck=dict(dataset="bair", seed=12, nf=16, n_flows=3, with_encoder=False),
                               B=1, vid_length=16, transfer=False, keep=(1, 1)),
    "landscape_small_bn": dict(ck=dict(dataset="landscape", seed=13, nf=16, n_flows=4, spade_gain=1.0,
                                       enc_channels=[64, 32, 32, 64, 64]), B=1, vid_length=24,
                               transfer=True, keep=(3, 2)),
    "bair_control": dict(ck=dict(dataset="bair", seed=14, nf=16, n_flows=8, control=True,
                                 with_encoder=False), B=3, vid_length=16, transfer=False, keep=(4, 2)),
}


def inputs_for(case, img_size):
    g = torch.Generator().manual_seed(1234)
    x0 = torch.rand(case["B"], 3, img_size, img_size, generator=g) * 2 - 1
    q = torch.rand(1, 16, 3, img_size, img_size, generator=g) * 2 - 1
    pos = torch.rand(case["B"], 3, generator=g)
    return x0, q, pos


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, case in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            mp = synthetic.write_synthetic_checkpoints(td, **case["ck"])
            m = rh.build_reference_model(mp, case["vid_length"], transfer=case["transfer"])
            img = m.config.Data["img_size"]
            x0, q, pos = inputs_for(case, img)
            control = bool(m.flow.control)
            kt, ks = case["keep"]
            rec = {}
            with rh.cpu_cuda_identity(), torch.no_grad():
                torch.manual_seed(77)
                residual = torch.randn(case["B"], m.z_dim)          # get_model.py:59 (CPU RNG)
                cond = [x0, pos if control else None]
                z = m.flow(residual, cond, reverse=True).view(case["B"], -1)
                embed = m.flow.embedder.encode(x0).mode().reshape(case["B"], -1)
                torch.manual_seed(77)
                frames = m(x0, pos if control else None)             # full Model.forward, same seed
                rec.update(residual=residual, embed=embed, z=z,
                           frames=frames[:, ::kt, :, ::ks, ::ks].contiguous(),
                           frames_shape=torch.tensor(frames.shape),
                           frames_sum=frames.double().sum(dim=(2, 3, 4)),
                           frames_sqsum=(frames.double() ** 2).sum(dim=(2, 3, 4)))
                if not control:
                    res_f, logdet = m.flow(z, [x0])                  # forward direction + logdet
                    rec.update(fwd_res=res_f.view(case["B"], -1), fwd_logdet=logdet)
                if case["transfer"]:
                    _, mu, _ = m.encoder(q[:, 1:].transpose(1, 2))
                    tr = m.transfer(q, x0)
                    res_q, ld_q = m.flow(mu, [q[:, 0]])
                    rec.update(t_mu=mu, t_res=res_q.view(1, -1), t_logdet=ld_q,
                               t_frames=tr[:, ::kt, :, ::ks, ::ks].contiguous(),
                               t_frames_shape=torch.tensor(tr.shape),
                               t_frames_sum=tr.double().sum(dim=(2, 3, 4)))
            arrays = {k: v.detach().cpu().numpy() for k, v in rec.items()}
            for k, v in arrays.items():
                assert np.isfinite(v).all(), f"{name}:{k} not finite -- synthetic weights ill-conditioned"
            arrays["meta"] = np.frombuffer(json.dumps(
                dict(case=name, ck=case["ck"], B=case["B"], vid_length=case["vid_length"],
                     transfer=case["transfer"], keep=list(case["keep"]), seed_inputs=1234,
                     seed_residual=77, torch=torch.__version__)).encode(), dtype=np.uint8)
            path = os.path.join(out_dir, name + ".npz")
            np.savez_compressed(path, **arrays)
            print(f"{name}: frames {tuple(frames.shape)} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
