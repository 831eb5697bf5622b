"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the *unmodified* reference (CompVis/image2video-synthesis-using-cINNs) from
``/root/reference`` so that (a) the oracle port in ``oracle/oracle_torch.py`` can be pinned
against it and (b) golden vectors under ``tests/golden/`` can be generated
(``oracle/make_golden.py``).  The reference tree only exists in the build container, never on the
GPU box, so nothing that runs under ``-m gpu``, ``smoke()`` or ``bench.py`` may import this file.

Two non-invasive shims are needed, no reference file is touched or copied:

* ``omegaconf`` is not installed: a stub module exposing ``OmegaConf.load`` backed by PyYAML whose
  nodes answer ``None`` for missing keys (omegaconf 2.0 semantics the reference relies on,
  get_model.py:42 reads ``opt.Training['control']`` which only the BAIR config defines).
* the reference hard-codes ``.cuda()`` (get_model.py:22,28,42,59; INN.py:39,57;
  normalization_layer.py:20; resnet3D.py:204): on a CPU-only host ``Tensor.cuda`` /
  ``Module.cuda`` are patched to identity for the duration of the call.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import torch
import yaml

REFERENCE_ROOT = os.environ.get("I2V_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "get_model.py"))


class _Node(dict):
    """dict with attribute access; missing keys read as None (omegaconf 2.0.x behaviour)."""

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return self.get(k)

    def __getitem__(self, k):
        return self.get(k)


def _wrap(o):
    if isinstance(o, dict):
        return _Node({k: _wrap(v) for k, v in o.items()})
    if isinstance(o, list):
        return [_wrap(v) for v in o]
    return o


def _install_omegaconf_stub():
    if "omegaconf" in sys.modules:
        return
    m = types.ModuleType("omegaconf")

    class OmegaConf:  # noqa: D401 - tiny stub
        @staticmethod
        def load(path):
            with open(path) as f:
                return _wrap(yaml.safe_load(f))

    m.OmegaConf = OmegaConf
    sys.modules["omegaconf"] = m


@contextlib.contextmanager
def cpu_cuda_identity():
    """Patch ``.cuda()`` to identity when no GPU is present (reference hard-codes it)."""
    if torch.cuda.is_available():
        yield
        return
    t_cuda, m_cuda = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_cuda, m_cuda


def import_reference():
    """Return a namespace with the reference's hot-path modules (imported untouched)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_omegaconf_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import get_model as ref_get_model  # noqa: F401  (reference get_model.py)
        from stage1_VAE.modules import decoder as ref_decoder
        from stage1_VAE.modules import resnet3D as ref_resnet3d
        from stage2_cINN.AE.modules import AE as ref_ae
        from stage2_cINN.modules import INN as ref_inn
        from stage2_cINN.modules import flow_blocks as ref_flow_blocks
    ns = types.SimpleNamespace(
        get_model=ref_get_model,
        decoder=ref_decoder,
        resnet3D=ref_resnet3d,
        AE=ref_ae,
        INN=ref_inn,
        flow_blocks=ref_flow_blocks,
    )
    return ns


def build_reference_model(model_path: str, vid_length: int, transfer: bool = False):
    """``get_model.Model(model_path, vid_length, transfer)`` of the reference, on CPU if no GPU."""
    ns = import_reference()
    import warnings

    with cpu_cuda_identity(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = ns.get_model.Model(model_path, vid_length, transfer=transfer)
    return m


def run_reference_forward(model, x_0, seed: int, cond=None):
    """Reference ``Model.forward`` with the residual drawn exactly as get_model.py:59 (CPU RNG)."""
    with cpu_cuda_identity(), torch.no_grad():
        torch.manual_seed(seed)
        return model(x_0, cond)


def run_reference_transfer(model, seq_query, x_0):
    with cpu_cuda_identity(), torch.no_grad():
        return model.transfer(seq_query, x_0)


def run_reference_validation_step(model, seq, seed: int):
    """The stage-2 validator's loop body (stage2_cINN/main.py:55-63) on the reference's own modules, with the
    reference's FlowLoss arithmetic (stage2_cINN/modules/loss.py:9-28; the wandb logging is not part of the value).
    Returns (loss, gauss, logdet, post)."""
    with cpu_cuda_identity(), torch.no_grad():
        torch.manual_seed(seed)
        post, mean, *_ = model.encoder(seq[:, 1:].transpose(1, 2))
        gauss, logdet = model.flow(post.reshape(post.size(0), -1).detach(), [seq[:, 0]])
        nll = 0.5 * torch.sum(torch.pow(gauss, 2), dim=[1, 2, 3])
        loss = torch.mean(nll) - torch.mean(logdet)
        return loss, gauss, logdet, post


def run_reference_reconstruction(model, seq, seed: int):
    """evaluate_FVD_posterior's loop body (utils/auxiliaries.py:73-75) on the reference's own modules."""
    with cpu_cuda_identity(), torch.no_grad():
        torch.manual_seed(seed)
        motion, *_ = model.encoder(seq[:, 1:].transpose(1, 2))
        return model.decoder(seq[:, 0], motion)
