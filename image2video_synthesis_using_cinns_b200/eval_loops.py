"""Throughput consumers of the sampling path: the *sampling* halves of the
reference's evaluation scripts and the stage-2 validation step, restated over the drop-in ``Model``.

    collect_synthesis_pairs(model, batches, dataset)   eval_synthesis_quality.py:39-58
    collect_realizations(model, batches, n_realiz)     eval_diversity.py:40-50
    FlowLoss / nll / flow_validation_step              stage2_cINN/modules/loss.py:9-28, stage2_cINN/main.py:55-63
    reconstruct_posterior(model, seq)                  utils/auxiliaries.py:66-84 (evaluate_FVD_posterior's loop body)
    sample_prior(model, seq, cond)                     utils/auxiliaries.py:87-101 (evaluate_FVD_prior's loop body)

The metric networks (I3D / VGG / LPIPS, FVD / FID) stay out of scope: these functions return
the tensors those metrics consume.  ``batches`` is any iterable of ``{"seq": (B, T, 3, H, W) [, "cond": (B, 3)]}``
dictionaries -- what the reference's data loaders yield.  Everything runs under ``torch.no_grad()`` on the
model's device; results are moved to the host per batch like the reference does.
"""
from __future__ import annotations

import torch


def _seq(file_dict, device):
    return file_dict["seq"].type(torch.FloatTensor).to(device)


@torch.no_grad()
def collect_synthesis_pairs(model, batches, dataset):
    """(fake, real) video stacks exactly as eval_synthesis_quality.py:39-58 assembles them for FVD / FID / LPIPS.

    bair : fake = [x_0, generated[:-1]]  real = seq[:, :-1]   (conditioning frame prepended, length kept)
    iPER : fake = [x_0, generated]       real = seq
    else : fake = generated              real = seq[:, :-1]   (dynamic textures)
    The loader is expected to deliver seq_length + 1 frames (eval_synthesis_quality.py:36).
    """
    seq_real, seq_fake = [], []
    for file_dict in batches:
        seq = _seq(file_dict, model.device)
        seq_gen = model(seq[:, 0])
        if dataset == "bair":
            seq_gen = torch.cat((seq[:, :1], seq_gen[:, :-1]), dim=1)
            seq_real.append(seq[:, :-1].cpu())
        elif dataset == "iPER":
            seq_gen = torch.cat((seq[:, :1], seq_gen), dim=1)
            seq_real.append(seq.cpu())
        else:
            seq_real.append(seq[:, :-1].cpu())
        seq_fake.append(seq_gen.cpu())
    fake, real = torch.cat(seq_fake, 0), torch.cat(seq_real, 0)
    if fake.shape != real.shape:                                   # eval_synthesis_quality.py:63 asserts the same
        raise ValueError(f"generated stack {tuple(fake.shape)} and real stack {tuple(real.shape)} differ: the loader "
                         "must deliver seq_length + 1 frames and batches no larger than seq_length (Model.forward "
                         "slices the batch, get_model.py:75)")
    return fake, real


@torch.no_grad()
def collect_realizations(model, batches, n_realiz):
    """(N, n_realiz, T, 3, H, W): n_realiz independent samples per start frame (eval_diversity.py:40-50).
    ``batches`` is iterated n_realiz times, so it must be re-iterable (a list or a DataLoader)."""
    per_realiz = []
    for _ in range(n_realiz):
        fakes = []
        for file_dict in batches:
            seq = _seq(file_dict, model.device)
            fakes.append(model(seq[:, 0]).cpu())
        per_realiz.append(torch.cat(fakes))
    return torch.stack(per_realiz, 1)


def nll(sample):
    """0.5 * sum(sample^2) over all but the batch dimension (loss.py:27-28)."""
    return 0.5 * torch.sum(torch.pow(sample, 2), dim=[1, 2, 3])


class FlowLoss:
    """stage2_cINN FlowLoss.forward (loss.py:9-25) without the wandb side effect: returns the loss and, when a
    ``logger`` with ``append`` is given, records the same four entries."""

    def forward(self, sample, logdet, logger=None, mode="eval"):
        nll_loss = torch.mean(nll(sample))
        assert len(logdet.shape) == 1
        nlogdet_loss = -torch.mean(logdet)
        loss = nll_loss + nlogdet_loss
        reference_nll_loss = torch.mean(nll(torch.randn_like(sample)))
        self.last = {"Loss": loss.item(), "reference_nll_loss": reference_nll_loss.item(),
                     "nlogdet_loss": nlogdet_loss.item(), "nll_loss": nll_loss.item()}
        if logger is not None:
            logger.append(self.last)
        return loss

    __call__ = forward


@torch.no_grad()
def flow_validation_step(model, seq, cond=None, loss_func=None, logger=None):
    """One iteration of the stage-2 validator (stage2_cINN/main.py:55-63): posterior sample of the 3-D encoder ->
    forward flow conditioned on the first frame -> negative log-likelihood.  Returns (loss, gauss, logdet)."""
    if not hasattr(model, "encoder"):
        raise RuntimeError("Model was built with transfer=False: the 3-D encoder is not loaded")
    seq = seq.type(torch.FloatTensor).to(model.device)
    post, mean, *_ = model.encoder(seq[:, 1:].transpose(1, 2))
    c = [seq[:, 0]] if not model.flow.control else [seq[:, 0], cond]
    gauss, logdet = model.flow(post.reshape(post.size(0), -1), c)
    loss = (loss_func or FlowLoss())(gauss, logdet, logger, mode="eval")
    return loss, gauss, logdet


@torch.no_grad()
def reconstruct_posterior(model, seq):
    """Stage-1 reconstruction ``decoder(seq[:, 0], Encoder(seq[:, 1:]).sample)`` (utils/auxiliaries.py:73-75).
    Returns (generated, original[:, 1:]) on the host."""
    if not hasattr(model, "encoder"):
        raise RuntimeError("Model was built with transfer=False: the 3-D encoder is not loaded")
    seq = seq.type(torch.FloatTensor).to(model.device)
    motion, *_ = model.encoder(seq[:, 1:].transpose(1, 2))
    return model.decoder(seq[:, 0], motion).cpu(), seq[:, 1:].cpu()


@torch.no_grad()
def sample_prior(model, seq, cond=None):
    """Stage-2 sampling step of evaluate_FVD_prior (utils/auxiliaries.py:93-98): residual on the CPU generator,
    inverse flow, ONE decoder pass (no autoregressive extension).  Returns (generated, original[:, 1:])."""
    seq = seq.type(torch.FloatTensor).to(model.device)
    res = torch.randn(seq.size(0), model.z_dim).to(model.device)
    c = [seq[:, 0]] if not model.flow.control else [seq[:, 0], cond]
    z = model.flow(res, c, reverse=True).view(seq.size(0), -1)
    return model.decoder(seq[:, 0], z).cpu(), seq[:, 1:].cpu()
