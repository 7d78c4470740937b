"""Losses.  Mirrors lib/networks/losses.py: PointFlowNLL :7-20, GaussianFlowNLL :23-31,
GaussianEntropy :34-39, FlowMixtureNLL :81-137, Flow_Mixture_Loss :140-173."""
import math

import torch
import torch.nn as nn

_LOG_2PI = math.log(2.0 * math.pi)


class PointFlowNLL(nn.Module):
    def forward(self, output_decoder, **kwargs):
        z = output_decoder['p_prior_samples'][0]
        mu = output_decoder['p_prior_mus'][0]
        lv = output_decoder['p_prior_logvars'][0]
        logdet = sum(output_decoder['p_prior_logvars'])
        quad = (z - mu) ** 2 / torch.exp(lv)
        return 0.5 * ((logdet + quad).sum(dim=1, keepdim=True) + _LOG_2PI * z.shape[1])


class GaussianFlowNLL(nn.Module):
    def forward(self, samples, mus, logvars):
        z = samples[0]
        quad = (z - mus[0]) ** 2 / torch.exp(logvars[0])
        return 0.5 * (torch.sum(sum(logvars) + quad) / z.shape[0] + _LOG_2PI * z.shape[1])


class GaussianEntropy(nn.Module):
    def forward(self, logvars):
        return 0.5 * (logvars.shape[1] * (1.0 + _LOG_2PI) + logvars.sum(1).mean())


class FlowMixtureNLL(nn.Module):
    """mean over shapes of the summed per-point mixture NLL (losses.py:88-137).

    The fused decoder already holds the per-point NLL (log-sum-exp done in the CUDA kernel,
    key 'mixture_nll'); for list-style decoder outputs (fused_nll=False, or outputs of the
    reference's own modules) the same closed form is evaluated vectorised over B and K instead
    of the reference's B x K Python loop.
    """

    def forward(self, output_decoder, mixture_weights_logits):
        first = output_decoder[0]
        if 'mixture_nll' in first:
            return first['mixture_nll'].sum(dim=1).mean()
        logw = mixture_weights_logits - torch.logsumexp(mixture_weights_logits, dim=-1, keepdim=True)
        logps = []
        for od in output_decoder:
            z = od['p_prior_samples'][0]
            lv0 = od['p_prior_logvars'][0]
            quad = (z - od['p_prior_mus'][0]) ** 2 / torch.exp(lv0)
            logps.append(-0.5 * ((sum(od['p_prior_logvars']) + quad).sum(dim=1) + _LOG_2PI * z.shape[1]))
        logp = torch.stack(logps, dim=2) + logw.unsqueeze(1)
        return (-torch.logsumexp(logp, dim=-1)).sum(dim=1).mean()


class Flow_Mixture_Loss(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.pnll_weight = kwargs.get('pnll_weight')
        self.gnll_weight = kwargs.get('gnll_weight')
        self.gent_weight = kwargs.get('gent_weight')
        self.n_components = kwargs.get('n_components')
        self.PNLL = FlowMixtureNLL()
        self.GNLL = GaussianFlowNLL()
        self.GENT = GaussianEntropy()

    def forward(self, output_prior, output_decoder, mixture_weights_logits):
        pnll = self.PNLL(output_decoder, mixture_weights_logits)
        gnll = self.GNLL(output_prior['g_prior_samples'], output_prior['g_prior_mus'], output_prior['g_prior_logvars'])
        gent = self.GENT(output_prior['g_posterior_logvars'])
        return self.pnll_weight * pnll + self.gnll_weight * gnll - self.gent_weight * gent, pnll, gnll, gent
