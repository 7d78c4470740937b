"""Mixture-of-flows model.  Mirrors lib/networks/flow_mixture.py: Flow_Mixture_Model :11-179,
Flow_Mixture_SVR_Model :181-230.

`decode` is the drop-in boundary of the hot path.  Training branch (reference :163-166, K calls of
one_flow_decode + the B x K Python loop of FlowMixtureNLL): all K coupling stacks of every point
run in the CUDA kernels in one pass and the K-way log-sum-exp is done on the device.  Evaluation
branch (:141-177, host multinomial + K gather/scatter decoder calls, B == 1): one sampling kernel
draws the component and the base noise per point with Philox and runs the direct stacks, for any
batch size.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from .decoders import LocalCondRNVPDecoder
from .encoders import FeatureEncoder, WeightsEncoder
from .models import Local_Cond_RNVP_MC_Global_RNVP_VAE
from ..flowstack import FlowStack, mixture_nll, sample_mixture


class Flow_Mixture_Model(Local_Cond_RNVP_MC_Global_RNVP_VAE):
    """K decoder flows + mixture weights.

    Extra (non-reference) attributes:
      fused_nll (bool, default True): training-mode `decode` returns the per-point mixture NLL
          computed in-kernel under key 'mixture_nll'; with False it returns the reference's
          list structure ('p_prior_samples'[0] = z_j, 'p_prior_logvars' summing to S_j) so any
          FlowMixtureNLL implementation can consume it.
      sample_seed (int | None): Philox seed of the sampling kernel; None draws one from torch's
          global generator on every call (so torch.manual_seed governs it).  With a fixed seed every
          `decode` call still gets its own Philox stream: the per-model call counter `sample_calls` is
          folded into the stream id, so the shape-by-shape evaluation loops of the reference
          (training.py:355, evaluating.py:93, B = 1 per call) do not reuse one set of draws for every
          shape; set `sample_calls = 0` to replay a sequence.
      nonfinite_points(): how many per-point NLLs the kernels have written as NaN/inf since the last
          call (device counter, one 4-byte read) -- the reference stops without an update on a NaN
          loss (training.py:43-46); the loss itself is NaN in that case here too.
    """

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.n_components = kwargs['n_components']
        self.params_reduce_mode = kwargs['params_reduce_mode']
        self.weights_type = kwargs['weights_type']
        self.mixture_weights_logits = nn.Parameter(torch.zeros(self.n_components), requires_grad=True)
        n_flows, n_features = self._get_decoder_params()
        self.pc_decoder = nn.ModuleList([
            LocalCondRNVPDecoder(n_flows, n_features, self.g_latent_space_size, weight_std=0.01)
            for _ in range(self.n_components)])
        self.mixture_weights_encoder = WeightsEncoder(3, self.g_latent_space_size, self.n_components,
                                                      deterministic=True, mu_weight_std=0.001, mu_bias=0.0,
                                                      logvar_weight_std=0.01, logvar_bias=0.0)
        self.fused_nll = True
        self.sample_seed = None
        self.sample_stream = 0
        self.sample_calls = 0
        self._stack = None

    # ------------------------------------------------------------------ sizing (flow_mixture.py:44-102)
    def _budget(self):
        return LocalCondRNVPDecoder.get_param_count(self.p_decoder_n_flows, self.p_decoder_n_features,
                                                    self.g_latent_space_size)

    def _get_p_decoder_n_features(self, depth):
        budget = self._budget()
        feats = self.p_decoder_n_features
        total = budget * self.n_components
        while total > budget and feats > 4:
            feats -= 1
            total = self.n_components * LocalCondRNVPDecoder.get_param_count(depth, feats, self.g_latent_space_size)
        return feats, (total > budget, budget, total)

    def _get_decoder_params(self):
        n = self.n_components
        if n == 1 or self.params_reduce_mode == 'none':
            return self.p_decoder_n_flows, self.p_decoder_n_features
        mode = self.params_reduce_mode
        if mode == 'depth_and_feature':
            depth = math.ceil(self.p_decoder_n_flows / math.sqrt(n))
            feats, _ = self._get_p_decoder_n_features(depth)
        elif mode == 'depth_first':
            depth = math.ceil(self.p_decoder_n_flows / n)
            feats, _ = self._get_p_decoder_n_features(depth)
        elif mode == 'feature_first':
            depth = self.p_decoder_n_flows
            feats, (over, budget, total) = self._get_p_decoder_n_features(depth)
            if over:
                while total > budget:
                    depth -= 1
                    total = n * LocalCondRNVPDecoder.get_param_count(depth, feats, self.g_latent_space_size)
        else:
            raise ValueError(f'Unknown params_reduce_mode: {mode}')
        return depth, feats

    # ------------------------------------------------------------------ mixture weights (:104-120)
    def get_weights(self, g_sample, warmup=False):
        if warmup or self.weights_type == 'global_weights':
            return self.mixture_weights_logits.unsqueeze(0).expand(g_sample.shape[0], self.n_components)
        if self.weights_type == 'learned_weights':
            return self.mixture_weights_encoder(g_sample)
        raise ValueError('unknown weights_type %r' % (self.weights_type,))

    # ------------------------------------------------------------------ the hot path
    def flow_stack(self):
        if self._stack is None:
            self._stack = FlowStack([dec.coupling_layers() for dec in self.pc_decoder])
        return self._stack

    def nonfinite_points(self):
        return 0 if self._stack is None else self._stack.take_nonfinite()

    def _base(self, g_sample):
        """Base Gaussian; the reference re-evaluates p_prior once per component (models.py:171 via
        :163-166), which advances its BatchNorm running statistics K times per step with the same batch
        statistics.  Kept, in closed form: K momentum updates with one batch statistic b are
        r_K = (1-m)^K r_0 + (1 - (1-m)^K) b, and b follows from the single evaluation's own update."""
        K = self.n_components
        track = self.training and K > 1 and self.p_decoder_base_type in ('free', 'freevar')
        bns = [m for m in self.p_prior.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm)] if track else []
        before = [(bn.running_mean.clone(), bn.running_var.clone()) for bn in bns]
        mu_b, lv_b = self.base_gaussian(g_sample)
        with torch.no_grad():
            for bn, (rm0, rv0) in zip(bns, before):
                m = bn.momentum
                if m is None or not bn.track_running_stats:
                    for _ in range(K - 1):          # cumulative-average BatchNorm: no closed form worth having
                        self.p_prior(g_sample)
                    break
                keep = (1.0 - m) ** K
                for r, r0 in ((bn.running_mean, rm0), (bn.running_var, rv0)):
                    b = (r - (1.0 - m) * r0) / m
                    # (.data: batch_norm's backward holds the buffer; its own in-kernel update does not bump the
                    # version counter either)
                    r.data.copy_(keep * r0 + (1.0 - keep) * b)
                bn.num_batches_tracked += K - 1
        return mu_b, lv_b

    def decode(self, p_input, g_sample, n_sampled_points, labeled_samples=False, warmup=False):
        logits = self.get_weights(g_sample, warmup)
        stack = self.flow_stack()
        if self.mode == 'training':
            mu_b, lv_b = self._base(g_sample)
            z, ssum, nll = mixture_nll(stack, p_input, g_sample, mu_b, lv_b, logits, self.training,
                                       want_nll=self.fused_nll)
            B, N = p_input.shape[0], p_input.shape[2]
            out = []
            for j in range(self.n_components):
                if self.fused_nll:
                    out.append({'mixture_nll': nll, 'component': j})
                else:
                    out.append({'p_prior_samples': [z[j], p_input],
                                'p_prior_mus': [mu_b.unsqueeze(2).expand(B, 3, N)],
                                'p_prior_logvars': [lv_b.unsqueeze(2).expand(B, 3, N), ssum[j]]})
            if labeled_samples:
                raise ValueError('labeled_samples needs an evaluation mode (model.mode != "training")')
            return out, logits

        # evaluation: sample n_sampled_points per shape (reference asserts B == 1, :146)
        mu_b, lv_b = self._base(g_sample)
        seed = self.sample_seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        stream = (self.sample_stream + 0x9E3779B9 * self.sample_calls) & 0xFFFFFFFF
        self.sample_calls += 1
        samples, labels, _ = sample_mixture(stack, g_sample, mu_b, lv_b, logits, int(n_sampled_points), seed, stream)
        if labeled_samples:
            return samples, labels.to(samples.dtype), logits
        out = []
        for j in range(self.n_components):
            out.append({'p_prior_samples': [samples], 'component_labels': labels, 'component': j})
        return out, logits


class Flow_Mixture_SVR_Model(Flow_Mixture_Model):
    """Single-view reconstruction: image -> ResNet18 -> g0_prior -> prior flow -> decoder."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        from .resnet import resnet18
        self.img_encoder = resnet18(num_classes=self.g_latent_space_size)
        self.g_prior_n_layers = kwargs.get('g_prior_n_layers')
        self.g0_prior = FeatureEncoder(self.g_prior_n_layers, self.g_latent_space_size, self.g_latent_space_size,
                                       deterministic=False, mu_weight_std=0.0033, mu_bias=0.0,
                                       logvar_weight_std=0.033, logvar_bias=0.0)
        self.g0_prior_mus = None
        self.g0_prior_logvars = None

    def encode(self, g_input, images):
        out = {}
        mu0, lv0 = self.g0_prior(self.img_encoder(images))
        out['g_prior_mus'], out['g_prior_logvars'] = [mu0], [lv0]
        if self.mode == 'training':
            feats = torch.max(self.pc_encoder(g_input), dim=2)[0]
            out['g_posterior_mus'], out['g_posterior_logvars'] = self.g_posterior(feats)
            out['g_posterior_samples'] = self.reparameterize(out['g_posterior_mus'], out['g_posterior_logvars'])
            gs, mus, lvs = self.g_prior(out['g_posterior_samples'], mode='inverse')
            out['g_prior_samples'] = gs + [out['g_posterior_samples']]
        elif self.mode == 'reconstruction':
            gs, mus, lvs = self.g_prior(mu0, mode='direct')
            out['g_prior_samples'] = [mu0] + gs
        out['g_prior_mus'] += mus
        out['g_prior_logvars'] += lvs
        return out
