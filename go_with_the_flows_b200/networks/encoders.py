"""Latent-side encoders (plain PyTorch; dense GEMMs that cuBLAS serves -- SURVEY.md §2 #7).

Mirrors lib/networks/encoders.py: PointNetCloudEncoder :9-28, FeatureEncoder :31-83,
WeightsEncoder :85-89.  Parameter names and init order are identical so reference checkpoints
load with strict=True and a seeded construction reproduces the reference's weights.
"""
from collections import OrderedDict

import torch.nn as nn
import torch.nn.functional as F

from .layers import SharedDot, Swish


class PointNetCloudEncoder(nn.Module):
    def __init__(self, init_n_channels, init_n_features, n_features):
        super().__init__()
        self.init_n_channels = init_n_channels
        self.init_n_features = init_n_features
        self.n_features = n_features
        widths = [init_n_features] + list(n_features)
        blocks = [('init_sd', init_n_channels, init_n_features)]
        blocks += [('sd%d' % i, widths[i], widths[i + 1]) for i in range(len(n_features))]
        self.features = nn.Sequential()
        for name, cin, cout in blocks:
            self.features.add_module(name, SharedDot(cin, cout, 1, bias=False))
            self.features.add_module(name + '_bn', nn.BatchNorm1d(cout))
            self.features.add_module(name + '_relu', nn.ReLU(inplace=True))

    def forward(self, input):
        return self.features(input)


class FeatureEncoder(nn.Module):
    def __init__(self, n_layers, in_features, latent_space_size, deterministic=False, batch_norm=True,
                 mu_weight_std=0.001, mu_bias=0.0, logvar_weight_std=0.01, logvar_bias=0.0, easy_init=False):
        super().__init__()
        self.n_layers = n_layers
        self.in_features = in_features
        self.latent_space_size = latent_space_size
        self.deterministic = deterministic
        self.batch_norm = batch_norm
        self.mu_weight_std = mu_weight_std
        self.mu_bias = mu_bias
        self.logvar_weight_std = logvar_weight_std
        self.logvar_bias = logvar_bias
        self.easy_init = easy_init

        if n_layers > 0:
            self.features = nn.Sequential()
            for i in range(n_layers):
                self.features.add_module('mlp%d' % i, nn.Linear(in_features, in_features, bias=False))
                if batch_norm:
                    self.features.add_module('mlp%d_bn' % i, nn.BatchNorm1d(in_features))
                self.features.add_module('mlp%d_swish' % i, Swish())

        self.mus = self._head('mu_mlp0', mu_weight_std, mu_bias)
        if not deterministic:
            self.logvars = self._head('logvar_mlp0', logvar_weight_std, logvar_bias)

    def _head(self, name, std, bias):
        lin = nn.Linear(self.in_features, self.latent_space_size, bias=True)
        if not self.easy_init:
            lin.weight.data.normal_(std=std)
            nn.init.constant_(lin.bias.data, bias)
        return nn.Sequential(OrderedDict([(name, lin)]))

    def forward(self, input):
        feats = self.features(input) if self.n_layers > 0 else input
        if self.deterministic:
            return self.mus(feats)
        return self.mus(feats), self.logvars(feats)


class WeightsEncoder(FeatureEncoder):
    """Mixture-weight head: log_softmax over the K logits (encoders.py:85-89)."""

    def forward(self, input):
        return F.log_softmax(super().forward(input), dim=1)
