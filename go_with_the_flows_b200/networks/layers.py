"""Point-wise primitives.  Mirrors lib/networks/layers.py (Swish :6-11, SharedDot :13-45)."""
import torch
import torch.nn as nn


class Swish(nn.Module):
    """x * sigmoid(x)  (layers.py:10-11)."""

    def forward(self, x):
        return x * torch.sigmoid(x)


class SharedDot(nn.Module):
    """1x1 'convolution' over points: weight (n_channels, out, in) applied to x (B, in, N).

    Same parameters / init as layers.py:14-38 (kaiming-uniform on the 3-D tensor, so torch's
    fan_in is out*in -- the reference's quirk is kept because checkpoints depend on nothing but
    seeds depend on it).  Inside a coupling layer these weights are consumed by the fused CUDA
    kernels; this module's own forward is the plain matmul used by the PointNet encoder
    (encoders.py:16-24), which is outside the flow hot path.
    """

    def __init__(self, in_features, out_features, n_channels, bias=False, init_weight=None, init_bias=None):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.n_channels = n_channels
        self.init_weight = init_weight
        self.init_bias = init_bias
        self.weight = nn.Parameter(torch.empty(n_channels, out_features, in_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(n_channels, out_features))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        if self.init_weight:
            nn.init.uniform_(self.weight.data, a=-self.init_weight, b=self.init_weight)
        else:
            nn.init.kaiming_uniform_(self.weight.data, a=0.)
        if self.bias is not None:
            nn.init.constant_(self.bias.data, self.init_bias if self.init_bias else 0.)

    def forward(self, input):
        out = torch.matmul(self.weight, input.unsqueeze(1))
        if self.bias is not None:
            out = out + self.bias[None, :, :, None]
        return out.squeeze(1)
