"""SharedMLP container (reference nn/mlp.py:45-81), 1-D only (the hot path's flow_mlp)."""
from torch import nn

from .conv import Conv1d


class SharedMLP(nn.ModuleList):
    def __init__(self, in_channels, mlp_channels, ndim=1, bn=True, bn_momentum=0.1):
        super(SharedMLP, self).__init__()
        if ndim != 1:
            raise ValueError("SharedMLP: only ndim=1 is on the PointFlow path")
        self.in_channels = in_channels
        for out_channels in mlp_channels:
            self.append(Conv1d(in_channels, out_channels, 1, relu=True, bn=bn, bn_momentum=bn_momentum))
            in_channels = out_channels
        self.out_channels = in_channels

    def forward(self, x):
        for module in self:
            x = module(x)
        return x
