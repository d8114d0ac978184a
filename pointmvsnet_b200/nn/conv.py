"""Conv1d + BN + ReLU container (reference nn/conv.py:7-41).  Parameter names
(``conv.weight``, ``bn.*``) match the reference so its checkpoints load unchanged; inside
PointFlow the arithmetic is done by the fused sm_100a kernels, this forward is the
stock-library path for stand-alone use."""
from torch import nn
import torch.nn.functional as F

from .init import init_uniform, init_bn


class Conv1d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, relu=True, bn=True, bn_momentum=0.1, **kwargs):
        super(Conv1d, self).__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, bias=(not bn), **kwargs)
        self.bn = nn.BatchNorm1d(out_channels, momentum=bn_momentum) if bn else None
        self.relu = relu
        self.init_weights()

    def forward(self, x):
        x = self.conv(x)
        if self.bn is not None:
            x = self.bn(x)
        if self.relu:
            x = F.relu(x, inplace=True)
        return x

    def init_weights(self):
        init_uniform(self.conv)
        if self.bn is not None:
            init_bn(self.bn)
