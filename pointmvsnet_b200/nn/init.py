"""Parameter initialisers under the reference's names (nn/init.py:4-24): BatchNorm affine
parameters start at (1, 0); convolution / linear weights are Xavier-uniform with zero bias."""
import torch


def _fill(module, weight_fn):
    weight, bias = getattr(module, "weight", None), getattr(module, "bias", None)
    with torch.no_grad():
        if weight is not None:
            weight_fn(weight)
        if bias is not None:
            bias.zero_()


def init_bn(module):
    _fill(module, lambda w: w.fill_(1.0))


def init_uniform(module):
    _fill(module, torch.nn.init.xavier_uniform_)
