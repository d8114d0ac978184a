"""``gather_knn`` autograd function (reference functions/gather_knn.py:10-24)."""
import torch

from . import dgcnn_ext


class GatherKNN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature, index):
        ctx.save_for_backward(index)
        return dgcnn_ext.gather_knn_forward(feature, index)

    @staticmethod
    def backward(ctx, grad_output):
        (knn_inds,) = ctx.saved_tensors
        return dgcnn_ext.gather_knn_backward(grad_output, knn_inds), None


gather_knn = GatherKNN.apply
