"""Structured-grid kNN (reference utils/torch_utils.py:16-61)."""
import random

import numpy as np
import torch

from .._lib import lib, check, stream_ptr, ptr, require_cuda, f32c


def set_random_seed(seed):
    """reference utils/torch_utils.py:7-13"""
    if seed < 0:
        return
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def get_knn_3d(xyz, kernel_size=5, knn=20):
    """xyz [B,3,D,H,W] -> LongTensor [B, D*H*W, knn]: the knn nearest of the
    kernel_size^3 window candidates (zero padding, global index clamp), nearest first,
    ties by candidate id."""
    require_cuda(xyz)
    if xyz.dim() != 5 or xyz.shape[1] != 3:
        raise RuntimeError("get_knn_3d: xyz must be [B,3,D,H,W]")
    assert kernel_size % 2 == 1
    x = f32c(xyz)
    B, _, D, H, W = x.shape
    idx = torch.empty(B, D * H * W, knn, device=x.device, dtype=torch.int64)
    with torch.cuda.device(x.device):
        check(lib.pmvs_knn3d(ptr(x), ptr(idx), None, B, D, H, W, kernel_size, knn, stream_ptr()))
    return idx
