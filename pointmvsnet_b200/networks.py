"""EdgeConv / EdgeConvNoC (reference networks.py:9-81), CUDA-branch semantics.

Same constructor, parameter names (``conv1``, ``conv2``, ``bn``) and forward signature as
the reference, so its checkpoints load unchanged.  The forward runs three sm_100a kernels
(GEMM, gathered-difference statistics, normalise + ReLU + mean over K) on points-major
data; the [B,C,N,K] tensors of the reference are never materialised.  Forward only: the
backward of the fused layer is a later row of the scope table."""
import torch
import torch.nn as nn

from ._lib import lib, check, stream_ptr, ptr, require_cuda, f32c

_SUPPORTED_COUT = (16, 32, 64, 128)


def _edge_layer(mod, feature, knn_inds, concat_central):
    require_cuda(feature, knn_inds)
    if feature.dim() != 3 or knn_inds.dim() != 3:
        raise RuntimeError("EdgeConv: feature must be [B,C,N] and knn_inds [B,N,K]")
    if torch.is_grad_enabled() and (feature.requires_grad or any(p.requires_grad for p in mod.parameters())):
        # inference under torch.no_grad() is the supported mode (test.py:62)
        raise NotImplementedError("pointmvsnet_b200 EdgeConv is forward-only; wrap the call in torch.no_grad()")
    B, cin, N = feature.shape
    K = knn_inds.shape[2]
    cout = mod.conv1.out_channels
    if knn_inds.shape[0] != B or knn_inds.shape[1] != N:
        raise RuntimeError("EdgeConv: knn_inds shape %s does not match feature %s" % (tuple(knn_inds.shape), tuple(feature.shape)))
    if cout not in _SUPPORTED_COUT or cin % 8 != 0 or cin > 224:
        raise RuntimeError("EdgeConv: unsupported channels in=%d out=%d (out in %s, in %% 8 == 0, in <= 224)"
                           % (cin, cout, _SUPPORTED_COUT))
    dev = feature.device
    x = f32c(feature)
    ctot = 2 * cout if concat_central else cout
    with torch.cuda.device(dev):
        st = stream_ptr()
        x_pm = torch.empty(B, N, cin, device=dev, dtype=torch.float32)
        check(lib.pmvs_transpose(ptr(x), ptr(x_pm), B, cin, N, st))
        idx32 = torch.empty(B, N, K, device=dev, dtype=torch.int32)
        ind = knn_inds.contiguous()
        if ind.dtype != torch.int64:
            ind = ind.long()
        check(lib.pmvs_idx64_to_idx32(ptr(ind), ptr(idx32), ind.numel(), st))
        w12, gamma, beta = _layer_params(mod, dev)
        le = torch.empty(B * N, 2 * cout, device=dev, dtype=torch.float32)
        stats = torch.empty(4 * cout, device=dev, dtype=torch.float64)
        out_pm = torch.empty(B, N, ctot, device=dev, dtype=torch.float32)
        rows = B * N
        train = mod.training or not mod.bn.track_running_stats
        if not train:
            rm = mod.bn.running_mean.double()
            rv = mod.bn.running_var.double()
            if concat_central:
                mc, vc, mn, vn = rm[:cout], rv[:cout], rm[cout:], rv[cout:]
            else:
                mc, vc, mn, vn = rm, rv, rm, rv
            stats.copy_(torch.cat([mc * rows, (vc + mc * mc) * rows, mn * (rows * K), (vn + mn * mn) * (rows * K)]))
        check(lib.pmvs_edgeconv_pm(ptr(x_pm), cin, ptr(idx32), ptr(w12), ptr(gamma), ptr(beta), float(mod.bn.eps),
                                   1 if concat_central else 0, 1 if train else 0, ptr(out_pm), ctot, ptr(le),
                                   ptr(stats), 1, rows, N, K, cin, cout, st))
        if train and mod.bn.track_running_stats and mod.bn.running_mean is not None:
            _update_running(mod.bn, stats, cout, rows, K, concat_central)
        out = torch.empty(B, ctot, N, device=dev, dtype=torch.float32)
        check(lib.pmvs_transpose(ptr(out_pm), ptr(out), B, N, ctot, st))
    return out


def _layer_params(mod, dev):
    """[conv1.weight ; conv2.weight] stacked + BN affine, fp32 contiguous on `dev`; rebuilt only when a parameter
    changed (21 calls per pass reuse them)."""
    params = (mod.conv1.weight, mod.conv2.weight, mod.bn.weight, mod.bn.bias)
    key = (str(dev),) + tuple((p.data_ptr(), p._version) for p in params)
    cache = getattr(mod, "_pmvs_params", None)
    if cache is None or cache[0] != key:
        w12 = torch.cat([mod.conv1.weight.detach()[:, :, 0], mod.conv2.weight.detach()[:, :, 0]], dim=0)
        vals = tuple(t.detach().to(device=dev, dtype=torch.float32).contiguous() for t in (w12, mod.bn.weight, mod.bn.bias))
        cache = (key, vals)
        object.__setattr__(mod, "_pmvs_params", cache)
    return cache[1]


def _update_running(bn, stats, cout, rows, K, concat_central):
    """nn.BatchNorm2d train-mode side effect on the [B,C,N,K] tensor the reference feeds it."""
    s = stats.view(4, cout)
    n_corr = float(rows * K)
    mean_c = s[0] / rows
    var_c = (s[1] / rows - mean_c * mean_c).clamp_(min=0) * (n_corr / max(n_corr - 1.0, 1.0))
    mean_n = s[2] / n_corr
    var_n = (s[3] / n_corr - mean_n * mean_n).clamp_(min=0) * (n_corr / max(n_corr - 1.0, 1.0))
    if concat_central:
        mean, var = torch.cat([mean_c, mean_n]), torch.cat([var_c, var_n])
    else:
        mean, var = mean_n, var_n
    bn.num_batches_tracked.add_(1)
    m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked.item())
    bn.running_mean.mul_(1 - m).add_(mean.to(bn.running_mean.dtype), alpha=m)
    bn.running_var.mul_(1 - m).add_(var.to(bn.running_var.dtype), alpha=m)


class EdgeConv(nn.Module):
    """feature [B,in,N], knn_inds [B,N,K] -> [B, 2*out, N]  (networks.py:9-45)"""

    def __init__(self, in_channels, out_channels):
        super(EdgeConv, self).__init__()
        self.conv1 = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.conv2 = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.bn = nn.BatchNorm2d(2 * out_channels)

    def forward(self, feature, knn_inds):
        return _edge_layer(self, feature, knn_inds, True)


class EdgeConvNoC(nn.Module):
    """feature [B,in,N], knn_inds [B,N,K] -> [B, out, N]  (networks.py:48-81)"""

    def __init__(self, in_channels, out_channels):
        super(EdgeConvNoC, self).__init__()
        self.conv1 = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.conv2 = nn.Conv1d(in_channels, out_channels, 1, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, feature, knn_inds):
        return _edge_layer(self, feature, knn_inds, False)
