"""Drop-in for the reference's pybind module ``pointmvsnet.functions.dgcnn_ext``
(functions/csrc/main.cpp:3-6): same two entry points, same tensor contract
(gather_knn.h:7-13), implemented over the C ABI of libpmvs_b200.so."""
import torch

from .._lib import lib, check, stream_ptr, ptr


def _check_inputs(name, x, xdim, index):
    # the reference checks CUDA-ness and ranks (gather_knn_kernel.cu:34-39, 106-112)
    if not x.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not index.is_cuda:
        raise RuntimeError("index must be a CUDA tensor")
    if x.dim() != xdim:
        raise RuntimeError("%s.dim() does not equal to %d" % (name, xdim))
    if index.dim() != 3:
        raise RuntimeError("index.dim() does not equal to 3")
    if index.dtype != torch.int64:
        raise RuntimeError("index must be int64")
    if index.size(0) != x.size(0):
        raise RuntimeError("index.size(0) does not equal to batch_size")
    if index.size(1) != x.size(2):
        raise RuntimeError("index.size(1) does not equal to num_inst")


def gather_knn_forward(input, index):
    """input [B,C,N] fp32 CUDA, index [B,N,K] int64 -> new tensor [B,C,N,K]."""
    _check_inputs("input", input, 3, index)
    x = input.float().contiguous()
    ind = index.contiguous()
    B, Cc, N = x.shape
    K = ind.shape[2]
    out = torch.empty(B, Cc, N, K, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(lib.pmvs_gather_knn_forward(ptr(x), ptr(ind), ptr(out), B, Cc, N, K, stream_ptr()))
    return out


def gather_knn_backward(grad_output, index, deterministic=True):
    """grad_output [B,C,N,K], index [B,N,K] -> grad_input [B,C,N] (scatter-add).

    ``deterministic=True`` (default): segmented reduction with a fixed summation order
    (``pmvs_gather_knn_backward_det``) - bit-reproducible, unlike the reference's ``atomicAdd`` scatter
    (gather_knn_kernel.cu:50-89).  ``False`` selects the atomic scatter (same sums up to fp32 rounding order)."""
    _check_inputs("grad_output", grad_output, 4, index)
    if index.size(2) != grad_output.size(3):
        raise RuntimeError("index.size(2) does not equal to k")
    g = grad_output.float().contiguous()
    ind = index.contiguous()
    B, Cc, N, K = g.shape
    out = torch.empty(B, Cc, N, device=g.device, dtype=torch.float32)
    with torch.cuda.device(g.device):
        if deterministic:
            nbytes = lib.pmvs_gather_knn_backward_det_workspace_bytes(B, N, K)
            ws = torch.empty(nbytes + 256, device=g.device, dtype=torch.uint8)
            base = ws.data_ptr()
            aligned = (base + 255) & ~255
            check(lib.pmvs_gather_knn_backward_det(ptr(g), ptr(ind), ptr(out), B, Cc, N, K, aligned,
                                                   nbytes + 256 - (aligned - base), stream_ptr()))
        else:
            check(lib.pmvs_gather_knn_backward(ptr(g), ptr(ind), ptr(out), B, Cc, N, K, stream_ptr()))
    return out
