"""Multi-GPU plumbing: one process per GPU (torchrun), reference views sharded over ranks,
one collective -- the all-gather of final depth maps (NCCL over NVLink on GPUs; gloo in the
CPU tests).  The reference's only parallelism is nn.DataParallel over the batch dimension
(train.py:177, test.py:84); sharding whole reference views is its process-per-GPU
equivalent.  PointFlow itself needs no data-path collective: every reference view (and
every sub-cloud, with its own BatchNorm statistics) is independent."""
import torch
import torch.distributed as dist


def shard_views(num_views, rank, world):
    """Contiguous block partition of ``range(num_views)``; remainders go to the low ranks."""
    base, rem = divmod(num_views, world)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def gather_depth_maps(local_depth, out_list=None):
    """all-gather equally shaped depth maps [B,1,h,w] from every rank (model.py:302 output)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [local_depth]
    if out_list is None:
        out_list = [torch.empty_like(local_depth) for _ in range(dist.get_world_size())]
    dist.all_gather(out_list, local_depth.contiguous())
    return out_list


def gather_ragged_depth_maps(local_depths, num_views):
    """all-gather a ragged shard: local_depths [n_local,1,h,w] on every rank (n_local from
    shard_views) -> [num_views,1,h,w] in view order on every rank."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local_depths
    rank = dist.get_rank()
    n_max = (num_views + world - 1) // world
    h, w = local_depths.shape[-2:]
    pad = torch.zeros(n_max, 1, h, w, dtype=local_depths.dtype, device=local_depths.device)
    pad[:local_depths.shape[0]] = local_depths
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    parts = [outs[r][:len(shard_views(num_views, r, world))] for r in range(world)]
    return torch.cat(parts, dim=0)


def state_dict_from_params(params, template):
    """Flat oracle-style parameter dict (synthetic.make_flow_params) -> state_dict with the
    reference's key names; running statistics come from ``template``."""
    sd = dict(template)
    for l in range(3):
        sd["flow_edge_conv.%d.conv1.weight" % l] = params["ec%d_w1" % l]
        sd["flow_edge_conv.%d.conv2.weight" % l] = params["ec%d_w2" % l]
        sd["flow_edge_conv.%d.bn.weight" % l] = params["ec%d_gamma" % l]
        sd["flow_edge_conv.%d.bn.bias" % l] = params["ec%d_beta" % l]
        sd["flow_mlp.0.%d.conv.weight" % l] = params["mlp%d_w" % l]
        sd["flow_mlp.0.%d.bn.weight" % l] = params["mlp%d_gamma" % l]
        sd["flow_mlp.0.%d.bn.bias" % l] = params["mlp%d_beta" % l]
    sd["flow_mlp.1.weight"] = params["mlp3_w"]
    return sd
