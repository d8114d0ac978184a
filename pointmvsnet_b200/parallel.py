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


def shard_sub_clouds(num_sub, rank, world):
    """(first, count) of the sub-clouds of one iteration that `rank` processes: contiguous blocks in the
    reference's (i, j) loop order (model.py:244-245); with fewer sub-clouds than ranks the high ranks idle
    (count 0).  Iteration 1 has a single cloud, which every rank computes (replicated, no exchange)."""
    if num_sub <= 1:
        return 0, num_sub
    ids = shard_views(num_sub, rank, world)
    return (ids[0], len(ids)) if ids else (0, 0)


def gather_view_pyramids(owned, num_views, rank, world):
    """C5 'view-sharded' input: rank r owns the feature pyramids of the views
    ``shard_views(num_views, r, world)`` (in the full system it ran ImageConv on those images); one
    all-gather per level gives every rank all V views - 28*H*W bytes per view, far cheaper than exchanging
    partial feature sums.  ``owned``: list over levels of [B, n_owned, C, h, w] (n_owned may be 0).
    Returns the list over levels of [B, V, C, h, w]."""
    if world == 1:
        return owned
    n_max = (num_views + world - 1) // world
    out = []
    for lvl in owned:
        B, _, C, h, w = lvl.shape
        slot = torch.zeros(B, n_max, C, h, w, dtype=lvl.dtype, device=lvl.device)
        slot[:, :lvl.shape[1]] = lvl
        slots = [torch.empty_like(slot) for _ in range(world)]
        dist.all_gather(slots, slot)
        parts = [slots[r][:, :len(shard_views(num_views, r, world))] for r in range(world)]
        out.append(torch.cat(parts, dim=1).contiguous())
    return out


class SubCloudShardedPass(object):
    """One reference view refined by `world` GPUs (BASELINE config C5; SURVEY.md 8e "sub-cloud units").

    The ratio^2 strided sub-clouds of an iteration are independent calls in the reference
    (model.py:236-267: own kNN graph, own BatchNorm batch statistics, own softmax), so they are the
    units: iteration 1 (one cloud) is replicated on every rank, iteration 2 has 4 units, iteration 3
    has 16.  After an iteration every rank holds the pixels of its own sub-clouds in an otherwise
    zero depth map; one all-reduce(SUM) of that map (1.9 MB at C5) assembles the full map on every
    rank - the only data-path collective, NCCL over NVLink.  Critical path 1 + 1 + 2 of the 21 units
    at 8 ranks: a LATENCY configuration with a 5.25x bound, not the throughput one (that is whole
    views per rank, bench.py default).  BatchNorm running statistics are not updated (each rank
    sees only its sub-clouds; inference does not read them, test.py:58)."""

    def __init__(self, point_flow, rank, world, img_scales=(0.125, 0.25, 0.5), inter_scales=(1.0, 0.75, 0.15)):
        self.pf, self.rank, self.world = point_flow, rank, world
        self.img_scales, self.inter_scales = tuple(img_scales), tuple(inter_scales)
        point_flow.update_running_stats = False

    def plan(self):
        """[(first, count)] per iteration for this rank"""
        from .point_flow import _ratio_for
        return [shard_sub_clouds(_ratio_for(s, True) ** 2, self.rank, self.world) for s in self.img_scales]

    def run(self, pyramids_cl, coarse_depth, cam_params_list, depth_interval, mean, std, img_hw, outs=None):
        """pyramids_cl: channels-last levels of ALL views (PointFlow.pyramids_to_channels_last of the
        gathered pyramids).  Returns the final [B,1,h,w] depth map (identical on every rank)."""
        depth = coarse_depth
        B = coarse_depth.shape[0]
        for i, (s, isc) in enumerate(zip(self.img_scales, self.inter_scales)):
            first, count = self.plan()[i]
            h, w = int(img_hw[0] * s), int(img_hw[1] * s)
            d_out, p_out = outs[i] if outs is not None else (
                torch.empty(B, 1, h, w, device=depth.device), torch.empty(B, 5, h, w, device=depth.device))
            sharded = _num_sub(s) > 1 and self.world > 1
            if sharded:
                d_out.zero_()
            if count > 0:
                self.pf(depth, depth_interval, s, i, interval_scale=isc, feature_pyramids=None,
                        cam_params_list=cam_params_list, mean=mean, std=std, is_test=True, img_hw=img_hw,
                        pyramids_channels_last=pyramids_cl, out=(d_out, p_out),
                        sub_range=(first, count) if sharded else None)
            if sharded:
                dist.all_reduce(d_out)  # disjoint supports: the sum assembles the map
            depth = d_out
        return depth


def _num_sub(scale):
    from .point_flow import _ratio_for
    return _ratio_for(scale, True) ** 2


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
