#!/usr/bin/env python
"""torchrun worker: the sub-cloud sharded pass (parallel.SubCloudShardedPass, BASELINE C5 style) over WORLD_SIZE
GPUs must give the depth map of the single-GPU pass.  Launched by tests/test_gpu_multi.py (needs >= 2 GPUs)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from pointmvsnet_b200.point_flow import PointFlow, PointFlowPass  # noqa: E402
from pointmvsnet_b200.parallel import SubCloudShardedPass, gather_view_pyramids, shard_views, gather_depth_maps  # noqa: E402
from pointmvsnet_b200.synthetic import make_pointflow_inputs  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
z = np.load(os.path.join(ROOT, "tests", "golden", "flow_weights.npz"))
weights = {k: torch.from_numpy(z[k]) for k in z.files}
cpu = make_pointflow_inputs(296, 400, 5, 1, 96, seed=3)  # 37 x 50 sub-grid: ragged tiles, V = 5 over `world` ranks
V = 5
gpu = {k: ([t.to(dev) for t in v] if isinstance(v, list) else (v.to(dev) if torch.is_tensor(v) else v)) for k, v in cpu.items()}
pf = PointFlow().to(dev)
pf.load_reference_state_dict(weights)
pf.train()
with torch.no_grad():
    own = shard_views(V, rank, world)
    pyr = gather_view_pyramids([lv[:, own].contiguous() for lv in gpu["pyramids"]], V, rank, world)
    assert all(torch.equal(a, b) for a, b in zip(pyr, gpu["pyramids"])), "pyramid all-gather"
    cl = PointFlow.pyramids_to_channels_last(pyr)
    sp = SubCloudShardedPass(pf, rank, world)
    got = sp.run(cl, gpu["coarse_depth"], gpu["cam_params_list"], gpu["depth_interval"], gpu["mean"], gpu["std"],
                 cpu["img_hw"]).clone()
    pf2 = PointFlow().to(dev)
    pf2.load_reference_state_dict(weights)
    pf2.train()
    want = PointFlowPass(pf2).run(gpu["pyramids"], gpu["coarse_depth"], gpu["cam_params_list"], gpu["depth_interval"],
                                  gpu["mean"], gpu["std"], cpu["img_hw"])[-1][0]
    torch.cuda.synchronize()
err = (got - want).abs().max().item()
assert err < 5e-4, err
# the data-parallel collective of the throughput configuration: all-gather of final depth maps
maps = gather_depth_maps(torch.full((1, 1, 4, 6), float(rank), device=dev))
assert [float(m[0, 0, 0, 0]) for m in maps] == [float(r) for r in range(world)]
dist.barrier()
dist.destroy_process_group()
print("SUBCLOUD-OK rank %d of %d: max |sharded - single GPU| = %.2e mm" % (rank, world, err))
