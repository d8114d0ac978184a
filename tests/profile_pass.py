#!/usr/bin/env python
"""Two eager (un-captured) PointFlow passes at a BASELINE config, for ncu:
pass 1 warms up, pass 2 is the one to capture (54 launches each)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pointmvsnet_b200.point_flow import PointFlow, PointFlowPass  # noqa: E402
from pointmvsnet_b200.parallel import state_dict_from_params  # noqa: E402
from pointmvsnet_b200.synthetic import make_pointflow_inputs, make_flow_params  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W, V, D = bench.CONFIGS[cfg]
dev = torch.device("cuda:0")
inp = make_pointflow_inputs(H, W, V, 1, D, seed=0, device=dev)
pf = PointFlow().to(dev)
pf.load_state_dict(state_dict_from_params(make_flow_params(seed=1), pf.state_dict()))
pf.train()
pfp = PointFlowPass(pf)
with torch.no_grad():
    for _ in range(passes):
        pfp.run(inp["pyramids"], inp["coarse_depth"], inp["cam_params_list"], inp["depth_interval"], inp["mean"],
                inp["std"], inp["img_hw"])
        torch.cuda.synchronize()
print("done")
