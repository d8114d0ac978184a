#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): one pass on a 64x128,
3-view input plus the stand-alone operators."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pointmvsnet_b200.point_flow import PointFlow, PointFlowPass
from pointmvsnet_b200.parallel import state_dict_from_params
from pointmvsnet_b200.synthetic import make_pointflow_inputs, make_flow_params
from pointmvsnet_b200.utils.torch_utils import get_knn_3d
from pointmvsnet_b200.utils.feature_fetcher import FeatureFetcher
from pointmvsnet_b200.networks import EdgeConv
dev = torch.device("cuda:0")
inp = make_pointflow_inputs(72, 136, 3, 1, 48, seed=0, device=dev)   # 9 x 17 sub-grid: ragged everywhere
pf = PointFlow().to(dev); pf.load_state_dict(state_dict_from_params(make_flow_params(seed=1), pf.state_dict())); pf.train()
with torch.no_grad():
    outs = PointFlowPass(pf).run(inp["pyramids"], inp["coarse_depth"], inp["cam_params_list"], inp["depth_interval"],
                                 inp["mean"], inp["std"], inp["img_hw"])
    idx = get_knn_3d(torch.randn(1, 3, 5, 9, 17, device=dev), 5, 16)
    m = EdgeConv(32, 32).to(dev).train()
    y = m(torch.randn(2, 32, 765, device=dev), torch.randint(0, 765, (2, 765, 16), device=dev))
    K = torch.tensor([[9., 0, 5], [0, 9., 4], [0, 0, 1]], device=dev).view(1, 1, 3, 3).expand(1, 2, 3, 3).contiguous()
    f = FeatureFetcher()(torch.randn(1, 2, 4, 9, 11, device=dev), torch.randn(1, 3, 60, device=dev) + torch.tensor([0., 0., 6.], device=dev).view(1, 3, 1), K, None)
torch.cuda.synchronize()
print("ok", outs[-1][0].mean().item(), int(idx.max()), y.abs().mean().item(), f.abs().mean().item())
