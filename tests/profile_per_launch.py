#!/usr/bin/env python
"""Per-launch CUDA-event times of one eager PointFlow pass (library profiling switch), for comparing kernel options:
    python tests/profile_per_launch.py C2 "edge=1" "edge=2"
prints, per option set, the launches of each of the three iterations (median of 5 passes, L2 flushed before each)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pointmvsnet_b200 import _lib  # noqa: E402
from pointmvsnet_b200.point_flow import PointFlow, PointFlowPass  # noqa: E402
from pointmvsnet_b200.parallel import state_dict_from_params  # noqa: E402
from pointmvsnet_b200.synthetic import make_pointflow_inputs, make_flow_params  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
option_sets = sys.argv[2:] or [""]
H, W, V, D = bench.CONFIGS[cfg]
dev = torch.device("cuda:0")
inp = make_pointflow_inputs(H, W, V, 1, D, seed=0, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for opts in option_sets:
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("=")
        _lib.set_option(k, int(v))
    pf = PointFlow().to(dev)
    pf.load_state_dict(state_dict_from_params(make_flow_params(seed=1), pf.state_dict()))
    pf.train()
    pfp = PointFlowPass(pf)
    runs = []
    with torch.no_grad():
        for rep in range(7):
            flush.zero_()
            torch.cuda.synchronize()
            _lib.profile_enable(rep >= 2)
            pfp.run(inp["pyramids"], inp["coarse_depth"], inp["cam_params_list"], inp["depth_interval"], inp["mean"],
                    inp["std"], inp["img_hw"])
            torch.cuda.synchronize()
            if rep >= 2:
                _lib.profile_enable(False)
                runs.append(_lib.profile_collect())
    n = len(runs[0])
    med = [(runs[0][i][0], sorted(r[i][1] for r in runs)[len(runs) // 2]) for i in range(n)]
    per_it = (n - 3) // 3
    print("== options [%s]: %d launches, pass %.1f us" % (opts, n, 1e3 * sum(m for _, m in med)))
    for it in range(3):
        seg = med[3 + it * per_it: 3 + (it + 1) * per_it]
        print("  it.%d %.1f us: " % (it + 1, 1e3 * sum(m for _, m in seg)) + "  ".join("%s %.1f" % (nm, 1e3 * m) for nm, m in seg))
