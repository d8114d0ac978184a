#!/usr/bin/env python
"""Experiment: throughput with G independent reference views in flight on one GPU
(G captured pass graphs replayed on G streams)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from pointmvsnet_b200.point_flow import PointFlow, PointFlowPass
from pointmvsnet_b200.parallel import state_dict_from_params
from pointmvsnet_b200.synthetic import make_pointflow_inputs, make_flow_params
dev = torch.device("cuda:0")
H, W, V, D = bench.CONFIGS["C2"]
base = PointFlow().to(dev); base.load_state_dict(state_dict_from_params(make_flow_params(seed=1), base.state_dict())); base.train()
G = 4
pipes = []
with torch.no_grad():
    for g in range(G):
        inp = make_pointflow_inputs(H, W, V, 1, D, seed=g, device=dev)
        pf = PointFlow(flow_edge_conv=base.flow_edge_conv, flow_mlp=base.flow_mlp, update_running_stats=(g == 0)).to(dev); pf.train()
        pipes.append(PointFlowPass(pf).capture(inp))
streams = [torch.cuda.Stream(device=dev) for _ in range(G)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
main = torch.cuda.current_stream(dev)
for g_used in (1, 2, 3, 4):
    times = []
    for rep in range(12):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(main)
        for g in range(g_used):
            streams[g].wait_stream(main)
            with torch.cuda.stream(streams[g]):
                pipes[g].replay()
        for g in range(g_used):
            main.wait_stream(streams[g])
        b.record(main)
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    t = sorted(times[2:])[len(times[2:]) // 2]
    print("views in flight %d: %.3f ms per step -> %.1f iters/s" % (g_used, t, g_used * 3 / (t * 1e-3)))
