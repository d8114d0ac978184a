#!/usr/bin/env python
"""Where does a small-K GEMM of the pass spend its time?  Times pmvs_linear_pm (the kernel behind every 1x1 convolution
of the path) on iteration-3-sized inputs of BASELINE C2 (16 groups x 20 480 rows) with the input-BN, the output
statistics and the output width switched on and off.  CUDA events, L2 flushed before each launch, median of 7."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointmvsnet_b200 import _lib  # noqa: E402

dev = torch.device("cuda:0")
G, R = 16, 20480
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
gen = torch.Generator().manual_seed(0)


def run(cin, cout, ldx, bn, stats):
    x = torch.randn(G * R, ldx, generator=gen).to(dev)
    w = (torch.randn(cout, cin, generator=gen) / cin ** 0.5).to(dev)
    gamma = torch.ones(cin, device=dev)
    beta = torch.zeros(cin, device=dev)
    xs = x[:, :cin].double().view(G, R, cin)
    in_stats = torch.cat([xs.sum(1), (xs * xs).sum(1)], dim=1).contiguous()
    y = torch.empty(G * R, cout, device=dev)
    out_stats = torch.zeros(G, 2 * cout, device=dev, dtype=torch.float64)
    st = _lib.stream_ptr()
    ts = []
    for rep in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib.pmvs_linear_pm(x.data_ptr(), ldx, w.data_ptr(), y.data_ptr(), cout, G, R, cin, cout,
                                           in_stats.data_ptr() if bn else None, gamma.data_ptr() if bn else None,
                                           beta.data_ptr() if bn else None, float(R), 1e-5,
                                           out_stats.data_ptr() if stats else None, st))
        e1.record()
        torch.cuda.synchronize()
        if rep >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    mb = G * R * 4 * (cin + cout) / 1e6
    print("cin %3d (ld %3d) -> cout %3d  in-BN %d  out-stats %d : %7.1f us  %6.1f MB  %5.2f TB/s" %
          (cin, ldx, cout, bn, stats, us, mb, mb / us))


for cin, cout, ldx in ((64, 16, 64), (64, 64, 64), (64, 128, 64), (32, 64, 224), (32, 64, 32), (136, 64, 136), (224, 64, 224)):
    for bn in (0, 1):
        for stats in (0, 1):
            run(cin, cout, ldx, bn, stats)
