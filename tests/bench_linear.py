#!/usr/bin/env python
"""Micro-benchmark of the per-point contraction kernels (pmvs_linear_pm) in the three
arithmetic modes, cold L2, CUDA events.  Usage: python tests/bench_linear.py [rows]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointmvsnet_b200 import _lib  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 409600
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for cin, cout, ldx in ((136, 64, 136), (224, 64, 224), (64, 128, 224), (64, 64, 64), (64, 16, 64), (32, 64, 224)):
    x = torch.randn(rows, ldx, device=dev)
    w = torch.randn(cout, cin, device=dev) / cin ** 0.5
    y = torch.empty(rows, cout, device=dev)
    stats = torch.zeros(2 * cout, device=dev, dtype=torch.float64)
    line = "%3dx%-3d ldx %3d:" % (cin, cout, ldx)
    for mode in (0, 1, 3):
        _lib.set_gemm_mode(mode)
        ts = []
        for rep in range(8):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(_lib.lib.pmvs_linear_pm(x.data_ptr(), ldx, w.data_ptr(), y.data_ptr(), cout, 16, rows // 16, cin,
                                               cout, None, None, None, 0.0, 1e-5, stats.data_ptr(), _lib.stream_ptr()))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        t = sorted(ts)[len(ts) // 2]
        gb = rows * 4 * (cin + cout) / 1e9
        line += "  mode%d %.1f us (%.0f GB/s)" % (mode, t * 1e3, gb / (t * 1e-3))
    print(line)
_lib.set_gemm_mode(3)
