"""Multi-GPU checks (NCCL, one process per GPU): skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sub_cloud_sharded_pass_equals_single_gpu_pass():
    n = min(4, torch.cuda.device_count())
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", "29577", os.path.join(ROOT, "tests", "multi_gpu_subcloud_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and res.stdout.count("SUBCLOUD-OK") == n, res.stdout[-2000:] + res.stderr[-4000:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_devices_in_one_process():
    """nn.DataParallel-style use (train.py:177): one process driving two GPUs - the per-device shared-memory opt-in
    of every kernel with > 48 KB of dynamic shared memory (GEMM, EdgeConv tile kernels)."""
    from pointmvsnet_b200.point_flow import PointFlow
    from pointmvsnet_b200.synthetic import make_pointflow_inputs
    from tests.conftest import load_golden
    weights = load_golden("flow_weights.npz")
    cpu = make_pointflow_inputs(64, 128, 3, 1, 48, seed=5)
    outs = []
    for d in (0, 1):
        dev = torch.device("cuda", d)
        pf = PointFlow().to(dev)
        pf.load_reference_state_dict(weights)
        pf.train()
        with torch.no_grad():
            depth, _ = pf(cpu["coarse_depth"].to(dev), (0.75 * cpu["depth_interval"]).to(dev), 0.25, 1,
                          feature_pyramids=[p.to(dev) for p in cpu["pyramids"]], cam_params_list=cpu["cam_params_list"].to(dev),
                          mean=cpu["mean"].to(dev), std=cpu["std"].to(dev), img_hw=cpu["img_hw"])
        torch.cuda.synchronize(dev)
        outs.append(depth.cpu())
    assert torch.allclose(outs[0], outs[1], atol=2e-4)
