"""GPU parity at the BENCHMARKED sizes (BASELINE.json configs C2..C5), pretrained hot-path weights.

Every iteration is compared on IDENTICAL inputs: the previous depth fed to the CUDA path and to
the oracle is the same tensor (the oracle's own previous result for C2/C3, the CUDA path's for
the sub-cloud checks of C4/C5), exactly as tests/test_gpu_parity.py does on the small golden
inputs.  Tolerances are the ones stated there:

  variance features atol 3e-5 + rtol 1e-5, xyz 1e-6, depth 5e-4 mm, probabilities 5e-5,
  kNN indices bit exact on every point.

C4 / C5: iteration 1 completely, plus ONE of the 16 strided sub-clouds of iteration 3 (the
sub-clouds are independent calls in the reference, model.py:236-267; the oracle evaluates only
that sub-cloud's pixels, oracle.build_point_features(sub=...)).  The oracle runs on the host
cores (torch CPU, 16 threads): about 15 s for a C2 pass.
"""
import pytest
import torch

from oracle import pointflow_oracle as O
from tests.test_gpu_parity import _pf, _run_iteration, _check_stages, sub_to_ref

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCALES = (0.125, 0.25, 0.5)
INTERS = (1.0, 0.75, 0.15)


@pytest.fixture(autouse=True)
def _threads():
    old = torch.get_num_threads()
    torch.set_num_threads(min(16, old))  # the oracle's CPU kernels are slower with 100+ threads
    yield
    torch.set_num_threads(old)


def _inputs(H, W, V, seed):
    from pointmvsnet_b200.synthetic import make_pointflow_inputs
    return make_pointflow_inputs(H, W, V, 1, 96, seed=seed)


def _full_iterations(cpu, golden_weights, golden_params, iters, check_knn_it=None):
    """iterations `iters` (0-based, contiguous from 0) with the oracle's previous depth as input"""
    pf = _pf(golden_weights)
    depth = cpu["coarse_depth"]
    for it in iters:
        s, isc = SCALES[it], INTERS[it]
        res, prob, stg, d_gpu, p_gpu = _run_iteration(pf, cpu, depth, s, isc, it, golden_params)
        dbg = _check_stages(pf, stg, 1)
        derr = (d_gpu - res).abs().max().item()
        assert derr <= 5e-4, (it, derr)
        assert torch.allclose(p_gpu, prob, atol=5e-5, rtol=0), (it, (p_gpu - prob).abs().max())
        if check_knn_it == it:
            # the fused path's neighbour lists of ALL sub-clouds against the oracle's get_knn_3d
            S, hs, ws, N = dbg["S"], dbg["hs"], dbg["ws"], dbg["N"]
            r = int(round(S ** 0.5))
            x7 = stg["xyz"].view(1, 3, 5, hs, r, ws, r)
            for i in range(r):
                for j in range(r):
                    want = O.knn3d(x7[:, :, :, :, i, :, j].contiguous(), 5, 16)
                    got = dbg["idx"][i * r + j].cpu().long()
                    assert torch.equal(got, want), ("kNN", it, i, j, (got != want).float().mean())
        depth = res
    return pf


def test_c2_all_iterations_full_size_vs_oracle(golden_weights, golden_params):
    """BASELINE C2 (640x512, V=4): iterations 1, 2 and 3 (25 600 / 102 400 / 409 600 points),
    every stage, plus the 409 600-point kNN of iteration 3 bit-exact against the oracle."""
    _full_iterations(_inputs(512, 640, 4, 0), golden_weights, golden_params, (0, 1, 2), check_knn_it=2)


def test_c3_all_iterations_full_size_vs_oracle(golden_weights, golden_params):
    """BASELINE C3 (640x512, V=6 = 5 source views, variance aggregation)."""
    _full_iterations(_inputs(512, 640, 6, 3), golden_weights, golden_params, (0, 1, 2))


def _sub_cloud_check(H, W, V, seed, sub, golden_weights, golden_params):
    """iteration 1 in full, iterations 2 on the GPU only, then sub-cloud `sub` of iteration 3"""
    cpu = _inputs(H, W, V, seed)
    pf = _full_iterations(cpu, golden_weights, golden_params, (0,))
    gpu = {k: ([t.to(DEV) for t in v] if isinstance(v, list) else (v.to(DEV) if torch.is_tensor(v) else v))
           for k, v in cpu.items()}
    with torch.no_grad():
        depth = gpu["coarse_depth"]
        for it in (0, 1):
            depth, _ = pf(depth, (INTERS[it] * gpu["depth_interval"]), SCALES[it], it,
                          feature_pyramids=gpu["pyramids"], cam_params_list=gpu["cam_params_list"],
                          mean=gpu["mean"], std=gpu["std"], img_hw=cpu["img_hw"])
        depth2 = depth.clone()
        d3, p3 = pf(depth2, (INTERS[2] * gpu["depth_interval"]), SCALES[2], 2, feature_pyramids=gpu["pyramids"],
                    cam_params_list=gpu["cam_params_list"], mean=gpu["mean"], std=gpu["std"], img_hw=cpu["img_hw"])
        torch.cuda.synchronize()
        i, j = sub
        res, prob, stg = O.point_flow(depth2.cpu(), INTERS[2] * cpu["depth_interval"], SCALES[2], cpu["pyramids"],
                                      cpu["cam_params_list"], cpu["mean"], cpu["std"], cpu["img_hw"], golden_params,
                                      return_stages=True, sub=sub)
    dbg = pf.debug_stages()
    S, hs, ws = dbg["S"], dbg["hs"], dbg["ws"]
    r = int(round(S ** 0.5))
    s = i * r + j
    feat = dbg["feature"][s].cpu().view(1, 5, hs, ws, 136).permute(0, 4, 1, 2, 3)
    assert torch.allclose(feat[:, :112], stg["feature"][:, :112], atol=3e-5, rtol=1e-5)
    assert torch.allclose(feat[:, 112:], stg["feature"][:, 112:], atol=1e-6)
    xyz = dbg["xyz"][s].cpu().view(1, 3, 5, hs, ws)
    assert torch.allclose(xyz, stg["xyz"], atol=1e-6)
    want_idx = O.knn3d(stg["xyz"], 5, 16)
    assert torch.equal(dbg["idx"][s].cpu().long(), want_idx)
    d_sub = d3.cpu()[:, :, i::r, j::r]
    p_sub = p3.cpu()[:, :, i::r, j::r]
    assert (d_sub - res).abs().max().item() <= 5e-4, (d_sub - res).abs().max()
    assert torch.allclose(p_sub, prob, atol=5e-5, rtol=0)


def test_c4_iteration1_and_one_sub_cloud_of_iteration3(golden_weights, golden_params):
    """BASELINE C4 (1280x960, V=4): sub-grid 160x120, 96 000 points per sub-cloud."""
    _sub_cloud_check(960, 1280, 4, 5, (1, 2), golden_weights, golden_params)


def test_c5_iteration1_and_one_sub_cloud_of_iteration3(golden_weights, golden_params):
    """BASELINE C5 (1600x1184, V=6): sub-grid 200x148 (ragged for every tile size), 148 000 points."""
    _sub_cloud_check(1184, 1600, 6, 7, (3, 0), golden_weights, golden_params)
