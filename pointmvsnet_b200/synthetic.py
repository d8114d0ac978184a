"""Synthetic DTU-shaped inputs for the PointFlow hot path (no dataset, no network).

Shapes and constants follow the reference's data contract:
  * batch dict layout          dataset.py:141-149, 298-307
  * cam_params_list [B,V,2,4,4] io.py:31-45  ([...,0] extrinsic, [...,1,:3,:3] K,
    [...,1,3,:] = depth_start, interval, num_depth, depth_end)
  * mean / std constants        dataset.py:26-27
  * focal length 2892.33 at 1600 px width (DTU calibration)
The camera rig (look-at arc around a target at z = 650 mm) is the generator
described in SURVEY.md Appendix A; everything is seeded.
"""
import math

import torch

DTU_MEAN = (1.97145182, -1.52387525, 651.07223895)  # dataset.py:26
DTU_STD = (84.45612252, 93.22252387, 80.08551226)  # dataset.py:27


def make_cameras(batch, views, height, width, num_depth=96, dtype=torch.float32):
    """cam_params_list [B,V,2,4,4] at FULL image resolution (isTest=True convention)."""
    s_int = 4.24 if num_depth == 48 else 2.13  # config.py:28 / configs/dtu_wde3.yaml:12
    cams = torch.zeros(batch, views, 2, 4, 4, dtype=torch.float64)
    f = 2892.33 * width / 1600.0
    for v in range(views):
        ang = 0.12 * v * (1.0 if v % 2 else -1.0)
        c = torch.tensor([650.0 * math.sin(ang), 20.0 * v, 650.0 - 650.0 * math.cos(ang)], dtype=torch.float64)
        target = torch.tensor([0.0, 0.0, 650.0], dtype=torch.float64)
        up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)
        z = target - c
        z = z / z.norm()
        x = torch.linalg.cross(up, z)
        x = x / x.norm()
        y = torch.linalg.cross(z, x)
        R = torch.stack([x, y, z], dim=0)
        t = -R @ c
        ext = torch.eye(4, dtype=torch.float64)
        ext[:3, :3] = R
        ext[:3, 3] = t
        K = torch.tensor([[f, 0.0, width / 2.0], [0.0, f, height / 2.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
        cams[:, v, 0] = ext
        cams[:, v, 1, :3, :3] = K
        cams[:, v, 1, 3, 0] = 425.0
        cams[:, v, 1, 3, 1] = 2.5 * s_int
        cams[:, v, 1, 3, 2] = num_depth
        cams[:, v, 1, 3, 3] = 425.0 + 2.5 * s_int * (num_depth - 1)
    return cams.to(dtype)


def make_pointflow_inputs(height=512, width=640, views=4, batch=1, num_depth=96, seed=0,
                          device="cpu", pin_memory=False):
    """Inputs of the point_flow loop (model.py:297-303) for a PointFlow-only run:
    pyramids conv1/conv2/conv3 ~ N(0,1) in the reference's [B,V,C,h,w] layout
    (model.py:133-148), a smooth coarse depth at (H/8, W/8), cameras, mean, std."""
    g = torch.Generator().manual_seed(seed)
    pyr = [
        torch.randn(batch, views, 16, height // 2, width // 2, generator=g),
        torch.randn(batch, views, 32, height // 4, width // 4, generator=g),
        torch.randn(batch, views, 64, height // 8, width // 8, generator=g),
    ]
    h8, w8 = height // 8, width // 8
    yy = torch.linspace(0, 1, h8).view(h8, 1)
    xx = torch.linspace(0, 1, w8).view(1, w8)
    smooth = torch.sin(2.3 * yy + 0.4) * torch.cos(3.1 * xx - 0.7)
    depth = 650.0 + 40.0 * smooth + 2.0 * torch.randn(h8, w8, generator=g)
    depth = depth.view(1, 1, h8, w8).repeat(batch, 1, 1, 1)
    if batch > 1:
        depth = depth + 3.0 * torch.randn(batch, 1, 1, 1, generator=g)
    cams = make_cameras(batch, views, height, width, num_depth)
    out = {
        "pyramids": pyr,
        "coarse_depth": depth.contiguous(),
        "cam_params_list": cams,
        "mean": torch.tensor(DTU_MEAN).view(1, 3).repeat(batch, 1),
        "std": torch.tensor(DTU_STD).view(1, 3).repeat(batch, 1),
        "depth_interval": cams[:, 0, 1, 3, 1].clone(),
        "img_hw": (height, width),
    }

    def mv(x):
        if isinstance(x, torch.Tensor):
            if pin_memory and device == "cpu":
                return x.pin_memory()
            return x.to(device)
        return x

    out["pyramids"] = [mv(p) for p in pyr]
    for k in ("coarse_depth", "cam_params_list", "mean", "std", "depth_interval"):
        out[k] = mv(out[k])
    return out


def make_flow_params(seed=1):
    """Random-init hot-path weights with the reference's shapes:
    xavier-uniform convs (nn/init.py:17-24), BN gamma/beta perturbed from (1, 0)
    so that the affine part is exercised."""
    g = torch.Generator().manual_seed(seed)

    def xavier(cout, cin):
        bound = math.sqrt(6.0 / (cin + cout))
        return (torch.rand(cout, cin, 1, generator=g) * 2 - 1) * bound

    p = {}
    for l, (cin, cout, cbn) in enumerate(((136, 32, 32), (32, 32, 64), (64, 64, 128))):
        p["ec%d_w1" % l] = xavier(cout, cin)
        p["ec%d_w2" % l] = xavier(cout, cin)
        p["ec%d_gamma" % l] = 1.0 + 0.1 * torch.randn(cbn, generator=g)
        p["ec%d_beta" % l] = 0.1 * torch.randn(cbn, generator=g)
    for i, (cin, cout) in enumerate(((224, 64), (64, 64), (64, 16))):
        p["mlp%d_w" % i] = xavier(cout, cin)
        p["mlp%d_gamma" % i] = 1.0 + 0.1 * torch.randn(cout, generator=g)
        p["mlp%d_beta" % i] = 0.1 * torch.randn(cout, generator=g)
    p["mlp3_w"] = xavier(1, 16)
    return p
