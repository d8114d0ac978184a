"""Output side of a test pass: what `test.py:76` writes per reference view.

Mirror of the reference's `pointmvsnet/utils/eval_file_logger.py:12-105` — same file names and file
bytes (pinned in `tests/test_host.py` against files written by the reference itself).  The
probability post-processing (`eval_file_logger.py:50-67`) is exposed as `flow_confidence` and runs on
whatever device the probability tensor lives on, so only one [H,W] map crosses PCIe instead of five.
"""
import os.path as osp

import numpy as np
import torch

from .io import mkdir, write_cam_dtu, write_pfm

__all__ = ["eval_file_logger", "flow_confidence", "depth2pts_np", "get_pixel_grids_np", "save_points"]


def flow_confidence(flow_prob):
    """[5,H,W] (or [B,5,H,W]) softmax over the hypotheses -> [H,W] ([B,H,W]) confidence.

    eval_file_logger.py:50-66: the expected hypothesis position e = sum_m p_m * (m - 2) + 2 is formed
    in float64, the probabilities of the two hypotheses that bracket it are added in float32:
    p[floor(e)] + p[min(floor(e) + 1, 4)].  A floor of -1 (e rounding below 0) wraps to the last
    hypothesis exactly as numpy's negative index does in the reference.
    """
    p = flow_prob
    m = p.shape[-3]
    steps = torch.arange(m, dtype=torch.float64, device=p.device) - 2.0
    shape = [1] * p.dim()
    shape[-3] = m
    prod = p.to(torch.float64) * steps.view(shape)
    e = prod.select(-3, 0)
    for i in range(1, m):  # numpy reduces the 5 contiguous products left to right
        e = e + prod.select(-3, i)
    e = e + 2.0
    lo = torch.floor(e).to(torch.int64)
    hi = torch.clamp(lo + 1, 0, m - 1)
    lo = torch.where(lo < 0, lo + m, lo)
    return (torch.gather(p, -3, lo.unsqueeze(-3)) + torch.gather(p, -3, hi.unsqueeze(-3))).squeeze(-3)


def get_pixel_grids_np(height, width):
    """3 x (H*W) homogeneous pixel centres, row-major (eval_file_logger.py:94-103)."""
    x = np.linspace(0.5, width - 0.5, width)
    y = np.linspace(0.5, height - 0.5, height)
    xx, yy = np.meshgrid(x, y)
    return np.stack([xx.reshape(-1), yy.reshape(-1), np.ones(height * width)], axis=0)


def depth2pts_np(depth_map, cam_intrinsic, cam_extrinsic):
    """Back-project a depth map to world points [H*W,3] (eval_file_logger.py:80-91)."""
    grid = get_pixel_grids_np(depth_map.shape[0], depth_map.shape[1])
    cam_points = np.matmul(np.linalg.inv(cam_intrinsic), grid) * np.reshape(depth_map, (1, -1))
    R = cam_extrinsic[:3, :3]
    t = cam_extrinsic[:3, 3:4]
    return np.matmul(np.linalg.inv(R), cam_points - t).transpose()


def save_points(path, points):
    np.savetxt(path, points, delimiter=" ", fmt="%.4f")


def _scaled_cam(ref_cam, rows_out, rows_in):
    cam = ref_cam.copy()
    cam[1, :2, :3] *= (float(rows_out) / float(rows_in))
    return cam


def eval_file_logger(data_batch, preds, ref_img_path, folder):
    """Write <eval>/<folder>/<scene>/ files of one reference view (eval_file_logger.py:12-77):
    %08d_init.pfm, %08d_init_prob.pfm, %08d.jpg, cam_%08d_init.txt, and per flow stage
    %08d_flowN.pfm, cam_%08d_flowN.txt, %08d_flowNpts.xyz, %08d_flowN_prob.pfm."""
    import cv2  # image writer of the reference (eval_file_logger.py:40)

    parts = ref_img_path.split("/")
    scene_folder = osp.join("/".join(parts[:-3]), folder, parts[-2])
    if not osp.isdir(scene_folder):
        mkdir(scene_folder)
        print("**** {} ****".format(parts[-2]))
    out_index = int(parts[-1][5:8]) - 1  # "rect_012_..." -> 11

    ref_cam = data_batch["cam_params_list"].cpu().numpy()[0, 0]
    init_depth = preds["coarse_depth_map"].cpu().numpy()[0, 0]
    init_prob = preds["coarse_prob_map"].cpu().numpy()[0, 0]
    ref_image = data_batch["ref_img"][0].cpu().numpy()

    write_pfm(scene_folder + "/%08d_init.pfm" % out_index, init_depth)
    write_pfm(scene_folder + "/%08d_init_prob.pfm" % out_index, init_prob)
    cv2.imwrite(scene_folder + "/%08d.jpg" % out_index, ref_image)
    write_cam_dtu(scene_folder + "/cam_%08d_init.txt" % out_index,
                  _scaled_cam(ref_cam, init_depth.shape[0], ref_image.shape[0]))

    for k in preds.keys():
        if "flow" not in k:
            continue
        if "prob" in k:
            conf = flow_confidence(preds[k][0]).cpu().numpy()
            write_pfm(scene_folder + "/{:08d}_{}.pfm".format(out_index, k), conf)
        else:
            depth = preds[k][0, 0].cpu().numpy()
            write_pfm(scene_folder + "/{:08d}_{}.pfm".format(out_index, k), depth)
            cam = _scaled_cam(ref_cam, depth.shape[0], ref_image.shape[0])
            write_cam_dtu(scene_folder + "/cam_{:08d}_{}.txt".format(out_index, k), cam)
            save_points(osp.join(scene_folder, "{:08d}_{}pts.xyz".format(out_index, k)),
                        depth2pts_np(depth, cam[1][:3, :3], cam[0]))
