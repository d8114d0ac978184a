"""Hand-over of the depth maps to the fusion stage.

The two steps of the reference's `tools/depthfusion.py` that touch the files this path produces:
`probability_filter` (`depthfusion.py:153-170`, zero the depths whose coarse / flow confidence is
below a threshold) and `mvsnet_to_gipuma` (`depthfusion.py:64-150`, camera -> 3x4 projection text,
depth -> `.dmb`, constant normals, image copy).  Same file names and file bytes as the reference
(pinned in `tests/test_output_formats.py`).  Running the external `fusibile` binary
(`depthfusion.py:173-194`) is out of scope.
"""
import os

import numpy as np

from .io import load_cam_dtu, load_pfm, mkdir, read_gipuma_dmb, write_gipuma_dmb, write_pfm

__all__ = ["probability_filter", "mvsnet_to_gipuma", "mvsnet_to_gipuma_cam", "mvsnet_to_gipuma_dmb",
           "fake_colmap_normal"]


def _resized_to(prob, shape, mode):
    if prob.shape == shape:
        return prob
    import cv2
    return cv2.resize(prob, (shape[1], shape[0]), interpolation=mode)


def probability_filter(scene_folder, init_prob_threshold, flow_prob_threshold, name, view_num, mode):
    """%08d_<name>.pfm -> %08d_<name>_prob_filtered.pfm: depth 0 where the flow confidence
    (%08d_<name>_prob.pfm) or the coarse confidence (%08d_init_prob.pfm) is below its threshold;
    confidence maps of another size are resized with the OpenCV interpolation `mode` first."""
    for v in range(view_num):
        stem = os.path.join(scene_folder, "{:08d}_".format(v))
        depth = load_pfm(stem + name + ".pfm")[0]
        flow_prob = _resized_to(load_pfm(stem + name + "_prob.pfm")[0], depth.shape, mode)
        init_prob = _resized_to(load_pfm(stem + "init_prob.pfm")[0], depth.shape, mode)
        out = depth.copy()
        out[(flow_prob < flow_prob_threshold) | (init_prob < init_prob_threshold)] = 0
        write_pfm(stem + name + "_prob_filtered.pfm", out)


def mvsnet_to_gipuma_dmb(in_path, out_path):
    write_gipuma_dmb(out_path, load_pfm(in_path)[0])


def mvsnet_to_gipuma_cam(in_path, out_path):
    """Camera text -> the 3x4 projection K[R|t] Gipuma reads, one row per line, `str()` per number."""
    with open(in_path) as f:
        cam = load_cam_dtu(f)
    K = cam[1].copy()
    K[3] = 0.0  # the depth-range row is not part of the intrinsic matrix
    P = np.matmul(K, cam[0])[:3]
    with open(out_path, "w") as f:
        for row in P:
            f.write("".join(str(x) + " " for x in row) + "\n")
        f.write("\n")


def fake_colmap_normal(in_depth_path, out_normal_path):
    """Constant unit normal (1,1,1)/sqrt(3) wherever the depth is positive, 0 elsewhere."""
    depth = read_gipuma_dmb(in_depth_path)
    valid = (depth > 0).astype(np.float32)[..., None]
    normal = np.ones(depth.shape + (3,), dtype=depth.dtype) / 1.732050808
    write_gipuma_dmb(out_normal_path, np.float32(normal * valid))


def mvsnet_to_gipuma(scene_folder, gipuma_point_folder, name, view_num):
    """Lay out <point_folder>/{cams,images,2333__%08d/{disp,normals}.dmb} for `fusibile`."""
    import cv2
    cam_folder = os.path.join(gipuma_point_folder, "cams")
    image_folder = os.path.join(gipuma_point_folder, "images")
    mkdir(cam_folder)
    mkdir(image_folder)
    for v in range(view_num):
        mvsnet_to_gipuma_cam(os.path.join(scene_folder, "cam_{:08d}_{}.txt".format(v, name)),
                             os.path.join(cam_folder, "{:08d}.jpg.P".format(v)))
        sub = os.path.join(gipuma_point_folder, "2333__{:08d}".format(v))
        mkdir(sub)
        depth_pfm = os.path.join(scene_folder, "{:08d}_{}_prob_filtered.pfm".format(v, name))
        mvsnet_to_gipuma_dmb(depth_pfm, os.path.join(sub, "disp.dmb"))
        fake_colmap_normal(os.path.join(sub, "disp.dmb"), os.path.join(sub, "normals.dmb"))
        image = cv2.imread(os.path.join(scene_folder, "{:08d}.jpg".format(v)))
        depth = load_pfm(depth_pfm)[0]
        if image.shape[:2] != depth.shape[:2]:
            image = cv2.resize(image, (depth.shape[1], depth.shape[0]), interpolation=cv2.INTER_NEAREST)
        cv2.imwrite(os.path.join(image_folder, "{:08d}.jpg".format(v)), image)
